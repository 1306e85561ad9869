"""include/rlitg.hpp -- the C++ host mirror of torm::TormIKSolver / torm::TormOptimizer over the C ABI.
CPU tier: the header and its test program compile and link with plain g++ (no CUDA toolkit on the include or link
line).  GPU tier: the program runs on the device and checks itself against the CPU oracle."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "host_mirror_test.cpp")
OUT_DIR = os.path.join(ROOT, "tests", "cpp", "_build")
EXE = os.path.join(OUT_DIR, "host_mirror_test")
LIBDIR = os.path.join(ROOT, "rl-itg_b200")
ORCDIR = os.path.join(ROOT, "oracle", "_build")


def build_exe():
    from oracle import oracle as O
    O.build()
    os.makedirs(OUT_DIR, exist_ok=True)
    deps = [SRC, os.path.join(ROOT, "include", "rlitg.hpp"), os.path.join(ROOT, "include", "rlitg.h"),
            os.path.join(LIBDIR, "librlitg.so")]
    if os.path.exists(EXE) and all(os.path.getmtime(EXE) >= os.path.getmtime(d) for d in deps):
        return EXE
    cmd = ["/usr/bin/g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
           "-I", os.path.join(ROOT, "oracle"), SRC, "-o", EXE, "-L", LIBDIR, "-lrlitg", "-L", ORCDIR, "-loracle",
           "-Wl,-rpath," + LIBDIR, "-Wl,-rpath," + ORCDIR]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    return EXE


def test_cpp_host_mirror_compiles_without_cuda_toolkit():
    assert os.path.exists(os.path.join(LIBDIR, "librlitg.so")), "build librlitg.so first (__graft_entry__.build())"
    exe = build_exe()
    assert os.access(exe, os.X_OK)


@pytest.mark.gpu
def test_cpp_host_mirror_runs(cuda_device):
    exe = build_exe()
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.strip().endswith("OK"), r.stdout
