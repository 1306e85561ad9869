#!/usr/bin/env python
"""Builds tests/golden/*.npz from the reference's own DATA fixtures (never its sources).

Run in the dev container, where /root/reference is mounted:  python tests/golden/make_golden.py
The GPU box has no /root/reference, so the vectors are committed.  What is harvested:

  fk_pairs.npz
    hello_init_q / hello_init_pose   pirl/launch/cfg/target_path/hello.yaml:2-3 (17-digit known answer)
    named_q / named_pose / named_id  torm/launch/config/start_conf/{hello,zig,rotation,kaist,sgvr,square,s}_config.csv
                                     rows, each an IK solution of row 0 of torm/launch/config/path/<name>
    rand_q / rand_pose               start_conf/random_<n>_config.csv row i <-> path/random_<n> row i, n = 0..99
                                     (62,610 pairs written by torm/src/pirl_problem_generation.cpp:40-66)
    obs_q / obs_pose                 start_conf/random_obs_<s>_<p>_config.csv rows <-> row 0 of path/random_obs_<s>_<p>
                                     (every 4th file to keep the fixture small: 5,000 of the 20,000 pairs)
  init_traj.npz
    <prob>_path, <prob>_<method>     torm/launch/config/path/<prob> and initial_traj/<prob>-<method>-0-0.csv
                                     for prob in hello, zig, rotation, square
Poses are stored [x, y, z, qx, qy, qz, qw]; the path files hold x,y,z;w,x,y,z (torm/src/torm_problem.cpp:169-174).
"""
import glob
import os

import numpy as np

REF = "/root/reference"
CFG = os.path.join(REF, "torm/launch/config")
OUT = os.path.dirname(os.path.abspath(__file__))


def load_path(name):
    rows = []
    with open(os.path.join(CFG, "path", name)) as f:
        for line in f:
            line = line.strip()
            if not line:
                continue
            p, q = line.split(";")
            x, y, z = [float(v) for v in p.split(",")]
            w, qx, qy, qz = [float(v) for v in q.split(",")]
            rows.append([x, y, z, qx, qy, qz, w])
    return np.array(rows)


def load_csv(path):
    rows = []
    with open(path) as f:
        for line in f:
            line = line.strip()
            if line:
                rows.append([float(v) for v in line.split(",")])
    return np.array(rows)


def main():
    out = {}
    # hello.yaml known answer
    with open(os.path.join(REF, "pirl/launch/cfg/target_path/hello.yaml")) as f:
        lines = f.read().splitlines()
    q = [float(v) for v in lines[1].split("[")[1].split("]")[0].split(",")]
    pose = [float(v) for v in lines[2].split("[")[1].split("]")[0].split(",")]
    out["hello_init_q"] = np.array(q)
    out["hello_init_pose_raw"] = np.array(pose)  # as written in the yaml (layout checked in the test)

    nq, npz, nid = [], [], []
    for k, name in enumerate(["hello", "zig", "rotation", "kaist", "sgvr", "square", "s"]):
        qs = load_csv(os.path.join(CFG, "start_conf", name + "_config.csv"))
        p0 = load_path(name)[0]
        nq.append(qs); npz.append(np.tile(p0, (qs.shape[0], 1))); nid.append(np.full(qs.shape[0], k))
    out["named_q"] = np.concatenate(nq); out["named_pose"] = np.concatenate(npz); out["named_id"] = np.concatenate(nid)

    rq, rp = [], []
    for n in range(100):
        qs = load_csv(os.path.join(CFG, "start_conf", "random_%d_config.csv" % n))
        ps = load_path("random_%d" % n)
        assert qs.shape[0] == ps.shape[0], (n, qs.shape, ps.shape)
        rq.append(qs); rp.append(ps)
    out["rand_q"] = np.concatenate(rq); out["rand_pose"] = np.concatenate(rp)

    oq, op = [], []
    files = sorted(glob.glob(os.path.join(CFG, "start_conf", "random_obs_*_config.csv")))
    for fpath in files[::4]:
        name = os.path.basename(fpath)[: -len("_config.csv")]
        qs = load_csv(fpath)
        p0 = load_path(name)[0]
        oq.append(qs); op.append(np.tile(p0, (qs.shape[0], 1)))
    out["obs_q"] = np.concatenate(oq); out["obs_pose"] = np.concatenate(op)
    np.savez_compressed(os.path.join(OUT, "fk_pairs.npz"), **out)
    print({k: v.shape for k, v in out.items()})

    tr = {}
    for prob in ["hello", "zig", "rotation", "square"]:
        tr[prob + "_path"] = load_path(prob)
        for method in ["torm_test", "torm_jli_test", "pirl_rl_test_0", "pirl_bc_test"]:
            fpath = os.path.join(CFG, "initial_traj", "%s-%s-0-0.csv" % (prob, method))
            if os.path.exists(fpath):
                tr["%s_%s" % (prob, method)] = load_csv(fpath)
    np.savez_compressed(os.path.join(OUT, "init_traj.npz"), **tr)
    print({k: v.shape for k, v in tr.items()})


if __name__ == "__main__":
    main()
