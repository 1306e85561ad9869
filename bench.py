#!/usr/bin/env python
"""bench.py -- headline benchmark of the rlitg hot path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Headline workload (BASELINE.json configs[1]): batched FK + geometric Jacobian over 1,048,576 random Fetch
configurations per GPU (D = 7, FP32, [B,7] in -> pos [B,3], rot [B,9], J [B,6,7] out).  A "step" is one pass
over that batch.  `value` = configs/s over all ranks with inputs resident in HBM; `e2e` = the same through the
host-buffer C-ABI entry point (pinned host buffers, H2D + D2H inside the timed region).  One JSON line on rank 0.
The TORM cost+gradient and PIRL env-step workloads are reported in the same line under "also".
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B_PER_GPU = 1 << 20
D = 7
BYTES_PER_CONFIG = 4 * (D + 12 + 6 * D)  # 244 B: read q, write p[3] R[9] J[6x7]


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """samples SM clock + throttle reasons of one GPU through NVML while the timed region runs"""

    def __init__(self, index, period=0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._halt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
                 nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake"}
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(self.period)

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def synth_q(B, seed, torch):
    """q_j ~ U(ll_j, ul_j) (pirl/launch/pirl.yaml:11-12), seeded"""
    ll = torch.tensor([-1.6056, -1.221, -3.141592, -2.251, -3.141592, -2.16, -3.141592])
    ul = torch.tensor([1.6056, 1.518, 3.141592, 2.251, 3.141592, 2.16, 3.141592])
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return (torch.rand(B, D, generator=g) * (ul - ll) + ll).float()


def cpu_reference_fk_jac(seconds_target=12.0, threads=None):
    """times the CPU oracle (FP64 restatement of the reference's KDL calls) on a bounded sample"""
    import numpy as np
    from oracle import oracle as O
    # all host threads this process may use (torchrun exports OMP_NUM_THREADS=1: override it explicitly)
    O.set_num_threads(threads or len(os.sched_getaffinity(0)))
    cores = O.num_threads()
    n = 1 << 18
    rng = np.random.default_rng(0)
    ll = np.array([-1.6056, -1.221, -3.141592, -2.251, -3.141592, -2.16, -3.141592]); ul = -ll.copy(); ul[1] = 1.518
    q = rng.uniform(ll, ul, size=(n, D))
    O.fk_jac(q[:1024])
    t0 = time.perf_counter(); O.fk_jac(q); t1 = time.perf_counter()
    reps = max(1, int(seconds_target / max(t1 - t0, 1e-6)))
    reps = min(reps, 2048)
    t0 = time.perf_counter()
    for _ in range(reps):
        O.fk_jac(q)
    dt = time.perf_counter() - t0
    return {"value": n * reps / dt, "unit": "configs/s", "cores": cores, "kind": "port",
            "sample": "%d x %d random configs, FK+Jacobian, FP64 oracle (OpenMP), %.1f s" % (reps, n, dt)}, n


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warm = args.steps, args.warmup
    import numpy as np
    from oracle import oracle as O
    O.set_num_threads(len(os.sched_getaffinity(0)))  # torchrun exports OMP_NUM_THREADS=1: use every host thread
    cores = O.num_threads()
    n = 1 << 18  # bounded sample of the 1M-config workload per step
    rng = np.random.default_rng(0)
    ll = np.array([-1.6056, -1.221, -3.141592, -2.251, -3.141592, -2.16, -3.141592]); ul = -ll.copy(); ul[1] = 1.518
    q = rng.uniform(ll, ul, size=(n, D))
    for _ in range(warm):
        O.fk_jac(q)
    t0 = time.perf_counter()
    for _ in range(steps):
        O.fk_jac(q)
    dt = time.perf_counter() - t0
    v = n * steps / dt
    line = {"impl": "reference", "metric": "fk_jacobian_configs_per_s", "value": v, "unit": "configs/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": 1e3 * dt / steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "fk_jacobian_1M_fetch_configs", "D": D, "sample_per_step": n,
                       "note": "reference CPU path = FP64 oracle port of the KDL calls (reference not buildable here)"},
            "cpu_baseline": {"value": v, "unit": "configs/s", "cores": cores, "kind": "port",
                             "sample": "%d configs per step" % n},
            "e2e": {"value": v, "unit": "configs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="skip the secondary workloads")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from rlitg_b200 import kinematics as K, _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    B = B_PER_GPU
    q_host = synth_q(B, 1234 + rank, torch)
    q = q_host.to(dev)
    out = {"pos": torch.empty(B, 3, device=dev), "rot": torch.empty(B, 3, 3, device=dev),
           "jac": torch.empty(B, 6, D, device=dev)}

    def step():
        K.fk_jac(q, out=out)

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    n0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    launches = _lib.launch_count() - n0
    ms_local = e0.elapsed_time(e1)
    # keep the GPU busy a little longer so that NVML gets samples under load on very short runs
    t_end = time.time() + 0.3
    while time.time() < t_end:
        step()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ms = max_over_ranks(ms_local)
    ms_per_step = ms / args.steps
    value = world * B * args.steps / (ms * 1e-3)

    peak, peak_src = measured_peaks()
    achieved = BYTES_PER_CONFIG * B / (ms_local / args.steps * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "kernel": "fk_jac_persist_kernel<float,7,aos,2 stages>", "bytes_per_launch": BYTES_PER_CONFIG * B,
                "peak_source": peak_src}
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof):
        try:
            roofline["traffic"] = json.load(open(prof)).get("fk_jac_kernel")
        except Exception:  # noqa: BLE001
            pass

    # ---- end to end: host buffers through rlitg_fk_jac_host (H2D q, kernel, D2H pos/rot/J)
    from rlitg_b200.dist import bind_host_to_gpu
    prev_affinity = None if os.environ.get("RLITG_BENCH_NO_NUMA") else bind_host_to_gpu(local)  # host buffers next to this rank's GPU (NUMA); undone after the e2e leg
    hq = q_host.pin_memory()
    hpos = torch.empty(B, 3).pin_memory(); hrot = torch.empty(B, 3, 3).pin_memory(); hjac = torch.empty(B, 6, D).pin_memory()
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        K.fk_jac_host(hq, pos=hpos, rot=hrot, jac=hjac)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        K.fk_jac_host(hq, pos=hpos, rot=hrot, jac=hjac)
    torch.cuda.synchronize()
    dt = max_over_ranks(time.perf_counter() - t0)
    e2e = {"value": world * B * e2e_steps / dt, "unit": "configs/s", "h2d_bytes_per_step": B * D * 4,
           "d2h_bytes_per_step": B * (3 + 9 + 6 * D) * 4, "steps": e2e_steps,
           "api": "rlitg_fk_jac_host (pinned host buffers)", "numa_bound": prev_affinity is not None}
    if prev_affinity is not None:
        os.sched_setaffinity(0, prev_affinity)

    also = {}
    if not args.no_also:
        try:
            import bench_also
            also = bench_also.run(dev, rank, world, max_over_ranks, barrier)
        except ImportError:
            also = {}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu, _ = cpu_reference_fk_jac()

    if rank == 0:
        line = {"metric": "fk_jacobian_configs_per_s", "value": value, "unit": "configs/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "fk_jacobian_1M_fetch_configs", "configs_per_gpu": B, "D": D,
                           "layout": "q[B,7] -> pos[B,3], rot[B,9], J[B,6,7] (array-of-structs)",
                           "l2": "working set 256 MB per step > 126 MB L2 (no flush needed)",
                           "parallelism": "dp%d (independent shards, no data-path collective)" % world},
                "roofline": roofline, "e2e": e2e, "gpu_launches": launches, "clocks": clocks}
        if cpu:
            line["cpu_baseline"] = cpu
        if also:
            line["also"] = also
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
