#!/usr/bin/env python
"""bench.py -- benchmark of the rlitg hot path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload all|fk|torm|pirl]

BASELINE.json's metric has two halves; both are in the line the driver keeps:
  headline    batched FK + geometric Jacobian over 1,048,576 random Fetch configurations per GPU (configs[1]; D = 7, FP32,
              q [B,7] -> pos [B,3], rot [B,9], J [B,6,7]).  A "step" is one pass over that batch.  `value` = configs/s
              over all ranks with inputs resident in HBM; `e2e` = the same through the host-buffer C-ABI entry point
              (pinned host buffers, H2D + D2H inside the timed region).
  torm, pirl  TORM cost + gradient for 1,024 x 200-waypoint trajectories against a 128^3 grid (configs[3]) and the PIRL env
              step over 16,384 envs (configs[2]): each with its own value / roofline / cpu_baseline / e2e, nested under
              those keys of the line (`roofline.torm`, `cpu_baseline.pirl`, ...) and repeated in a compact `summary` at
              the very end of the line.
  config5     the strong split of configs[4] (8M configs / 65,536 envs / 8,192 trajectories over the ranks).
One JSON line on rank 0.
"""
import argparse
import hashlib
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
CPU_SAMPLE = 1 << 18                     # configurations per step of the CPU arm: a bounded sample of the 1M-config workload


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


def cpu_sample_q():
    import numpy as np
    rng = np.random.default_rng(0)
    ll = np.array([-1.6056, -1.221, -3.141592, -2.251, -3.141592, -2.16, -3.141592]); ul = -ll.copy(); ul[1] = 1.518
    return rng.uniform(ll, ul, size=(CPU_SAMPLE, D))


def cpu_reference_fk_jac(seconds_target=8.0):
    """times the CPU oracle (FP64 restatement of the reference's two KDL calls per configuration) on a bounded sample"""
    from oracle import oracle as O
    # all host threads this process may use (torchrun exports OMP_NUM_THREADS=1: override it explicitly)
    O.set_num_threads(len(os.sched_getaffinity(0)))
    q = cpu_sample_q()
    O.fk_jac(q[:1024])
    t0 = time.perf_counter(); O.fk_jac(q); t1 = time.perf_counter()
    reps = min(max(1, int(seconds_target / max(t1 - t0, 1e-6))), 2048)
    t0 = time.perf_counter()
    for _ in range(reps):
        O.fk_jac(q)
    dt = time.perf_counter() - t0
    return {"value": CPU_SAMPLE * reps / dt, "unit": "configs/s", "cores": O.num_threads(), "kind": "port",
            "sample": "%d x %d of the 1,048,576 configs per step; FK + Jacobian as the reference's two chain walks per "
                      "configuration (JntToCart + JntToJac), FP64 oracle, OpenMP, %.1f s" % (reps, CPU_SAMPLE, dt)}


def run_reference(args):
    """--impl reference: the reference's CPU path for the headline workload (the FP64 oracle port: the reference itself
    needs ROS / KDL and cannot be built here), every host thread, a bounded sample of the workload per step."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    steps, warm = args.steps, args.warmup
    from oracle import oracle as O
    O.set_num_threads(len(os.sched_getaffinity(0)))
    cores = O.num_threads()
    q = cpu_sample_q()
    for _ in range(warm):
        O.fk_jac(q)
    t0 = time.perf_counter()
    for _ in range(steps):
        O.fk_jac(q)
    dt = time.perf_counter() - t0
    v = CPU_SAMPLE * steps / dt
    sample = ("%d of the 1,048,576 configs per step (throughput metric: the sample bounds the run time, the per-config work is "
              "identical); two chain walks per configuration like the reference's JntToCart + JntToJac" % CPU_SAMPLE)
    line = {"impl": "reference", "metric": "fk_jacobian_configs_per_s", "value": v, "unit": "configs/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": 1e3 * dt / steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "fk_jacobian_1M_fetch_configs", "D": D, "sample_per_step": CPU_SAMPLE,
                       "note": "reference CPU path = FP64 oracle port of the KDL calls (reference not buildable here)"},
            "cpu_baseline": {"value": v, "unit": "configs/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "configs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def library_provenance():
    from rlitg_b200 import _lib
    st = os.stat(_lib.LIB_PATH)
    h = hashlib.sha256()
    with open(_lib.LIB_PATH, "rb") as f:
        h.update(f.read())
    return {"path": os.path.relpath(_lib.LIB_PATH, ROOT), "built_in_this_process": bool(_lib.BUILT_IN_THIS_PROCESS),
            "shipped_prebuilt": not _lib.BUILT_IN_THIS_PROCESS, "mtime_utc": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime(st.st_mtime)),
            "sha16": h.hexdigest()[:16]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="all", choices=["all", "fk", "torm", "pirl"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="headline workload only")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from rlitg_b200 import kinematics as K, _lib
    import bench_also as BA

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

    hbm_peak, peak_src = measured_peaks()
    want = lambda w: args.workload in ("all", w) and not (args.no_also and w != "fk")
    cpu_ok = rank == 0 and world == 1 and not args.no_cpu_baseline

    # ------------------------------------------------------------------------------------------ headline: FK + Jacobian
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
    # EXACTLY K steps per timed block; 20 steps are ~0.9 ms of GPU time and move by ~5 % from block to block, so the block is
    # repeated and the median block reported (every block: barrier + synchronize on both sides, CUDA events, max over ranks)
    BLOCKS = 7
    blocks, launches = [], 0
    for _ in range(BLOCKS):
        n0 = _lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        barrier()
        launches = _lib.launch_count() - n0
        blocks.append((max_over_ranks(e0.elapsed_time(e1)), e0.elapsed_time(e1)))
    blocks.sort()
    ms, ms_local = blocks[len(blocks) // 2]
    # keep the GPU busy a little longer so that NVML gets samples under load on very short runs
    t_end = time.time() + 0.3
    while time.time() < t_end:
        step()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ms_per_step = ms / args.steps
    value = world * B * args.steps / (ms * 1e-3)
    achieved = BYTES_PER_CONFIG * B / (ms_local / args.steps * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "traffic": BA.profile_traffic("fk_jac_kernel"),
                "traffic_source": "profiles/traffic.json: dram__bytes of one ncu --set full capture of this kernel (not this run)",
                "kernel": "fk_jac_persist_kernel<float,7,aos,2 stages>", "bytes_per_launch": BYTES_PER_CONFIG * B,
                "peak_source": peak_src}

    # ---- end to end: host buffers through rlitg_fk_jac_host (H2D q, kernel, D2H pos/rot/J)
    from rlitg_b200.dist import bind_host_to_gpu
    prev_affinity = None if os.environ.get("RLITG_BENCH_NO_NUMA") else bind_host_to_gpu(local)
    hq = q_host.pin_memory()
    hpos = torch.empty(B, 3).pin_memory(); hrot = torch.empty(B, 3, 3).pin_memory(); hjac = torch.empty(B, 6, D).pin_memory()
    e2e_steps = max(5, min(args.steps, 20))
    for _ in range(2):
        K.fk_jac_host(hq, pos=hpos, rot=hrot, jac=hjac)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        K.fk_jac_host(hq, pos=hpos, rot=hrot, jac=hjac)
    torch.cuda.synchronize()
    dt = max_over_ranks(time.perf_counter() - t0)
    h2d, d2h = B * D * 4, B * (3 + 9 + 6 * D) * 4
    # the ceiling of that number: the same bytes as plain pinned copies into the SAME host buffers (both directions at once,
    # every rank at once), best of three blocks in each of two forms: one copy per output, and the pipe's own chunking
    dq_ = torch.empty(B, D, device=dev)
    dsrc = [torch.empty(B, 3, device=dev), torch.empty(B, 3, 3, device=dev), torch.empty(B, 6, D, device=dev)]
    hdst = [hpos, hrot, hjac]
    s_in, s_out = torch.cuda.Stream(device=dev), [torch.cuda.Stream(device=dev) for _ in range(3)]

    def copies(nchunks):
        with torch.cuda.stream(s_in):
            dq_.copy_(hq, non_blocking=True)
        for k in range(nchunks):
            with torch.cuda.stream(s_out[k % 3]):
                for h_, d_ in zip(hdst, dsrc):
                    h_.chunk(nchunks)[k].copy_(d_.chunk(nchunks)[k], non_blocking=True)

    dtc = None
    for nchunks in (1, 8):
        for _ in range(3):
            copies(nchunks)
            torch.cuda.synchronize()
            barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                copies(nchunks)
            torch.cuda.synchronize()
            d_ = max_over_ranks(time.perf_counter() - t0)
            dtc = d_ if dtc is None else min(dtc, d_)
    e2e = {"value": world * B * e2e_steps / dt, "unit": "configs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "steps": e2e_steps, "api": "rlitg_fk_jac_host (pinned host buffers)", "numa_bound": prev_affinity is not None,
           "pcie_ceiling_configs_per_s": world * B * e2e_steps / dtc,
           "pcie_ceiling_note": "the step's %d + %d bytes as bare pinned copies into the same host buffers, both directions at once, "
                                "all %d rank(s) at once (best of 3 blocks, whole and in 8 chunks over 3 streams): what bounds this "
                                "number is the host side of the box (one PCIe root / NUMA domain behind all GPUs)"
                                % (h2d, d2h, world)}
    del dq_, dsrc
    if prev_affinity is not None:
        os.sched_setaffinity(0, prev_affinity)

    cpu_baseline = cpu_reference_fk_jac() if cpu_ok else None
    config = {"workload": "fk_jacobian_1M_fetch_configs", "configs_per_gpu": B, "D": D,
              "layout": "q[B,7] -> pos[B,3], rot[B,9], J[B,6,7] (array-of-structs)",
              "l2": "working set 256 MB per step > 126 MB L2 (no flush needed)",
              "timing": "median of %d blocks of exactly %d steps (CUDA events, barrier + synchronize around each block, max over ranks)"
                        % (BLOCKS, args.steps),
              "parallelism": "dp%d (independent shards, no data-path collective)" % world}

    # -------------------------------------------------------------------------------------- the other half of the metric
    also, summary = {}, {}
    sec_steps = max(10, min(args.steps, 40))
    if want("torm") or want("pirl"):
        fma_peak, dfma_peak = BA.pipe_peaks(torch, dev)
        also["pipe_peaks_tflops"] = {"fp32_fma": fma_peak, "fp64_fma": dfma_peak}

    def nest(name, rec):
        """a workload's compact record under the keys the driver keeps + the summary at the end of the line"""
        roofline[name] = rec.get("roofline")
        e2e[name] = rec.get("e2e")
        if cpu_baseline is not None:
            cpu_baseline[name] = rec.get("cpu_baseline")
        rf, cb, ee = rec.get("roofline") or {}, rec.get("cpu_baseline") or {}, rec.get("e2e") or {}
        summary[name] = {"value": rec["value"], "unit": rec["unit"],
                         "roofline": {k: rf.get(k) for k in ("bound", "achieved", "peak", "unit", "frac", "traffic")},
                         "cpu_baseline": {k: cb.get(k) for k in ("value", "cores", "kind")} if cb else None,
                         "e2e": {k: ee.get(k) for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")}}

    if want("torm"):
        try:
            rec, det = BA.run_torm(torch, dev, rank, world, max_over_ranks, barrier, sec_steps, 5, fma_peak, cpu=cpu_ok)
            nest("torm", rec)
            summary["torm"]["ms_per_iteration"] = rec["ms_per_iteration"]
            config["torm"] = det.pop("config")
            also["torm_cost_grad"] = dict(det, ms_per_iteration=rec["ms_per_iteration"])
        except Exception as e:  # noqa: BLE001
            also["torm_cost_grad"] = {"error": repr(e)}
    if want("pirl"):
        try:
            rec, det = BA.run_pirl(torch, dev, rank, world, max_over_ranks, barrier, sec_steps, 5, hbm_peak, cpu=cpu_ok)
            nest("pirl", rec)
            summary["pirl"]["value_launch"] = rec["value_launch"]
            summary["pirl"]["us_per_step"] = rec["us_per_step"]
            summary["pirl"]["us_per_step_cuda_graph"] = rec["us_per_step_cuda_graph"]
            summary["pirl"]["us_per_step_50_steps_per_graph"] = rec.get("us_per_step_50_steps_per_graph")
            config["pirl"] = det.pop("config")
            also["pirl_env_step"] = det
        except Exception as e:  # noqa: BLE001
            also["pirl_env_step"] = {"error": repr(e)}
    if args.workload == "all" and not args.no_also:
        try:
            also["config5_strong_split"] = BA.run_config5(torch, dev, rank, world, max_over_ranks, barrier, max(5, sec_steps // 2), 3)
            c5 = also["config5_strong_split"]
            summary["config5"] = {"fk_configs_per_s": c5["fk_jacobian_8M"]["configs_per_s"],
                                  "pirl_env_steps_per_s": c5["pirl_65536"]["env_steps_per_s"],
                                  "pirl_env_steps_per_s_python_launch": c5["pirl_65536"]["env_steps_per_s_python_launch"],
                                  "torm_waypoints_per_s": c5["torm_8192"]["waypoints_per_s"], "ranks": world}
        except Exception as e:  # noqa: BLE001
            also["config5_strong_split"] = {"error": repr(e)}
        try:
            also["fk_other_layouts"] = BA.run_fk_layouts(torch, dev, rank, world, max_over_ranks, barrier, sec_steps, 5)
        except Exception as e:  # noqa: BLE001
            also["fk_other_layouts"] = {"error": repr(e)}
        if cpu_ok:
            try:
                also["cpu_reference_batch1"] = BA.cpu_batch1()
            except Exception as e:  # noqa: BLE001
                also["cpu_reference_batch1"] = {"error": repr(e)}

    if rank == 0:
        line = {"metric": "fk_jacobian_configs_per_s", "value": value, "unit": "configs/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "gpu_launches": launches,
                "clocks": clocks, "config": config, "library": library_provenance()}
        if also:
            line["also"] = also
        line["roofline"] = roofline
        if cpu_baseline:
            line["cpu_baseline"] = cpu_baseline
        line["e2e"] = e2e
        if args.workload in ("torm", "pirl") and args.workload in summary:
            # --workload torm | pirl: that workload's own record as the headline of the line
            s = summary[args.workload]
            line.update({"metric": {"torm": "torm_cost_grad_waypoints_per_s", "pirl": "pirl_env_steps_per_s"}[args.workload],
                         "value": s["value"], "unit": s["unit"],
                         "ms_per_step": s["ms_per_iteration"] if args.workload == "torm"
                         else 1e-3 * (s.get("us_per_step_50_steps_per_graph") or s["us_per_step"])})
            line["config"] = dict(config.get(args.workload) or {}, headline_of="--workload " + args.workload,
                                  fk_jacobian=dict((k, v) for k, v in config.items() if k not in ("torm", "pirl")))
        if summary:
            line["summary"] = summary  # last: survives a tail-truncated log
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
