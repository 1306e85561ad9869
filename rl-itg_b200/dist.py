"""Multi-GPU plumbing: one process per GPU, shards of the batch axis, statistics over NCCL (or gloo on CPU).

The hot path shards trivially: configurations, environments and trajectories are independent
units, the only coupling (7-point stencils, the smoothness band) is inside one trajectory, which never leaves its
CTA.  So there is NO collective on the kernel path.  What crosses ranks is a few dozen bytes per step: per-shard
cost / reward statistics, and -- for candidate scoring (TormOptimizer::getNewTrajectory,
torm/src/torm_optimizer.cpp:973-1003) -- the arg-min candidate and the winning segment."""
from __future__ import annotations

import torch
import torch.distributed as dist

STAT_FIELDS = ("count", "sum_cost", "min_cost", "argmin", "sum_reward", "n_done", "n_collision")


def world():
    return (dist.get_rank(), dist.get_world_size()) if dist.is_available() and dist.is_initialized() else (0, 1)


def shard_range(n: int, rank: int | None = None, world_size: int | None = None):
    """Contiguous split of range(n): the first n % world ranks get one extra unit.  -> (start, stop)."""
    r, w = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    base, extra = divmod(n, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def local_stats(cost=None, reward=None, done=None, collision=None, offset=0):
    """Per-shard statistics vector (float64 [7], on the tensors' device): STAT_FIELDS."""
    ref = cost if cost is not None else reward
    s = torch.zeros(len(STAT_FIELDS), dtype=torch.float64, device=ref.device)
    s[0] = ref.numel()
    s[2] = float("inf")
    s[3] = -1
    if cost is not None and cost.numel():
        c = cost.double()
        s[1] = c.sum()
        m, i = c.min(dim=0)
        s[2], s[3] = m, i.double() + offset
    if reward is not None:
        s[4] = reward.double().sum()
    if done is not None:
        s[5] = done.double().sum()
    if collision is not None:
        s[6] = collision.double().sum()
    return s


def gather_stats(s: torch.Tensor):
    """all_gather of the per-shard vectors -> dict of global statistics (+ 'per_rank' [W,7])."""
    rank, w = world()
    if w == 1:
        allv = s[None]
    else:
        buf = [torch.empty_like(s) for _ in range(w)]
        dist.all_gather(buf, s)
        allv = torch.stack(buf)
    best = int(torch.argmin(allv[:, 2]).item())
    cnt = allv[:, 0].sum().item()
    return {"count": int(cnt), "sum_cost": allv[:, 1].sum().item(), "mean_cost": allv[:, 1].sum().item() / max(cnt, 1),
            "min_cost": allv[best, 2].item(), "argmin": int(allv[best, 3].item()), "argmin_rank": best,
            "sum_reward": allv[:, 4].sum().item(), "mean_reward": allv[:, 4].sum().item() / max(cnt, 1),
            "n_done": int(allv[:, 5].sum().item()), "n_collision": int(allv[:, 6].sum().item()), "per_rank": allv}


def broadcast_best(rows: torch.Tensor, cost: torch.Tensor, offset: int):
    """Candidate selection across shards: every rank holds candidates `rows` [n_local, ...] with `cost` [n_local]
    (global ids offset .. offset + n_local).  Returns (best cost, global id, winning row) on every rank."""
    st = gather_stats(local_stats(cost=cost, offset=offset))
    rank, w = world()
    owner, gid = st["argmin_rank"], st["argmin"]
    win = rows[gid - offset].clone() if rank == owner else torch.empty_like(rows[0])
    if w > 1:
        dist.broadcast(win, src=owner)
    return st["min_cost"], gid, win


def bind_host_to_gpu(device_index: int):
    """Pins the calling thread to the CPUs NVML reports as local to GPU `device_index` (same NUMA node / PCIe root),
    so that pinned host buffers allocated afterwards land in memory next to that GPU -- with 8 ranks staging
    256 MB per step each, remote-node buffers halve the host<->device rate.  Returns the previous affinity set (pass it to
    os.sched_setaffinity(0, prev) to undo) or None when NVML, the GPU's CPU set or the cpuset cgroup do not allow it."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        uuid = getattr(torch.cuda.get_device_properties(device_index), "uuid", None)
        h = None
        if uuid is not None:
            try:
                h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(uuid)).encode())
            except Exception:  # noqa: BLE001
                h = None
        if h is None:
            h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        prev = os.sched_getaffinity(0)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        local = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
        want = local & prev
        if not want or want == prev:
            return None
        os.sched_setaffinity(0, want)
        return prev
    except Exception:  # noqa: BLE001
        return None
