"""Frame-queue sharding across GPUs (one process per GPU).

Frames are independent, so N GPUs = N decoder instances, each working through its own contiguous shard of the
frame queue.  There is NO collective on the data path; the only cross-rank step is the host-side merge of the
per-rank results and iteration statistics (torch.distributed all_gather_object / all_reduce; gloo or nccl).
"""
from __future__ import annotations

import numpy as np


def shard_bounds(n_frames: int, world: int, rank: int):
    """Contiguous balanced shard [lo, hi) of the frame queue for `rank`."""
    base, extra = divmod(n_frames, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def merge_reports(reports, frame_counts):
    """Host-side merge of per-rank reports into the fields decode() reports (test_report.h:20-23)."""
    total = int(sum(frame_counts))
    live = [(r, n) for r, n in zip(reports, frame_counts) if n > 0]
    if not live:
        return {"max_iter": 0, "min_iter": 0xFFFFFFFF, "avg_iter": 0.0, "n_frames": 0}
    return {
        "max_iter": max(int(r["max_iter"]) for r, _ in live),
        "min_iter": min(int(r["min_iter"]) for r, _ in live),
        "avg_iter": float(sum(float(r["avg_iter"]) * n for r, n in live) / total),
        "n_frames": total,
    }


def decode_sharded(decode_fn, inputs: np.ndarray, syndromes: np.ndarray, group=None):
    """Every rank calls this with the FULL host arrays (reference API layout, inputs[:, frame]); each decodes
    its shard with decode_fn(inputs_shard, syndromes_shard) -> (results, iter_counts, report) and the results
    are merged on every rank.  Returns (results[n_frames, N/32], iter_counts[n_frames], merged report)."""
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n_frames = inputs.shape[1]
    lo, hi = shard_bounds(n_frames, world, rank)
    if hi > lo:
        res, iters, rep = decode_fn(np.ascontiguousarray(inputs[:, lo:hi]), np.ascontiguousarray(syndromes[lo:hi]))
    else:
        res, iters, rep = np.zeros((0, 0), dtype=np.uint32), np.zeros(0, dtype=np.uint32), {}
    if world == 1:
        return res, iters, merge_reports([rep], [n_frames])
    gathered = [None] * world
    dist.all_gather_object(gathered, (lo, hi, res, iters, rep), group=group)
    width = max(g[2].shape[1] for g in gathered if g[1] > g[0])
    results = np.zeros((n_frames, width), dtype=np.uint32)
    all_iters = np.zeros(n_frames, dtype=np.uint32)
    for g_lo, g_hi, g_res, g_iters, _ in gathered:
        if g_hi > g_lo:
            results[g_lo:g_hi] = g_res
            all_iters[g_lo:g_hi] = g_iters
    merged = merge_reports([g[4] for g in gathered], [g[1] - g[0] for g in gathered])
    return results, all_iters, merged
