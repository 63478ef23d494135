"""Utterance sharding across the GPUs of one node: utterances are independent in inference, so
each rank takes its own utterances and there is NO data-path collective - the only cross-rank traffic is the timing /
bookkeeping reduction of a benchmark or evaluation run (a few scalars over torch.distributed)."""
from __future__ import annotations

import torch


def chunk_count(T: int, K: int = 250, P: int = 125, ksz: int = 2, stride: int = 1) -> int:
    """S of DPRNN._segmentation for a T-sample utterance (dprnn.py:189-201): the cost unit of the masker."""
    L = (T - ksz) // stride + 1
    return (L + K) // P + 1


def length_buckets(lengths, bucket: int):
    """Length-sorted buckets of `bucket` utterances (indices into `lengths`): padding-free ragged batches of similar
    cost per utterance (cfg 3)."""
    order = sorted(range(len(lengths)), key=lambda i: lengths[i])
    return [order[i:i + bucket] for i in range(0, len(order), bucket)]


def lpt_assign(costs, world: int):
    """Longest-processing-time-first assignment of items with the given costs to `world` ranks.
    Returns per rank the list of item indices (in assignment order); deterministic, identical on every rank."""
    load, per_rank = [0] * world, [[] for _ in range(world)]
    for i in sorted(range(len(costs)), key=lambda i: (-costs[i], i)):
        r = load.index(min(load))
        per_rank[r].append(i)
        load[r] += costs[i]
    return per_rank


def reduce_timing(ms_local: float, work_local: float, device, group=None):
    """(max over ranks of the device time, sum over ranks of the work): the whole-job rate is sum / max."""
    import torch.distributed as dist
    t = torch.tensor([ms_local], dtype=torch.float64, device=device)
    w = torch.tensor([work_local], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        dist.all_reduce(w, op=dist.ReduceOp.SUM, group=group)
    return float(t.item()), float(w.item())
