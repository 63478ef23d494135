"""The N > 1 path on the CPU (gloo, world_size 2): utterance sharding has no data-path collective, so what runs across
ranks is the deterministic shard assignment and the timing reduction bench.py / an evaluation driver performs."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tss_with_dprnn_b200.sharding import chunk_count, length_buckets, lpt_assign, reduce_timing  # noqa: E402


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    import bench
    per_rank = bench.cfg3_buckets(world, 64)
    mine = per_rank[rank]
    # every rank computes the same global assignment; exchange a checksum of the own shard and the shard sizes
    ids = torch.tensor([sum(t for bk in mine for t, _ in bk), sum(len(bk) for bk in mine),
                        sum(chunk_count(t) for bk in mine for t, _ in bk)], dtype=torch.int64)
    gathered = [torch.zeros_like(ids) for _ in range(world)]
    dist.all_gather(gathered, ids)
    ms, work = reduce_timing(10.0 + rank, float(ids[0]) / 8000, torch.device('cpu'))
    q.put((rank, [g.tolist() for g in gathered], ms, work))
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_sharding_and_timing_reduction():
    world, port = 2, 29655
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=100) for _ in range(world))
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    import bench
    rows = bench.load_test_set_lengths()
    (r0, g0, ms0, w0), (r1, g1, ms1, w1) = res
    assert g0 == g1                                           # both ranks see the same gathered picture
    assert g0[0][0] + g0[1][0] == sum(t for t, _ in rows)     # the shards partition the test set (samples ...
    assert g0[0][1] + g0[1][1] == len(rows) == 3000           # ... and utterances)
    c0, c1 = g0[0][2], g0[1][2]
    assert abs(c0 - c1) / max(c0, c1) < 0.03                  # LPT balance on the masker cost (chunk count)
    assert ms0 == ms1 == 11.0                                 # max over ranks
    assert abs(w0 - sum(t for t, _ in rows) / 8000) < 1e-6 and w0 == w1      # whole-job audio seconds


def test_lpt_and_buckets_properties():
    lengths = [24000 + 37 * i % 9000 for i in range(1000)]
    buckets = length_buckets(lengths, 64)
    assert sorted(i for b in buckets for i in b) == list(range(1000))
    assert all(max(lengths[i] for i in a) <= min(lengths[i] for i in b) for a, b in zip(buckets, buckets[1:]))
    costs = [sum(chunk_count(lengths[i]) for i in b) for b in buckets]
    for world in (1, 2, 4, 8):
        parts = lpt_assign(costs, world)
        assert sorted(i for p in parts for i in p) == list(range(len(buckets)))
        loads = [sum(costs[i] for i in p) for p in parts]
        assert max(loads) - min(loads) <= max(costs)
    assert chunk_count(24000) == 194 and chunk_count(111920) == 898


def _grad_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from tss_with_dprnn_b200.dp import FlatParams, allreduce_mean
    torch.manual_seed(0)                                     # identical replicas
    net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.PReLU(), torch.nn.Linear(16, 4))
    net[1].weight.requires_grad_(False)                      # frozen parameters stay out of the flat buffer
    fp = FlatParams(net)
    for _, p in fp.named:
        p.grad.fill_(float(rank + 1))
    allreduce_mean(fp.grad)
    q.put((rank, fp.numel, min(float(p.grad.min()) for _, p in fp.named), float(fp.grad.max()),
           bool(net[0].weight.data_ptr() == fp.flat.data_ptr())))
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_flat_gradient_allreduce_two_ranks():
    """cfg 5's single exchange step: mean all-reduce of one flat gradient buffer (gloo on the CPU here, NCCL on GPUs)."""
    world, port = 2, 29656
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_grad_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=100) for _ in range(world))
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    for rank, numel, lo, hi, aliased in res:
        assert numel == 8 * 16 + 16 + 16 * 4 + 4             # PReLU weight excluded
        assert lo == hi == 1.5                               # mean of 1 and 2 on both ranks
        assert aliased                                       # parameters are views of the flat buffer


def _bcast_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from tss_with_dprnn_b200.dp import FlatParams, allreduce_mean, average_buffers, broadcast_state
    torch.manual_seed(100 + rank)                            # replicas built from DIFFERENT seeds
    net = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.BatchNorm1d(16), torch.nn.Linear(16, 4))
    net[1].running_mean.fill_(float(rank))
    fp = FlatParams(net)
    before = float(fp.flat.double().sum())
    moments = torch.full_like(fp.flat, float(rank))
    broadcast_state([fp.flat, moments] + list(net.buffers()))
    after = float(fp.flat.double().sum())
    # one data-parallel SGD step on rank-dependent gradients: the replicas must stay identical
    for _, p in fp.named:
        p.grad.fill_(float(rank + 1))
    allreduce_mean(fp.grad)
    fp.flat.add_(fp.grad, alpha=-0.1)
    net[1].running_var.fill_(1.0 + rank)                     # rank-local statistics of the step
    average_buffers([b for b in net.buffers() if b.dtype.is_floating_point])
    q.put((rank, before, after, float(moments.max()), float(net[1].running_mean.max()), float(fp.flat.double().sum()),
           float(net[0].weight.double().sum()), float(net[1].running_var.mean())))
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_replicas_start_identical_from_different_seeds():
    """SpeTrainStep broadcasts rank 0's parameters, Adam moments and BatchNorm buffers at construction (and
    after load_checkpoint); here the plumbing it uses (dp.broadcast_state / average_buffers) on gloo, world size 2."""
    world, port = 2, 29657
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_bcast_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=100) for _ in range(world))
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    (_, b0, a0, m0, rm0, s0, w0, rv0), (_, b1, a1, m1, rm1, s1, w1, rv1) = res
    assert b0 != b1                                           # different seeds really gave different replicas
    assert a0 == a1 == b0                                     # both now hold rank 0's parameters
    assert m0 == m1 == 0.0 and rm0 == rm1 == 0.0              # moments / BatchNorm buffers likewise
    assert s0 == s1 and w0 == w1                              # still identical after an all-reduced update
    assert rv0 == rv1 == 1.5                                  # BatchNorm statistics averaged for the checkpoint
