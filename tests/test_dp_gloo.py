"""Host-side logic of the data-parallel step with world_size 2 over gloo on CPU: the per-layer gradient buckets cover the
flat buffer exactly once, in backward order, and reducing them bucket by bucket equals one allreduce of the whole buffer."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from spokennlp_b200.trainer import allreduce_bucket, gradient_buckets
    numel, firsts = 10_000, [1_000, 2_500, 4_000, 7_000]
    buckets = gradient_buckets(firsts, numel)
    g = torch.Generator().manual_seed(100 + rank)
    flat = torch.randn(numel, generator=g)
    whole = flat.clone()
    dist.all_reduce(whole)
    works = [allreduce_bucket(flat, lo, hi) for lo, hi in buckets]      # async, in backward order
    for w in works:
        w.wait()
    ok = torch.equal(flat, whole)
    # averaging is applied later as a multiplier: mean gradient == sum * (1 / world)
    ref_mean = sum(torch.randn(numel, generator=torch.Generator().manual_seed(100 + r)) for r in range(world)) / world
    ok = ok and torch.allclose(flat * (1.0 / world), ref_mean, atol=1e-6)
    # the trainer can send the overlapped buckets through a second communicator and keep the last one on the default
    # group: mixing groups per bucket gives the same sums
    bg = dist.new_group(ranks=list(range(world)))
    flat2 = torch.randn(numel, generator=torch.Generator().manual_seed(100 + rank))
    works = [allreduce_bucket(flat2, lo, hi, group=(bg if i + 1 < len(buckets) else None)) for i, (lo, hi) in enumerate(buckets)]
    for w in works:
        w.wait()
    ok = ok and torch.equal(flat2, whole)
    q.put((rank, ok))
    dist.destroy_process_group()


def test_gradient_buckets_cover_flat_buffer_in_backward_order():
    from spokennlp_b200.trainer import gradient_buckets
    firsts, numel = [100, 300, 700], 1000
    b = gradient_buckets(firsts, numel)
    assert b == [(700, 1000), (300, 700), (100, 300), (0, 100)]       # last layer (+head) first, embeddings last
    covered = sorted(b)
    assert covered[0][0] == 0 and covered[-1][1] == numel
    assert all(covered[i][1] == covered[i + 1][0] for i in range(len(covered) - 1))


def test_bucketed_allreduce_world2_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True), (1, True)]
