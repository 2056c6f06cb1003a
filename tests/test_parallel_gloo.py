"""world_size-2 gloo test of the multi-rank path: proof sharding + the single accept-bit all-gather."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from zkp_b200 import parallel


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, bad_rank, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    N = 11
    rows = np.arange(3 * N * 32, dtype=np.uint8).reshape(3, N, 32)
    mine = parallel.shard_columns(rows, rank, world)
    lo, hi = parallel.shard_range(N, rank, world)
    assert mine.shape == (3, hi - lo, 32) and (mine == rows[:, lo:hi]).all()
    bits, verdict = parallel.gather_accept_bits(rank != bad_rank)
    # single-verdict mode plumbing: 160-byte partial sums in rank order; a rank with an invalid shard voids the gather
    part = (np.arange(20, dtype=np.uint64) + np.uint64(1000 * rank) + np.uint64(1 << 63)).reshape(4, 5)
    allp = parallel.gather_partial_sums(part)
    assert allp.shape == (world, 20) and allp.dtype == np.uint64
    for r in range(world):
        assert (allp[r] == np.arange(20, dtype=np.uint64) + np.uint64(1000 * r) + np.uint64(1 << 63)).all()
    voided = parallel.gather_partial_sums(None if rank == bad_rank else part)
    assert (voided is None) == (bad_rank >= 0)
    q.put((rank, lo, hi, bits, verdict))
    dist.destroy_process_group()


def _run(bad_rank):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, bad_rank, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    return res


def test_shards_cover_and_bits_gather():
    res = _run(bad_rank=-1)
    assert [r[1:3] for r in res] == [(0, 6), (6, 11)]
    assert all(r[3] == [1, 1] and r[4] for r in res)


def test_one_rejecting_shard_fails_the_batch():
    res = _run(bad_rank=1)
    assert all(r[3] == [1, 0] and not r[4] for r in res)


def test_shard_range_properties():
    for n in (0, 1, 7, 8, 2**21 + 3):
        for w in (1, 2, 3, 8):
            spans = [parallel.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
