"""World-size-2 CPU tests (gloo) of the multi-GPU host logic: bucketed gradient
all-reduce driven by the backward pass, parameter broadcast, and tile sharding for dense
inference.  The device path is the same code with NCCL tensors (bench.py --gpus N)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from elektronn2_b200 import parallel


class FakeParam(object):
    def __init__(self, shape):
        self.shape = shape


class FakeStore(object):
    """Same layout rules as executor.ParamStore ([weights | biases], 4-float aligned)."""

    def __init__(self, w_sizes, b_sizes, rank):
        self.entries, off = [], 0
        for i, s in enumerate(w_sizes):
            self.entries.append(('w%d' % i, FakeParam((s,)), off, s))
            off += (s + 3) // 4 * 4
        self.n_reg = off
        for i, s in enumerate(b_sizes):
            self.entries.append(('b%d' % i, FakeParam((s,)), off, s))
            off += (s + 3) // 4 * 4
        self.total = off
        self.G = torch.arange(off, dtype=torch.float32) * (rank + 1)
        self.P = torch.full((off,), float(rank))
        self.version = 0


class FakeModel(object):
    data_parallel = None
    _store = None


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, w, _ = parallel.init_from_env(backend='gloo')
    assert (r, w) == (rank, world)
    model = FakeModel()
    dp = parallel.DataParallel(model, bucket_mb=4 * 1000 / (1024.0 * 1024.0))   # ~1000-float buckets
    store = FakeStore([700, 900, 300, 1500, 64], [16, 16, 8, 2], rank)
    model._store = store
    assert dp.grad_scale() == 1.0 / world
    dp.broadcast_parameters(store)
    assert float(store.P.abs().max()) == 0.0                                    # rank 0's values everywhere
    dp.begin_step(store)
    # the layer nearest the input has a bucket of its own, merged with the (contiguous) bias region
    assert dp._buckets[0][0] == 0 and dp._buckets[-1] == (store.entries[4][2], store.total)
    assert all(any(s == e[2] for s, _ in dp._buckets) for e in store.entries[3:5])   # E2_DP_TAIL=2: one bucket per tail layer
    assert all(a[1] == b[0] for a, b in zip(dp._buckets, dp._buckets[1:]))      # contiguous, no gaps
    launched = []
    for _, _, off, size in store.entries:
        if off >= store.n_reg:
            break
        dp.on_gradients_ready(off + size)      # what executor.Plan calls after each wgrad launch
        launched.append(dp._next)
    assert launched[0] <= 1 and launched[-1] == len(dp._buckets) - 1 and sorted(launched) == launched
    dp.finish_step(store)
    expect = torch.arange(store.total, dtype=torch.float32) * sum(range(1, world + 1))
    ok = bool(torch.equal(store.G, expect)) and dp.bytes_reduced == store.total * 4
    # the simple whole-buffer form must not reduce a second time in the same step
    dp.allreduce_gradients(store)
    ok = ok and bool(torch.equal(store.G, expect))
    # second step, the way the fused optimiser step drives it (executor.Plan._train_body_bwd): a split offset S
    # that is NOT a natural bucket end becomes one; once the wgrads below S have been launched, on_gradients_ready(S)
    # + wait_launched() must leave [0, S) fully reduced while the rest is still local
    store.G = torch.arange(store.total, dtype=torch.float32) * (rank + 1)
    S = store.entries[3][2]                    # offset of the 4th weight: inside the second ~1000-float bucket
    dp.begin_step(store, split=S)
    ok = ok and any(e == S for _, e in dp._buckets) and all(a[1] == b[0] for a, b in zip(dp._buckets, dp._buckets[1:]))
    ok = ok and not any(s < S < e for s, e in dp._buckets)
    for _, _, off, size in store.entries[:3]:
        dp.on_gradients_ready(off + size)
    dp.on_gradients_ready(S)
    dp.wait_launched()
    ok = ok and bool(torch.equal(store.G[:S], expect[:S]))
    ok = ok and bool(torch.equal(store.G[S:store.n_reg], (torch.arange(store.total, dtype=torch.float32) * (rank + 1))[S:store.n_reg]))
    dp.finish_step(store)
    ok = ok and bool(torch.equal(store.G, expect))
    q.put((rank, ok))
    dist.destroy_process_group()


def test_bucketed_allreduce_two_ranks_gloo():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_bucket_ranges_cover_buffer():
    s = FakeStore([5, 4096, 100, 7, 9000, 1], [], 0)
    b = parallel.bucket_ranges(s.entries, s.total, 4000)
    assert b[0][0] == 0 and b[-1][1] == s.total
    assert all(x[1] == y[0] for x, y in zip(b, b[1:]))
    offs = {e[2] for e in s.entries} | {s.total}
    assert all(x[0] in offs and x[1] in offs for x in b)       # buckets end on parameter boundaries


def test_tile_sharding_is_a_partition():
    from elektronn2_b200.neuromancer.dense import tile_list, shard_tiles
    tiles = tile_list([7, 5, 3])
    for world in (1, 2, 4, 8):
        parts = [shard_tiles(tiles, r, world) for r in range(world)]
        assert sum(parts, []) == tiles                           # contiguous blocks, reference loop order
