"""world_size-2 (and 3) CPU test of the N>1 host logic: the library's row planner
(tsq_plan_partition, host only) + the slab gather (gloo), with the oracle filling each rank's
slab.  Rank 0 must end up with exactly the single-process matrix."""
import os
import socket
import sys
import time

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, lens_list, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import pyoracle as o
    from tweakseq_b200 import capi
    from tweakseq_b200.distributed import gather_slabs
    rng = np.random.default_rng(123)                      # same sequences on every rank
    enc = [rng.integers(0, 20, l).astype(np.uint8) for l in lens_list]
    n = len(enc)
    ranges = capi.plan_partition(lens_list, world)
    b, e = ranges[rank]
    buf = torch.full((n * (n - 1) // 2,), -999, dtype=torch.int32)
    if e > b:
        part, _ = o.all_pairs(enc, o.matrix(0), 11, 1, nthreads=1, pair_begin=b, pair_end=e)
        buf[b:e] = torch.from_numpy(part)
    gather_slabs(buf, ranges)
    if rank == 0:
        full, _ = o.all_pairs(enc, o.matrix(0), 11, 1, nthreads=2)
        q.put((bool((buf.numpy() == full).all()), ranges))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_slab_gather_reassembles_the_matrix(world):
    lens = sorted([int(x) for x in np.random.default_rng(5).integers(20, 60, 41)])  # sorted => identity order
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, lens, q)) for r in range(world)]
    for p in procs:
        p.start()
    ok, ranges = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok
    n = len(lens)
    assert ranges[0][0] == 0 and ranges[-1][1] == n * (n - 1) // 2
    for a, b in zip(ranges[:-1], ranges[1:]):
        assert a[1] == b[0]


def test_planner_balances_cells_and_cuts_on_pair_boundaries():
    from tweakseq_b200 import capi
    n = 1000
    for world in (2, 4, 8):
        r = capi.plan_partition([300] * n, world)
        sizes = [e - b for b, e in r]
        assert sum(sizes) == n * (n - 1) // 2
        assert max(sizes) / (sum(sizes) / world) < 1.02
        # row starts are even (rows travel in pairs through the packed kernel)
        for b, _ in r[1:]:
            rows = [i for i in range(n) if i * n - i * (i + 1) // 2 == b]
            assert rows and rows[0] % 2 == 0
    # ragged lengths: balance is by DP cells, not by pair count
    lens = list(range(1, 401))
    r = capi.plan_partition(lens, 4)
    L = np.array(lens, dtype=np.float64)
    suffix = np.concatenate([np.cumsum(L[::-1])[::-1][1:], [0]])
    rowcost = L * suffix
    starts = [next(i for i in range(len(lens) + 1) if (i * 400 - i * (i + 1) // 2 if i < 399 else 400 * 399 // 2) >= b) for b, _ in r] + [400]
    cells = [rowcost[a:b].sum() for a, b in zip(starts[:-1], starts[1:])]
    assert max(cells) / (sum(cells) / 4) < 1.05


def _shared_worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tweakseq_b200 import capi
    from tweakseq_b200.distributed import SharedResult
    count = n * (n - 1) // 2
    ranges = capi.plan_partition([300] * n, world)            # fixed length: the slabs tile the FINAL triangle
    res = SharedResult(count, want_dist=True)                 # rank 0 creates the segment, everyone maps it
    b, e = ranges[rank]
    ok = True
    for epoch in (1, 2):                                      # two jobs through the same segment
        if rank == world - 1:
            time.sleep(0.2)                                   # a late rank: rank 0 must wait for its mark
        res.scores[b:e] = np.arange(b, e, dtype=np.int32) * epoch   # what tsq_download does with each rank's own slab
        res.distances[b:e] = np.arange(b, e, dtype=np.float64) * 0.5 * epoch
        res.arrive(rank, epoch)                               # ShardedRun.finish: a mark in the segment, no collective
        if rank == 0:
            res.wait_all(epoch)
            ok = ok and bool((res.scores == np.arange(count, dtype=np.int32) * epoch).all() and
                             (res.distances == np.arange(count, dtype=np.float64) * 0.5 * epoch).all())
        dist.barrier()                                        # (the caller's own step boundary)
    if rank == 0:
        with pytest.raises(RuntimeError):
            res.wait_all(3, timeout_s=0.05)                   # nobody arrives at job 3: an error, not a hang
        q.put((ok, res.path, ranges))
    dist.barrier()
    res.close()                                               # rank 0 removes the file
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_shared_result_segment_receives_every_ranks_slab(world):
    """The sharded result path of tweakseq_b200/distributed.py without a GPU: one host segment (a file in /dev/shm or
    /tmp), created by rank 0 and mapped by all ranks; each rank writes only the slab the library's planner gives it;
    after the barrier rank 0 holds the whole matrix; the segment is removed on close."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_shared_worker, args=(r, world, port, 257, q)) for r in range(world)]
    for p in procs:
        p.start()
    ok, path, ranges = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok and not os.path.exists(path)
    assert ranges[0][0] == 0 and ranges[-1][1] == 257 * 256 // 2
