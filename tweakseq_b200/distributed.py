"""Multi-GPU driver, one process per GPU (torchrun).

The pair space shards with no exchange during compute: every rank packs the
same database and computes the contiguous slab of sorted-order rows the library's planner assigns it
(tsq_plan_partition / tsq_partition).  How the slabs come together depends on the input
(tsq_results_sharded):

* fixed-length input (every BASELINE configuration with a large result): the length sort is the
  identity, so a rank's slab IS a contiguous piece of the final packed triangle.  Every rank finalizes
  its own slab (distances next to the scores) and copies it over ITS OWN PCIe link straight into one
  host result that all ranks map -- a POSIX shared-memory segment handed to the library with
  tsq_set_result_buffers.  No gather, N links instead of one.
* ragged input: the un-sort scatters a slab over the triangle, so the slabs travel to rank 0 in one
  grouped send/recv over NCCL (NVLink 5 / NVSwitch) and rank 0 runs tsq_finalize on the whole.

One process driving several devices does the same behind the C ABI (tsq_params.n_devices); this module
is the torchrun flavour.  PyTorch is plumbing here: the process group and a zero-copy view of the
library's device buffer.
"""
from __future__ import annotations

import os
import time
import uuid

import numpy as np
import torch
import torch.distributed as dist

from . import capi


def gather_slabs(buf: torch.Tensor, ranges, group=None, root: int = 0, first: int = 0):
    """buf: this rank's packed score buffer -- the whole triangle on `root`, elsewhere at least the
    rank's own slab, buf[0] being packed index `first`.

    After the call rank `root` holds every rank's slab.  One grouped batch of point-to-point
    ops (ncclGroupStart/End under NCCL), since slab sizes are uneven and there is no gatherv."""
    rank = dist.get_rank(group)
    world = dist.get_world_size(group)
    assert len(ranges) == world
    ops = []
    if rank == root:
        assert first == 0
        for r in range(world):
            b, e = ranges[r]
            if r != root and e > b:
                ops.append(dist.P2POp(dist.irecv, buf[b:e], r if group is None else dist.get_global_rank(group, r), group))
    else:
        b, e = ranges[rank]
        if e > b:
            ops.append(dist.P2POp(dist.isend, buf[b - first:e - first], root if group is None else dist.get_global_rank(group, root), group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return buf


class SharedResult:
    """One host result for all ranks: a file in /dev/shm that every rank maps (rank 0 creates and
    removes it).  scores: int32[count]; distances: float64[count] or None."""

    def __init__(self, count: int, want_dist: bool, group=None):
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        # layout: scores | distances | one page of arrival flags (int64 per rank), each part page-aligned
        off_d = (count * 4 + 4095) & ~4095
        off_f = (off_d + (count * 8 if want_dist else 0) + 4095) & ~4095
        size = off_f + max(4096, (world * 8 + 4095) & ~4095)
        name = [None]
        if rank == 0:
            # /dev/shm when it has the room (a container's default is small: a page touched beyond it is a SIGBUS,
            # not an error code), else a temporary directory on disk: page-locked and mapped the same way
            where = None
            for d in ("/dev/shm", os.environ.get("TMPDIR", "/tmp"), "/tmp"):
                try:
                    st = os.statvfs(d)
                    if st.f_bavail * st.f_frsize > size + (256 << 20):
                        where = d
                        break
                except OSError:
                    pass
            if where is None:
                raise RuntimeError(f"no room for a shared result of {size} bytes in /dev/shm or /tmp")
            name = [f"{where}/tsq_b200_{os.getpid()}_{uuid.uuid4().hex[:12]}"]
            with open(name[0], "wb") as f:
                f.truncate(size)
        if dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.broadcast_object_list(name, src=0 if group is None else dist.get_global_rank(group, 0), group=group)
        self.path, self.owner = name[0], rank == 0
        self._map = np.memmap(self.path, dtype=np.uint8, mode="r+", shape=(size,))
        self.scores = self._map[:count * 4].view(np.int32)
        self.distances = self._map[off_d:off_d + count * 8].view(np.float64) if want_dist else None
        self.flags = self._map[off_f:off_f + world * 8].view(np.int64)     # zero-filled by the truncate
        self.count = count

    def arrive(self, rank: int, epoch: int):
        """This rank's slab of job `epoch` has landed (called after its download returned)."""
        self.flags[rank] = epoch

    def wait_all(self, epoch: int, timeout_s: float = 600.0):
        """Rank 0: until every rank has arrived at `epoch`.  A flag in the mapped segment instead of a collective:
        the slabs already meet in host memory, and an NCCL barrier costs more than the whole copy of a small job."""
        t_end = time.monotonic() + timeout_s
        spins = 0
        while int(self.flags.min()) < epoch:
            spins += 1
            if (spins & 0x3ff) == 0 and time.monotonic() > t_end:
                raise RuntimeError(f"shared result: ranks {np.nonzero(self.flags < epoch)[0].tolist()} never arrived at job {epoch}")

    def close(self):
        self.scores = self.distances = self.flags = None
        self._map = None
        if self.owner and self.path and os.path.exists(self.path):
            os.unlink(self.path)
        self.path = None


class ShardedRun:
    """All-vs-all over the ranks of a process group.  Rank 0 ends up with scores/distances."""

    def __init__(self, seqs, alphabet=capi.PROTEIN, gap_open=-1, gap_extend=-1, flags=0, group=None,
                 device: int | None = None):
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        dev = torch.cuda.current_device() if device is None else device
        self.flags = flags
        self.ctx = capi.Context(alphabet=alphabet, gap_open=gap_open, gap_extend=gap_extend, device=dev,
                                part_rank=self.rank, part_world=self.world, flags=flags)
        self.ctx.set_sequences(seqs)
        # a non-default stream: its handle is what tsq_set_stream needs (0 would mean "own stream"),
        # and torch.cuda.Event / NCCL ordering then see the same stream the kernels run on
        self.stream = torch.cuda.Stream(device=dev)
        self.ctx.set_stream(self.stream.cuda_stream)
        self.ranges = None
        self.buf = None
        self.first = 0
        self.sharded = False
        self.shared: SharedResult | None = None
        self.epoch = 0                        # jobs finished through the current shared result

    def upload(self):
        self.ctx.upload()
        # every rank plans all ranks' slabs from the same sequences: no exchange of ranges before the gather
        self.ranges = [self.ctx.partition_of(r) for r in range(self.world)]
        assert self.ranges[self.rank] == self.ctx.partition()
        self.sharded = self.world > 1 and self.ctx.results_sharded()
        if self.sharded:
            count = self.ctx.npairs
            if self.shared is None or self.shared.count != count:
                if self.shared is not None:
                    self.ctx.set_result_buffers(None)
                    self.shared.close()
                self.shared = SharedResult(count, not (self.flags & capi.FLAG_NO_DISTANCES), self.group)
                self.epoch = 0
                self.ctx.set_result_buffers(self.shared.scores, self.shared.distances)
        else:
            arr, self.first = self.ctx.device_slab()
            self.buf = torch.as_tensor(arr, device="cuda")

    def compute(self):
        """Kernels of this rank (+ the gather where the input needs one), all on one stream (no host sync)."""
        with torch.cuda.stream(self.stream):
            self.ctx.compute()
            if self.world > 1 and not self.sharded:
                gather_slabs(self.buf, self.ranges, self.group, first=self.first)

    def finish(self):
        """Results to the host.  Sharded: every rank finalizes its slab and copies it over its own PCIe
        link into the shared result and marks its arrival there; rank 0 waits for all the marks (the others are free).  Gathered: rank 0 un-sorts, derives the distances and
        downloads; the others wait for their stream."""
        if self.sharded:
            self.ctx.download()
            self.epoch += 1                   # rank 0 may read once every rank's copy has landed
            self.shared.arrive(self.rank, self.epoch)
            if self.rank == 0:
                self.shared.wait_all(self.epoch)
        elif self.rank == 0:
            self.ctx.finalize()
            self.ctx.download()
        else:
            self.ctx.synchronize()

    def scores(self) -> np.ndarray:
        """Rank 0, after finish(): the packed int32 matrix (a view of the shared result when sharded)."""
        return self.shared.scores if self.sharded else self.ctx.scores()

    def distances(self) -> np.ndarray:
        return self.shared.distances if self.sharded else self.ctx.distances()

    def close(self):
        self.ctx.close()
        if self.shared is not None:
            self.shared.close()
            self.shared = None
