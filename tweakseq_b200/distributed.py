"""Multi-GPU driver: one process per GPU (torchrun), row slabs gathered to rank 0.

The pair space shards with no exchange during compute: every rank
packs the same database, computes the contiguous slab of sorted-order rows the library's
planner assigns it (tsq_plan_partition / tsq_partition), and one grouped send/recv over NCCL
(NVLink 5 / NVSwitch) moves the slabs into rank 0's score buffer, where tsq_finalize un-sorts
and derives the distances.  PyTorch is plumbing here: process group + a zero-copy view of the
library's device buffer.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import capi


def gather_slabs(buf: torch.Tensor, ranges, group=None, root: int = 0):
    """buf: this rank's full-size packed score buffer with ranges[rank] filled.

    After the call rank `root` holds every rank's slab.  One grouped batch of point-to-point
    ops (ncclGroupStart/End under NCCL), since slab sizes are uneven and there is no gatherv."""
    rank = dist.get_rank(group)
    world = dist.get_world_size(group)
    assert len(ranges) == world
    ops = []
    if rank == root:
        for r in range(world):
            b, e = ranges[r]
            if r != root and e > b:
                ops.append(dist.P2POp(dist.irecv, buf[b:e], r if group is None else dist.get_global_rank(group, r), group))
    else:
        b, e = ranges[rank]
        if e > b:
            ops.append(dist.P2POp(dist.isend, buf[b:e], root if group is None else dist.get_global_rank(group, root), group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return buf


class ShardedRun:
    """All-vs-all over the ranks of a process group.  Rank 0 ends up with scores/distances."""

    def __init__(self, seqs, alphabet=capi.PROTEIN, gap_open=-1, gap_extend=-1, flags=0, group=None,
                 device: int | None = None):
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        dev = torch.cuda.current_device() if device is None else device
        self.ctx = capi.Context(alphabet=alphabet, gap_open=gap_open, gap_extend=gap_extend, device=dev,
                                part_rank=self.rank, part_world=self.world, flags=flags)
        self.ctx.set_sequences(seqs)
        # a non-default stream: its handle is what tsq_set_stream needs (0 would mean "own stream"),
        # and torch.cuda.Event / NCCL ordering then see the same stream the kernels run on
        self.stream = torch.cuda.Stream(device=dev)
        self.ctx.set_stream(self.stream.cuda_stream)
        self.ranges = None
        self.buf = None

    def upload(self):
        self.ctx.upload()
        # every rank plans all ranks' slabs from the same sequences: no exchange of ranges before the gather
        self.ranges = [self.ctx.partition_of(r) for r in range(self.world)]
        assert self.ranges[self.rank] == self.ctx.partition()
        self.buf = torch.as_tensor(self.ctx.device_scores(), device="cuda")

    def compute(self):
        """Kernels of this rank + the gather, all on the current stream (no host sync)."""
        with torch.cuda.stream(self.stream):
            self.ctx.compute()
            if self.world > 1:
                gather_slabs(self.buf, self.ranges, self.group)

    def finish(self):
        """Rank 0: un-sort + distances + D2H.  Others: wait for their stream."""
        if self.rank == 0:
            self.ctx.finalize()
            self.ctx.download()
        else:
            self.ctx.synchronize()

    def close(self):
        self.ctx.close()
