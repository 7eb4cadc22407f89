"""Multi-GPU correctness check (run under torchrun, NCCL): ShardedRun on ragged protein and mixed-regime
nucleotide inputs; rank 0 compares the gathered, finalized matrix with the CPU oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import tweakseq_b200 as t
from tweakseq_b200 import synth
from tweakseq_b200.distributed import ShardedRun
from oracle import pyoracle as o

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rng = np.random.default_rng(5)
AA = "ARNDCQEGHILKMFPSTWYVBZX"
cases = [
    (0, ["".join(rng.choice(list(AA), int(l))) for l in rng.integers(0, 400, 300)] + [""]),
    (1, ["".join(rng.choice(list("ACGT"), int(l))) for l in rng.integers(20, 900, 60)] + synth.nucleotide(4, 8000, 9500, 3)),
    (0, synth.protein(256, 300, 2)),          # fixed length: results stay sharded, every rank downloads its slab
    (0, synth.protein(1001, 120, 7)),
    (1, synth.nucleotide(24, 9000, 9000, 9)),  # fixed length, wavefront kernel
]
for alphabet, seqs in cases:
    run = ShardedRun(seqs, alphabet=alphabet, device=local)
    run.upload(); run.compute(); run.finish()
    torch.cuda.synchronize()
    if rank == 0:
        enc = [o.encode(s, alphabet) for s in seqs]
        mat = o.matrix(alphabet)
        ref, _ = o.all_pairs(enc, mat, 10 if alphabet else 11, 1, nthreads=os.cpu_count())
        selfs = np.array([o.self_score(e, mat) for e in enc], dtype=np.int32)
        assert (run.scores() == ref).all(), "scores differ"
        assert np.asarray(run.distances()).tobytes() == o.distances(ref, selfs).tobytes(), "distances differ"
        print(f"sharded ok: world {world}, alphabet {alphabet}, n {len(seqs)}, results {'sharded over the ranks (own PCIe links)' if run.sharded else 'gathered to rank 0 (NCCL)'}, ranges {run.ranges}", flush=True)
    run.close()
dist.barrier()
dist.destroy_process_group()
