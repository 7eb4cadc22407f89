"""configs[1] (1000 unrelated x 300 aa): tsq_msa three times in one process (wall time per call)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tweakseq_b200 as t
from tweakseq_b200 import synth
seqs = synth.config(2)[1]
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for rep in range(reps):
    with t.Context() as ctx:
        ctx.set_sequences(seqs); ctx.run(); ctx.guide_tree()
        t0 = time.perf_counter()
        rows, order = ctx.msa()
        print(f"rep {rep}: msa wall {1e3 * (time.perf_counter() - t0):.1f} ms, msa_ms {ctx.stats()['msa_ms']:.1f}, cols {len(rows[0])}", flush=True)
