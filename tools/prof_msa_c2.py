"""tsq_msa on configs[1] with the per-launch trace (TSQ_MSA_DEBUG=2), twice (the second run has warm caches)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tweakseq_b200 as t
from tweakseq_b200 import synth
seqs = synth.config(2)[1]
with t.Context() as ctx:
    ctx.set_sequences(seqs); ctx.run(); ctx.guide_tree()
    for rep in range(2):
        print("run", rep, flush=True)
        sys.stderr.flush()
        t0 = time.perf_counter(); ctx.msa(); w = (time.perf_counter() - t0) * 1e3
        print("msa_ms", round(ctx.stats()["msa_ms"], 2), "wall", round(w, 2), flush=True)
