"""tsq_msa timings, best of three per workload (plan + kernels + copies), for A/B runs of two library builds in one call."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tweakseq_b200 as t
from tweakseq_b200 import synth

def one(name, seqs):
    best = 1e9
    with t.Context() as ctx:
        ctx.set_sequences(seqs)
        for _ in range(4):                      # the first pass warms the block cache; a new run() invalidates tree and alignment
            ctx.run(); ctx.guide_tree()
            ctx.msa()
            if _:
                best = min(best, ctx.stats()["msa_ms"])
    print(f"{name}: n={len(seqs)} msa_ms(best of 3)={best:.1f}", flush=True)

one("configs[0] 100 x 100-500 aa", synth.config(1)[1])
one("family 100 x ~300", synth.protein(100, (200, 400, 300, 30), 1, family=True))
one("family 300 x 300", synth.protein(300, 300, 2, family=True))
one("family 1000 x 300", synth.protein(1000, 300, 2, family=True))
one("configs[1] 1000 unrelated x 300", synth.config(2)[1])
