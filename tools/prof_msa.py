"""End-to-end time of tsq_msa (plan, kernels, copies) after a run, on protein families and on configs[1]."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import tweakseq_b200 as t
from tweakseq_b200 import synth

def one(name, seqs):
    with t.Context() as ctx:
        ctx.set_sequences(seqs); ctx.run(); ctx.guide_tree()
        t0 = time.perf_counter()
        rows, order = ctx.msa()
        wall = (time.perf_counter() - t0) * 1e3
        st = ctx.stats()
    print(f"{name}: n={len(seqs)} cols={len(rows[0]) if rows else 0} msa_ms={st['msa_ms']:.1f} wall_ms={wall:.1f} "
          f"tree_ms={st['tree_ms']:.2f} kernel_ms={st['kernel_ms']:.2f}", flush=True)

one("family 100 x ~300 (configs[0] family variant)", synth.protein(100, (200, 400, 300, 30), 1, family=True))
one("family 300 x 300", synth.protein(300, 300, 2, family=True))
one("family 1000 x 300", synth.protein(1000, 300, 2, family=True))
one("configs[1]: 1000 unrelated x 300", synth.config(2)[1])
