"""UPGMA guide-tree kernel time by CTA size (TSQ_UPGMA_THREADS) on configs[1] and on 4 000 sequences; trees must be identical."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import tweakseq_b200 as t
from tweakseq_b200 import synth
for name, seqs in (("configs[1] n=1000", synth.config(2)[1]), ("n=4000 x 120 aa", synth.protein(4000, 120, 7))):
    ref = None
    with t.Context() as ctx:
        ctx.set_sequences(seqs); ctx.run()
        for th in ("1024", "512", "256", "128", ""):
            if th: os.environ["TSQ_UPGMA_THREADS"] = th
            else: os.environ.pop("TSQ_UPGMA_THREADS", None)
            best = 1e9
            for _ in range(3):
                ctx.compute(); ctx.download()
                tree = ctx.guide_tree()
                best = min(best, ctx.stats()["tree_ms"])
            key = tuple(np.asarray(x).tobytes() for x in tree)
            if ref is None: ref = key
            print(name, "threads", th or "default", "tree_ms", round(best, 3), "same tree" if key == ref else "DIFFERENT TREE", flush=True)
