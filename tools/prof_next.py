"""One pass through the 'next' rows for ncu: identity mode (gotoh32), guide tree (upgma), consensus."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tweakseq_b200 as t
from tweakseq_b200 import synth
_, seqs = synth.config(2)
with t.Context(flags=t.FLAG_IDENTITY) as ctx:
    ctx.set_sequences(seqs); ctx.run(); ctx.guide_tree()
    rows = [s[:300] for s in seqs]
    ctx.consensus(rows)
    print(ctx.stats())
