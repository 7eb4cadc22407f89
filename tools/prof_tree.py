import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tweakseq_b200 as t
from tweakseq_b200 import synth
_, seqs = synth.config(2)
with t.Context() as ctx:
    ctx.set_sequences(seqs); ctx.run(); ctx.guide_tree()
    print("tree_ms", ctx.stats()["tree_ms"])
