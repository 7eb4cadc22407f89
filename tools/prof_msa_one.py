"""One progressive alignment of a 100-sequence protein family (a caterpillar tree: 99 dependent merges), for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tweakseq_b200 as t
from tweakseq_b200 import synth
seqs = synth.protein(100, (200, 400, 300, 30), 1, family=True)
with t.Context() as ctx:
    ctx.set_sequences(seqs); ctx.run(); ctx.msa()
    print(ctx.stats()["msa_ms"])
