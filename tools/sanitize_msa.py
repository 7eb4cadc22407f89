"""A small progressive alignment for compute-sanitizer (memcheck / initcheck), checked against the oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import tweakseq_b200 as t
from oracle import pyoracle as o

rng = np.random.default_rng(3)
root = rng.choice(list("ARNDCQEGHILKMFPSTWYV"), 70)
seqs = []
for _ in range(14):
    s = [c if rng.random() > 0.25 else rng.choice(list("ARNDCQEGHILKMFPSTWYV")) for c in root if rng.random() > 0.05]
    seqs.append("".join(s))
seqs[3] = ""
with t.Context() as ctx:
    ctx.set_sequences(seqs); ctx.run()
    left, right, _ = ctx.guide_tree()
    rows, order = ctx.msa()
want, _ = o.msa([o.encode(s) for s in seqs], o.matrix(0), 11, 1, left, right)
assert rows == want, (rows, want)
print("sanitize_msa ok", len(rows), "rows x", len(rows[0]), "columns")
