"""Developer GPU check for the wavefront kernel: timing + sampled oracle comparison on C4-like input."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import tweakseq_b200 as t
from tweakseq_b200 import synth
from oracle import pyoracle as o
n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
seqs = synth.nucleotide(n, 10000, 30000, 4)
ctx = t.Context(alphabet=1, flags=t.FLAG_NO_DISTANCES)
ctx.set_sequences(seqs)
t0 = time.time(); ctx.run(); tr = time.time() - t0
ctx.compute(); ctx.synchronize(); st = ctx.stats()
ctx.download(); s = ctx.scores()
rng = np.random.default_rng(1)
pi = rng.integers(0, n - 1, 48); pj = rng.integers(0, n, 48); keep = pi < pj; pi, pj = pi[keep], pj[keep]
enc = [o.encode(x, 1) for x in seqs]
t0 = time.time(); ref, cells = o.pair_list(enc, pi, pj, o.matrix(1), 10, 1, nthreads=os.cpu_count()); tc = time.time() - t0
idx = [t.pair_index(int(a), int(b), n) for a, b in zip(pi, pj)]
bad = int((ref != s[idx]).sum())
print(json.dumps({"case": f"c4_{n}x10-30kb", "pairs": len(s), "cells": st["cells"], "kernel_ms": st["kernel_ms"], "gcups": st["gcups_kernel"],
                  "checked_pairs": len(pi), "mismatch": bad, "cpu_gcups": cells / tc / 1e9, "run_s": tr}))
