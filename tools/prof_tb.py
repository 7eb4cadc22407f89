"""Times tsq_align_pair (single-CTA anti-diagonal traceback) on a few pair sizes."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import tweakseq_b200 as t
rng = np.random.default_rng(1)
for alphabet, letters, sizes in ((0, "ARNDCQEGHILKMFPSTWYV", (300, 1000)), (1, "ACGT", (3000, 10000, 30000))):
    for L in sizes:
        seqs = ["".join(rng.choice(list(letters), L)) for _ in range(2)]
        with t.Context(alphabet=alphabet) as ctx:
            ctx.set_sequences(seqs); ctx.upload()
            ctx.align_pair(0, 1)
            t0 = time.perf_counter(); ra, rb, sc = ctx.align_pair(0, 1); dt = time.perf_counter() - t0
        print(f"traceback {L} x {L}: {1e3*dt:.2f} ms  ({L*L/dt/1e9:.2f} GCUPS)  columns {len(ra)} score {sc}", flush=True)
