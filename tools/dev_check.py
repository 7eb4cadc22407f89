"""Developer GPU check: CUDA path vs oracle on a few sets + timing. Run under gpurun."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import tweakseq_b200 as t
from tweakseq_b200 import synth
from oracle import pyoracle as o

def check(name, seqs, alphabet=0, go=-1, ge=-1, full=True, sample=20000, reps=1, flags=0):
    ctx = t.Context(alphabet=alphabet, gap_open=go, gap_extend=ge, flags=flags)
    ctx.set_sequences(seqs)
    t0 = time.time(); ctx.run(); t_run = time.time() - t0
    best = 1e30
    for _ in range(reps):
        ctx.compute(); ctx.synchronize(); best = min(best, ctx.stats()["kernel_ms"])
    ctx.download()
    s = ctx.scores(); d = ctx.distances(); st = ctx.stats()
    enc = [o.encode(x, alphabet) for x in seqs]
    mat = o.matrix(alphabet)
    g = (11 if alphabet == 0 else 10) if go < 0 else go
    e = 1 if ge < 0 else ge
    n = len(seqs)
    t0 = time.time()
    if full:
        ref, cells = o.all_pairs(enc, mat, g, e, nthreads=os.cpu_count())
        bad = np.nonzero(ref != s)[0]
        selfs = np.array([o.self_score(x, mat) for x in enc], dtype=np.int32)
        dref = o.distances(ref, selfs)
        dbad = int((dref != d).sum())
    else:
        rng = np.random.default_rng(7)
        pi = rng.integers(0, n - 1, sample); pj = rng.integers(0, n, sample)
        keep = pi < pj; pi, pj = pi[keep], pj[keep]
        ref, cells = o.pair_list(enc, pi, pj, mat, g, e, nthreads=os.cpu_count())
        idx = np.array([t.pair_index(int(a), int(b), n) for a, b in zip(pi, pj)])
        bad = np.nonzero(ref != s[idx])[0]
        dbad = -1
    t_cpu = time.time() - t0
    print(json.dumps({"case": name, "n": n, "pairs": int(len(s)), "mismatch": int(len(bad)), "dist_mismatch": dbad,
                      "kernel_ms": best, "gcups": st["cells"] / best / 1e6 if best > 0 else 0, "run_s": round(t_run, 4),
                      "K": st["strip_width"], "cpu_gcups": cells / t_cpu / 1e9, "cells": st["cells"]}), flush=True)
    if len(bad):
        print("  first bad:", bad[:10], ref[bad[:10]], (s if full else s[idx])[bad[:10]])
    ctx.close()
    return len(bad) == 0

if __name__ == "__main__":
    rng = np.random.default_rng(3)
    aa = "ARNDCQEGHILKMFPSTWYVBZX"
    rag = ["".join(rng.choice(list(aa), int(l))) for l in rng.integers(0, 130, 75)]
    rag[5] = ""; rag[9] = "acdefg-hik.lmn pq"; rag[11] = "W"
    ok = check("ragged75", rag)
    ok &= check("ragged75_go5_ge2", rag, go=5, ge=2)
    ok &= check("ragged75_go0_ge0", rag, go=0, ge=0)
    ok &= check("c1_100x300", synth.config(1)[1])
    ok &= check("c1_family", synth.protein(100, (200, 400, 300, 30), 1, family=True))
    ok &= check("n333_len1..700", ["".join(rng.choice(list(aa[:20]), int(l))) for l in rng.integers(1, 700, 333)])
    which = sys.argv[1:] or ["c2"]
    if "c2" in which:
        ok &= check("c2_1000x300", synth.config(2)[1], full=True, reps=5)
    if "c3s" in which:
        ok &= check("c3_small_3000x400", synth.config(3, 0.3)[1], full=False, reps=3)
    if "c3" in which:
        ok &= check("c3_10000x400", synth.config(3)[1], full=False, reps=2, flags=t.FLAG_NO_DISTANCES)
    print("ALL OK" if ok else "FAILURES")
