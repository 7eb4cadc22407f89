"""Generates tests/golden/*.json.  Run from the repo root: python tests/golden/make_golden.py

The reference ships no vectors for this path, so the committed goldens come
from the independent numpy formulation in tests/np_gotoh.py, cross-checked here against the
naive pure-Python DP and the C oracle before anything is written.  Closed-form known answers
(hand-derived) are listed separately in tests/test_oracle.py.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from np_gotoh import gotoh_np, gotoh_py  # noqa: E402
from oracle import pyoracle as o  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
AA = "ARNDCQEGHILKMFPSTWYVBZX"


def main():
    rng = np.random.default_rng(20261017)
    pairs = []
    for alphabet, letters, params in ((0, AA, [(11, 1), (5, 2), (0, 0), (0, 3), (20, 0)]),
                                      (1, "ACGTN", [(10, 1), (3, 3), (0, 1)])):
        mat = o.matrix(alphabet)
        for go, ge in params:
            for _ in range(14):
                la, lb = int(rng.integers(0, 90)), int(rng.integers(0, 90))
                a = "".join(rng.choice(list(letters), la))
                b = "".join(rng.choice(list(letters), lb))
                if rng.random() < 0.4 and la > 4:   # related pair: mutate a copy
                    bb = list(a)
                    for k in range(len(bb)):
                        if rng.random() < 0.2:
                            bb[k] = str(rng.choice(list(letters)))
                    cut = int(rng.integers(0, len(bb)))
                    b = "".join(bb[:cut] + bb[cut + int(rng.integers(0, 4)):])
                ea, eb = o.encode(a, alphabet), o.encode(b, alphabet)
                s = gotoh_np(ea, eb, mat, go, ge)
                assert s == o.gotoh(ea, eb, mat, go, ge), (a, b, go, ge)
                if la * lb <= 1600:
                    assert s == gotoh_py(list(ea), list(eb), mat.tolist(), go, ge)
                pairs.append({"alphabet": alphabet, "go": go, "ge": ge, "a": a, "b": b, "score": s})
    json.dump(pairs, open(os.path.join(HERE, "pairs.json"), "w"), indent=0)

    # one small all-vs-all set with packed scores, self scores and bit-exact distances
    seqs = ["".join(rng.choice(list(AA[:20]), int(l))) for l in rng.integers(0, 70, 14)]
    seqs[3] = ""
    seqs[7] = seqs[2][:30] + seqs[2][33:]
    seqs[9] = seqs[2].lower()
    mat = o.matrix(0)
    enc = [o.encode(s, 0) for s in seqs]
    n = len(seqs)
    scores = []
    for i in range(n):
        for j in range(i + 1, n):
            s = gotoh_np(enc[i], enc[j], mat, 11, 1)
            assert s == o.gotoh(enc[i], enc[j], mat, 11, 1)
            scores.append(s)
    selfs = [int(sum(int(mat[x, x]) for x in e)) for e in enc]
    d = []
    k = 0
    for i in range(n):
        for j in range(i + 1, n):
            mn = min(selfs[i], selfs[j])
            d.append((1.0 - scores[k] / mn) if mn > 0 else 1.0)
            k += 1
    json.dump({"alphabet": 0, "go": 11, "ge": 1, "seqs": seqs, "scores": scores, "self": selfs,
               "distances_hex": [float(x).hex() for x in d]},
              open(os.path.join(HERE, "allpairs_small.json"), "w"), indent=0)
    print("wrote", len(pairs), "pairs and", len(scores), "all-vs-all scores")


if __name__ == "__main__":
    main()
