"""Generates tests/golden/msa_small.json.  Run from the repo root: python tests/golden/make_golden_msa.py

Progressive alignments along the UPGMA tree of small sequence sets.  The rows and per-merge scores come
from the independent pure-Python statement of the spec (tests/np_msa.py: pair sums over member pairs,
Python integers) and are cross-checked against the C oracle before anything is written; the tree is the
oracle's UPGMA of the oracle's distances (pinned against scipy in tests/test_guide_tree.py), stored with
the case so that every consumer aligns along the same merges.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import np_msa  # noqa: E402
from oracle import pyoracle as o  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
PROT = "ARNDCQEGHILKMFPSTWYVBZX"
NUC = "ACGTN"


def family(rng, n, length, letters, mut, indel):
    root = rng.choice(list(letters), size=length)
    out = []
    for _ in range(n):
        s = []
        for ch in root:
            u = rng.random()
            if u < indel / 2:
                continue
            if u < indel:
                s.append(str(rng.choice(list(letters))))
            s.append(str(rng.choice(list(letters))) if rng.random() < mut else str(ch))
        out.append("".join(s))
    return out


def main():
    rng = np.random.default_rng(20261018)
    cases = []
    spec = [(0, 6, 40, 11, 1, 0.2, 0.08), (0, 9, 25, 11, 1, 0.5, 0.2), (0, 5, 30, 4, 2, 0.3, 0.1), (0, 7, 20, 0, 0, 0.3, 0.1),
            (1, 8, 50, 10, 1, 0.15, 0.06), (1, 4, 35, 3, 3, 0.3, 0.15), (0, 2, 30, 11, 1, 0.3, 0.1), (0, 1, 12, 11, 1, 0.0, 0.0)]
    for alphabet, n, length, go, ge, mut, indel in spec:
        letters = NUC if alphabet else PROT
        seqs = family(rng, n, length, "ACGT" if alphabet else PROT[:20], mut, indel)
        if n >= 7:
            seqs[2] = ""                       # an empty sequence: an all-gap row
            seqs[4] = seqs[4].lower()          # case is folded by the encoder; rows come back canonical
        mat = o.matrix(alphabet)
        enc = [o.encode(s, alphabet) for s in seqs]
        sc, _ = o.all_pairs(enc, mat, go, ge)
        selfs = np.array([o.self_score(e, mat) for e in enc], dtype=np.int32)
        left, right, _ = o.upgma(o.distances(sc, selfs), n)
        S = {a: {b: int(mat[i, j]) for j, b in enumerate(letters)} for i, a in enumerate(letters)}
        canon = ["".join(letters[v] for v in e) for e in enc]
        rows, scores = np_msa.progressive(canon, left.tolist(), right.tolist(), S, go, ge)
        orows, oscores = o.msa(enc, mat, go, ge, left, right, alphabet)
        assert rows == orows and scores == oscores.tolist(), (alphabet, n)
        cases.append({"alphabet": alphabet, "go": go, "ge": ge, "seqs": seqs, "left": left.tolist(), "right": right.tolist(),
                      "rows": rows, "merge_scores": scores})
    json.dump(cases, open(os.path.join(HERE, "msa_small.json"), "w"), indent=0)
    print("wrote", len(cases), "alignments")


if __name__ == "__main__":
    main()
