"""Pins the oracle against the DEFINITION rather than another DP: every global alignment of two
short sequences is enumerated as a path of match / insert / delete moves and scored straight from
SURVEY.md 8c (substitution scores summed, every maximal run of k gap columns on one side costs
go + k*ge).  The maximum over all paths must equal the Gotoh score, and the most identities among
the maximal paths must equal the identity count of the keyed recurrence."""
import itertools

import numpy as np
import pytest

from oracle import pyoracle as o

MAT = o.matrix(o.PROTEIN)
DNA = o.matrix(o.NUCLEOTIDE)


def all_paths(n, m):
    """Every sequence of moves 'M' (1,1), 'D' (1,0), 'I' (0,1) from (0,0) to (n,m)."""
    out = []

    def rec(i, j, acc):
        if i == n and j == m:
            out.append("".join(acc))
            return
        if i < n and j < m:
            rec(i + 1, j + 1, acc + ["M"])
        if i < n:
            rec(i + 1, j, acc + ["D"])
        if j < m:
            rec(i, j + 1, acc + ["I"])

    rec(0, 0, [])
    return out


def score_path(path, a, b, mat, go, ge):
    i = j = 0
    s = nid = 0
    prev = ""
    for mv in path:
        if mv == "M":
            s += int(mat[a[i], b[j]])
            nid += int(a[i] == b[j])
            i += 1
            j += 1
        else:
            s -= ge + (go if mv != prev else 0)      # a new run opens whenever the move type changes
            if mv == "D":
                i += 1
            else:
                j += 1
        prev = mv
    return s, nid


def brute(a, b, mat, go, ge):
    best = None
    for p in all_paths(len(a), len(b)):
        k = score_path(p, a, b, mat, go, ge)
        if best is None or k > best:
            best = k
    return best


@pytest.mark.parametrize("go,ge", [(11, 1), (0, 1), (3, 0), (1, 2), (0, 0)])
def test_protein_all_paths(go, ge):
    rng = np.random.default_rng(100 + go * 7 + ge)
    for n, m in itertools.product(range(1, 6), range(1, 6)):
        for _ in range(3):
            a = rng.integers(0, 23, n).astype(np.uint8)
            b = rng.integers(0, 23, m).astype(np.uint8)
            if rng.random() < 0.5 and n <= m:                  # related pair: ties in the identity count
                b[:n] = a
            s, nid = brute(a, b, MAT, go, ge)
            assert o.gotoh(a, b, MAT, go, ge) == s
            assert o.gotoh_id(a, b, MAT, go, ge) == (s, nid)


@pytest.mark.parametrize("go,ge", [(10, 1), (2, 1), (0, 3)])
def test_nucleotide_all_paths(go, ge):
    rng = np.random.default_rng(7 + go)
    for n, m in [(1, 1), (2, 5), (5, 2), (4, 4), (6, 5), (3, 6)]:
        for _ in range(4):
            a = rng.integers(0, 5, n).astype(np.uint8)
            b = rng.integers(0, 5, m).astype(np.uint8)
            s, nid = brute(a, b, DNA, go, ge)
            assert o.gotoh(a, b, DNA, go, ge) == s
            assert o.gotoh_id(a, b, DNA, go, ge) == (s, nid)


def test_path_count_is_delannoy():
    assert [len(all_paths(k, k)) for k in range(5)] == [1, 3, 13, 63, 321]
