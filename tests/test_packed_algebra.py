"""The exactness argument of the packed 16-bit kernel, executed on the CPU.

Re-runs the kernel's arithmetic -- values stored as v + delta*(i+j) + BIAS in an unsigned
16-bit half, non-negative biased scores, finite E/F seeds instead of -inf, strip mining with a
(H, E) boundary column, padded query columns -- in plain Python integers with an explicit
[0, 65535] range check on every intermediate, and compares with the oracle.  This is a model
of csrc/gotoh16.cuh, not product code.
"""
import numpy as np
import pytest

from oracle import pyoracle as o


def bias_for(go, ge, delta, lpad):
    return 3 * go + 2 * ge + max(0, ge - delta) * 2 * (lpad + 1) + delta + 16


def packed_model(a, b, mat, go, ge, K):
    """a = query (columns), b = subject (rows); returns H[len(b)][len(a)]."""
    nsym = mat.shape[0]
    smin = int(mat.min())
    delta = (-smin + 1) // 2 if smin < 0 else 0
    gep, goep = ge - delta, go + ge - delta
    L, Ls = len(a), len(b)
    nstrips = (L + K - 1) // K
    lpad = max(nstrips * K, Ls)
    BIAS = bias_for(go, ge, delta, lpad)

    def chk(v):
        assert 0 <= v <= 65535, v
        return v

    sb = np.zeros((nsym + 1, nsym), dtype=np.int64)
    sb[:nsym] = mat.astype(np.int64) + 2 * delta
    assert sb.min() >= 0
    bnd = [None] * (Ls + 2)
    res = None
    for s in range(nstrips):
        j0 = s * K
        cols = [int(a[j0 + c]) if j0 + c < L else nsym for c in range(K)]
        H = [chk(BIAS - go - (j0 + c + 1) * gep) for c in range(K)]
        F = [chk(h - goep) for h in H]
        hdiag = chk(BIAS if j0 == 0 else BIAS - go - j0 * gep)
        for i in range(1, Ls + 1):
            if s == 0:
                Hl = chk(BIAS - go - i * gep)
                E = chk(Hl - goep)
            else:
                Hl, E = bnd[i]
            hd, hdiag = hdiag, Hl
            for c in range(K):
                t = chk(hd + int(sb[cols[c], int(b[i - 1])]))
                hd = H[c]
                h = max(t, E, F[c])
                H[c] = h
                hg = chk(h - goep)
                E = max(chk((E - gep)), hg)
                F[c] = max(chk(F[c] - gep), hg)
            if s + 1 < nstrips:
                bnd[i] = (H[K - 1], E)
        if j0 < L <= j0 + K:
            res = H[L - 1 - j0]
    return res - BIAS - delta * (L + Ls)


@pytest.mark.parametrize("K", [4, 7, 32])
@pytest.mark.parametrize("go,ge", [(11, 1), (0, 0), (5, 3), (2, 6)])
def test_packed_model_matches_oracle(K, go, ge):
    rng = np.random.default_rng(K * 100 + go * 10 + ge)
    mat = o.matrix(o.PROTEIN)
    for _ in range(12):
        a = rng.integers(0, 23, rng.integers(1, 60))
        b = rng.integers(0, 23, rng.integers(1, 60))
        assert packed_model(a, b, mat, go, ge, K) == o.gotoh(a, b, mat, go, ge)


def test_packed_model_nucleotide_and_positive_matrix():
    rng = np.random.default_rng(9)
    dna = o.matrix(o.NUCLEOTIDE)
    for _ in range(10):
        a, b = rng.integers(0, 5, rng.integers(1, 80)), rng.integers(0, 5, rng.integers(1, 80))
        assert packed_model(a, b, dna, 10, 1, 16) == o.gotoh(a, b, dna, 10, 1)
    pos = np.abs(o.matrix(o.PROTEIN))          # all-positive matrix: delta = 0
    for _ in range(5):
        a, b = rng.integers(0, 23, 40), rng.integers(0, 23, 33)
        assert packed_model(a, b, pos, 4, 2, 8) == o.gotoh(a, b, pos, 4, 2)


def test_range_bound_is_tight_enough_for_protein_defaults():
    # identical all-W sequences drive H to its maximum: 11 per column plus the skew
    L = 400
    a = np.full(L, 17)
    assert packed_model(a, a, o.matrix(o.PROTEIN), 11, 1, 50) == 11 * L
