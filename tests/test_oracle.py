"""Pins the CPU oracle: hand-derived known answers (SURVEY.md 8c list), the independent numpy
formulation, the committed golden vectors, and the reference's own matrix source."""
import json
import os
import re

import numpy as np
import pytest

from np_gotoh import gotoh_np, gotoh_py
from published_vectors import VECTORS
from oracle import pyoracle as o

HERE = os.path.dirname(os.path.abspath(__file__))
AA = "ARNDCQEGHILKMFPSTWYVBZX"
MAT = o.matrix(o.PROTEIN)
GO, GE = 11, 1


def sc(a, b, alphabet=o.PROTEIN, go=None, ge=1):
    return o.score_str(a, b, alphabet, go, ge)


def diag(s):
    e = o.encode(s)
    return sum(int(MAT[x, x]) for x in e)


# ---- matrix / alphabet come from the reference -------------------------------------------------
def test_matrix_is_symmetric_and_in_range():
    assert MAT.shape == (23, 23)
    assert (MAT == MAT.T).all()
    assert MAT.min() == -4 and MAT.max() == 11
    assert int(MAT[17, 17]) == 11 and int(MAT[4, 4]) == 9  # W/W, C/C


REF = "/root/reference/tweakseq/Core/Annotations/Consensus.cpp"


@pytest.mark.skipif(not os.path.exists(REF), reason="reference tree not present on this box")
def test_matrix_and_map_match_reference_source():
    src = open(REF).read()
    body = src[src.index("BLOSUM62[23][23]"):src.index("BLOSUM62map")]
    rows = re.findall(r"\{([^{}]*)\}", body)
    ref = np.array([[int(x) for x in r.replace(" ", "").split(",") if x] for r in rows])
    assert ref.shape == (23, 23)
    assert (ref == MAT).all()
    mbody = src[src.index("BLOSUM62map[26]"):]
    mbody = mbody[mbody.index("{") + 1:mbody.index("}")]
    refmap = [int(x) for x in mbody.replace("\n", " ").replace(" ", "").split(",") if x]
    assert len(refmap) == 26
    for k in range(26):
        assert o.encode(chr(65 + k))[0] == refmap[k]


def test_letter_mapping():
    assert list(o.encode(AA)) == list(range(23))
    assert list(o.encode(AA.lower())) == list(range(23))
    assert list(o.encode("JOUX*1?")) == [22] * 7           # J,O,U,X and non-letters -> X
    assert list(o.encode("A-C.D E\tF\nG")) == list(o.encode("ACDEFG"))  # gaps/whitespace dropped
    assert list(o.encode("ACGTUNRYacgtn-", o.NUCLEOTIDE)) == [0, 1, 2, 3, 3, 4, 4, 4, 0, 1, 2, 3, 4]


# ---- closed-form known answers -----------------------------------------------------------------
@pytest.mark.parametrize("k", [1, 2, 7, 50])
def test_identical_sequences_score_sum_of_diagonal(k):
    assert sc("W" * k, "W" * k) == 11 * k
    s = (AA * 3)[:k]
    assert sc(s, s) == diag(s)


def test_single_substitution():
    a = "ACDEFGHIKLMNPQRSTVWY"
    b = a[:7] + "W" + a[8:]      # I -> W at position 7
    assert sc(a, b) == diag(a) - int(MAT[9, 9]) + int(MAT[9, 17])


@pytest.mark.parametrize("k", [1, 2, 5])
@pytest.mark.parametrize("where", ["start", "middle", "end"])
def test_single_gap_closed_form(k, where):
    a = "WCWCWCWCWCWCWCWCWCWC"          # high-scoring, so the optimal path keeps the matches
    if where == "start":
        b = a[k:]
    elif where == "end":
        b = a[:-k]
    else:
        b = a[:10] + a[10 + k:]
    assert sc(a, b) == diag(b) - (GO + k * GE)
    assert sc(b, a) == diag(b) - (GO + k * GE)


def test_empty_sequences():
    assert sc("", "") == 0
    assert sc("", "ACD") == -(GO + 3 * GE)
    assert sc("ACDEF", "") == -(GO + 5 * GE)
    assert sc("---", "AC") == -(GO + 2 * GE)


def test_all_x():
    assert sc("XXXX", "XXXX") == -4
    assert sc("JOU*", "XXXX") == -4


def test_gapped_input_equals_ungapped():
    assert sc("AC-DE..FGH IK", "ACDQFGHIK") == sc("ACDEFGHIK", "ACDQFGHIK")


def test_tie_cases_small_gap_open():
    # go = 0 makes E/F/H ties common; the three implementations must still agree
    rng = np.random.default_rng(5)
    for _ in range(60):
        a = rng.integers(0, 23, rng.integers(1, 25))
        b = rng.integers(0, 23, rng.integers(1, 25))
        for go, ge in ((0, 0), (0, 1), (1, 0), (0, 4)):
            assert o.gotoh(a, b, MAT, go, ge) == gotoh_py(list(a), list(b), MAT.tolist(), go, ge)


def test_nucleotide_defaults():
    assert sc("ACGT", "ACGT", o.NUCLEOTIDE) == 20
    assert sc("ACGT", "ACGA", o.NUCLEOTIDE) == 15 - 4
    assert sc("ACGTN", "ACGTN", o.NUCLEOTIDE) == 19
    assert sc("ACGTACGT", "ACGACGT", o.NUCLEOTIDE) == 35 - 11


# ---- cross-checks --------------------------------------------------------------------------------
def test_symmetry():
    rng = np.random.default_rng(11)
    for _ in range(200):
        a = rng.integers(0, 23, rng.integers(0, 60))
        b = rng.integers(0, 23, rng.integers(0, 60))
        assert o.gotoh(a, b, MAT, GO, GE) == o.gotoh(b, a, MAT, GO, GE)


def test_oracle_vs_independent_numpy_random_pairs():
    rng = np.random.default_rng(2026)
    n = 0
    for go, ge in ((11, 1), (5, 2), (0, 0), (3, 0), (25, 3)):
        for _ in range(400):
            la, lb = int(rng.integers(0, 200)), int(rng.integers(0, 200))
            a, b = rng.integers(0, 23, la), rng.integers(0, 23, lb)
            assert o.gotoh(a, b, MAT, go, ge) == gotoh_np(a, b, MAT, go, ge)
            n += 1
    dna = o.matrix(o.NUCLEOTIDE)
    for _ in range(300):
        a, b = rng.integers(0, 5, rng.integers(0, 400)), rng.integers(0, 5, rng.integers(0, 400))
        assert o.gotoh(a, b, dna, 10, 1) == gotoh_np(a, b, dna, 10, 1)
    assert n == 2000


def test_long_pair_heap_path():
    rng = np.random.default_rng(1)
    a, b = rng.integers(0, 5, 3000), rng.integers(0, 5, 2500)   # > 1024 columns: malloc path
    dna = o.matrix(o.NUCLEOTIDE)
    assert o.gotoh(a, b, dna, 10, 1) == gotoh_np(a, b, dna, 10, 1)


def test_golden_pairs():
    pairs = json.load(open(os.path.join(HERE, "golden", "pairs.json")))
    assert len(pairs) >= 100
    for p in pairs:
        assert o.score_str(p["a"], p["b"], p["alphabet"], p["go"], p["ge"]) == p["score"], p


def test_golden_allpairs_scores_and_distances():
    g = json.load(open(os.path.join(HERE, "golden", "allpairs_small.json")))
    enc = [o.encode(s, g["alphabet"]) for s in g["seqs"]]
    for nt in (1, 3):
        s, cells = o.all_pairs(enc, MAT, g["go"], g["ge"], nthreads=nt)
        assert s.tolist() == g["scores"]
    selfs = np.array([o.self_score(e, MAT) for e in enc], dtype=np.int32)
    assert selfs.tolist() == g["self"]
    d = o.distances(s, selfs)
    assert [float(x).hex() for x in d] == g["distances_hex"]
    lens = np.array([len(e) for e in enc], dtype=np.int64)
    assert cells == (lens.sum() ** 2 - (lens ** 2).sum()) // 2


def test_pair_range_and_pair_list_agree():
    rng = np.random.default_rng(4)
    enc = [rng.integers(0, 23, rng.integers(0, 50)).astype(np.uint8) for _ in range(37)]
    full, _ = o.all_pairs(enc, MAT, GO, GE, nthreads=2)
    n = len(enc)
    part, _ = o.all_pairs(enc, MAT, GO, GE, nthreads=3, pair_begin=101, pair_end=555)
    assert (part == full[101:555]).all()
    iu, ju = np.triu_indices(n, 1)
    assert [o.pair_index(int(i), int(j), n) for i, j in zip(iu[:50], ju[:50])] == list(range(50))
    sel = rng.integers(0, len(iu), 200)
    lst, _ = o.pair_list(enc, iu[sel], ju[sel], MAT, GO, GE, nthreads=2)
    assert (lst == full[sel]).all()


def test_distance_function():
    assert o.distance(10, 10, 20) == 0.0
    assert o.distance(5, 10, 20) == 0.5
    assert o.distance(-5, 10, 20) == 1.5
    assert o.distance(3, 0, 20) == 1.0 and o.distance(3, -4, 20) == 1.0
    assert o.distance(7, 30, 21) == 1.0 - 7.0 / 21.0


def test_simd_cpu_baseline_kernel_is_bit_identical_to_the_scalar_oracle():
    """oracle/gotoh_simd.c (bench.py's CPU arm: one subject per int16 lane) against tsq_oracle_all_pairs: ragged
    lengths, empties, several gap models, both alphabets, a row sub-range, and sequences long enough that the int16
    bound fails and the kernel must hand the pair to the scalar routine."""
    rng = np.random.default_rng(77)
    AA = "ARNDCQEGHILKMFPSTWYVBZX"
    seqs = ["".join(rng.choice(list(AA), int(l))) for l in rng.integers(0, 130, 70)] + ["", "W", "acd-ef"]
    enc = [o.encode(s) for s in seqs]
    mat = o.matrix(0)
    for go, ge in [(11, 1), (5, 2), (0, 0), (0, 3), (30, 0), (1, 7)]:
        a, ca = o.all_pairs(enc, mat, go, ge, nthreads=2)
        b, cb = o.rows_simd(enc, mat, go, ge, nthreads=3)
        assert (a == b).all() and ca == cb, (go, ge)
    n = len(enc)
    a, _ = o.all_pairs(enc, mat, 11, 1, nthreads=2)
    part, _ = o.rows_simd(enc, mat, 11, 1, nthreads=2, row_begin=9, row_end=41)
    assert (part == a[o.pair_index(9, 10, n):o.pair_index(41, 42, n)]).all()
    nt = ["".join(rng.choice(list("ACGTN"), int(l))) for l in rng.integers(1, 300, 40)]
    encn = [o.encode(s, 1) for s in nt]
    a, _ = o.all_pairs(encn, o.matrix(1), 10, 1, nthreads=2)
    b, _ = o.rows_simd(encn, o.matrix(1), 10, 1, nthreads=2)
    assert (a == b).all()
    long_ = ["W" * 3100, "W" * 3100, "".join(rng.choice(list(AA[:20]), 3000)), "A" * 40]     # 11 * 3100 > int16
    encl = [o.encode(s) for s in long_]
    a, _ = o.all_pairs(encl, mat, 11, 1, nthreads=2)
    b, _ = o.rows_simd(encl, mat, 11, 1, nthreads=2)
    assert (a == b).all() and a[0] == 11 * 3100


@pytest.mark.parametrize("name,alphabet,a,b,mat,go,ge,published", VECTORS, ids=[v[0] for v in VECTORS])
def test_published_alignment_scores(name, alphabet, a, b, mat, go, ge, published):
    """The anchors outside this repository (tests/published_vectors.py): the scalar oracle, its SIMD kernel, the
    independent numpy and pure-Python statements all give the published optimum, in both argument orders."""
    ea, eb = o.encode(a, alphabet), o.encode(b, alphabet)
    mat = o.matrix(alphabet) if mat is None else mat
    assert o.gotoh(ea, eb, mat, go, ge) == published
    assert o.gotoh(eb, ea, mat, go, ge) == published
    assert gotoh_np(ea, eb, mat, go, ge) == published
    assert gotoh_py(ea, eb, mat, go, ge) == published
    simd, _ = o.rows_simd([ea, eb], mat, go, ge, nthreads=1)
    assert simd.tolist() == [published]
