"""SURVEY.md 8f-2, second half: emitting a pairwise alignment.  CPU leg pins the oracle's traceback
(the rows spell the inputs, re-scoring the rows from the gap-run definition gives the Gotoh score,
the fixed tie rules hold); the gpu leg requires the CUDA path to emit the SAME strings."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import pyoracle as o

MAT = o.matrix(o.PROTEIN)
DNA = o.matrix(o.NUCLEOTIDE)
AA = "ARNDCQEGHILKMFPSTWYVBZX"


def rescore(ra, rb, mat, letters, go, ge):
    """Score of two gapped rows straight from the definition (a run of k gap columns costs go + k*ge)."""
    assert len(ra) == len(rb)
    s, prev = 0, ""
    for x, y in zip(ra, rb):
        assert not (x == "-" and y == "-")
        if x == "-" or y == "-":
            kind = "a" if x == "-" else "b"
            s -= ge + (go if kind != prev else 0)
            prev = kind
        else:
            s += int(mat[letters.index(x), letters.index(y)])
            prev = ""
    return s


def check_oracle(a, b, mat, letters, go, ge, alphabet):
    ra, rb, sc = o.traceback(np.array(a, np.uint8), np.array(b, np.uint8), mat, go, ge, alphabet)
    assert ra.replace("-", "") == "".join(letters[x] for x in a)
    assert rb.replace("-", "") == "".join(letters[x] for x in b)
    assert sc == o.gotoh(np.array(a, np.uint8), np.array(b, np.uint8), mat, go, ge)
    assert rescore(ra, rb, mat, letters, go, ge) == sc
    return ra, rb, sc


@settings(max_examples=300, deadline=None)
@given(st.lists(st.integers(0, 22), max_size=40), st.lists(st.integers(0, 22), max_size=40),
       st.tuples(st.integers(0, 20), st.integers(0, 6)))
def test_oracle_traceback_is_a_valid_optimal_alignment_protein(a, b, g):
    check_oracle(a, b, MAT, AA, g[0], g[1], o.PROTEIN)


@settings(max_examples=150, deadline=None)
@given(st.lists(st.integers(0, 4), max_size=60), st.lists(st.integers(0, 4), max_size=60),
       st.tuples(st.integers(0, 12), st.integers(0, 4)))
def test_oracle_traceback_is_a_valid_optimal_alignment_nucleotide(a, b, g):
    check_oracle(a, b, DNA, "ACGTN", g[0], g[1], o.NUCLEOTIDE)


def test_oracle_traceback_known_answers_and_tie_rules():
    enc = lambda s: list(o.encode(s))
    assert check_oracle(enc("WWCWW"), enc("WWWW"), MAT, AA, 11, 1, 0)[:2] == ("WWCWW", "WW-WW")
    assert check_oracle(enc("ACD"), enc(""), MAT, AA, 11, 1, 0) == ("ACD", "---", -14)
    assert check_oracle(enc(""), enc(""), MAT, AA, 11, 1, 0) == ("", "", 0)
    # free gaps (go = ge = 0), AW vs WA: W/W aligned with two free gaps scores 11; the diagonal is preferred
    # wherever it is optimal, and the leading gap lands in the row the E-before-F rule selects
    ra, rb, sc = check_oracle(enc("AW"), enc("WA"), MAT, AA, 0, 0, 0)
    assert sc == 11 and (ra, rb) == ("AW-", "-WA")
    # equal-score choice between a diagonal mismatch and gaps: the diagonal wins
    ra, rb, _ = check_oracle(enc("AR"), enc("RA"), MAT, AA, 0, 0, 0)
    assert "-" in ra or (ra, rb) == ("AR", "RA")


# ---- gpu leg ---------------------------------------------------------------------------------------------
def _rand(rng, n, lo, hi, letters):
    return ["".join(rng.choice(list(letters), int(l))) for l in rng.integers(lo, hi, n)]


@pytest.mark.gpu
@pytest.mark.parametrize("go,ge", [(11, 1), (0, 0), (3, 2), (0, 4), (25, 0)])
def test_gpu_traceback_emits_the_oracle_strings_protein(go, ge):
    import tweakseq_b200 as t
    from tweakseq_b200 import synth
    rng = np.random.default_rng(go * 10 + ge)
    seqs = synth.protein(8, (40, 160, 100, 30), 5, family=True) + _rand(rng, 10, 0, 140, AA) + ["", "W", "acd-ef"]
    enc = [o.encode(s) for s in seqs]
    with t.Context(gap_open=go, gap_extend=ge) as ctx:
        ctx.set_sequences(seqs)
        ctx.run()
        scores = ctx.scores()
        n = len(seqs)
        pairs = [(int(i), int(j)) for i, j in rng.integers(0, n, (60, 2))] + [(0, 1), (n - 3, n - 1), (n - 1, 0), (2, 2)]
        for i, j in pairs:
            ra, rb, sc = ctx.align_pair(i, j)
            assert (ra, rb, sc) == o.traceback(enc[i], enc[j], MAT, go, ge)
            assert rescore(ra, rb, MAT, AA, go, ge) == sc
            if i != j:
                a, b = min(i, j), max(i, j)
                assert sc == scores[a * n - a * (a + 1) // 2 + (b - a - 1)]      # same score as the matrix entry


@pytest.mark.gpu
def test_gpu_traceback_nucleotide_and_long_pair():
    import tweakseq_b200 as t
    from tweakseq_b200 import synth
    fam = synth.nucleotide(3, 2500, 3300, 9, family=True)
    rng = np.random.default_rng(3)
    seqs = fam + _rand(rng, 4, 1, 1100, "ACGTN")
    enc = [o.encode(s, 1) for s in seqs]
    with t.Context(alphabet=1) as ctx:
        ctx.set_sequences(seqs)
        ctx.upload()                                  # no all-vs-all run needed for a single pair
        for i, j in [(0, 1), (2, 0), (1, 5), (6, 3), (4, 4)]:
            assert ctx.align_pair(i, j) == o.traceback(enc[i], enc[j], DNA, 10, 1, o.NUCLEOTIDE)


@pytest.mark.gpu
def test_gpu_traceback_errors_and_backend_object():
    import tweakseq_b200 as t
    with t.Context() as ctx:
        ctx.set_sequences(["ACDEF", "ACEF"])
        with pytest.raises(t.TsqError):
            ctx.align_pair(0, 1)                      # before tsq_upload
        ctx.upload()
        with pytest.raises(t.TsqError):
            ctx.align_pair(0, 2)                      # index out of range
        assert ctx.align_pair(0, 1) == ("ACDEF", "AC-EF", 4 + 9 + 5 + 6 - 12)
    tool = t.B200Gotoh()
    assert tool.pairwise_alignment("WWCWW", "wwww") == ("WWCWW", "WW-WW", 32)
