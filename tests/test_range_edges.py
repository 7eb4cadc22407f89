"""The packed 16-bit kernels AT the edge of their exactness range, on the GPU, against the oracle.

`gotoh16_kernel` keeps every value as v + delta*(i+j) + BIAS in an unsigned 16-bit half and is exact only
while nothing leaves [0, 65535]; the host admits a sequence only under a proven bound
(tsq_get_limits).  Random sequences never come near that bound -- these inputs do: the highest-scoring
alignments (poly-W, poly-A) at the longest admitted length and one below, the most negative ones (large gap
costs with ge > delta, unrelated and very unequal sequences at the limit), and matrices that use the whole
int8 range.  `wave16_kernel` holds a moving 16-bit window instead: its window bound is
exercised just under the limit and just over it, where the host must select the 32-bit kernel.  A TMA
barrier that never completes must surface as TSQ_ERR_CUDA, not as a wrong score."""
import os

import numpy as np
import pytest

import tweakseq_b200 as t
from tweakseq_b200 import synth
from oracle import pyoracle as o

pytestmark = pytest.mark.gpu
NT = os.cpu_count() or 1
AA20 = "ARNDCQEGHILKMFPSTWYV"


def rand(rng, l, letters=AA20):
    return "".join(rng.choice(list(letters), int(l)))


def check(seqs, alphabet=0, go=-1, ge=-1, matrix=None, flags=0):
    with t.Context(alphabet=alphabet, gap_open=go, gap_extend=ge, matrix=matrix, flags=flags | t.FLAG_NO_DISTANCES) as ctx:
        lim = ctx.limits()
        ctx.set_sequences(seqs)
        ctx.run()
        s, st = ctx.scores(), ctx.stats()
    enc = [o.encode(x, alphabet) for x in seqs]
    mat = o.matrix(alphabet) if matrix is None else np.ascontiguousarray(matrix, dtype=np.int8)
    g = (10 if alphabet else 11) if go < 0 else go
    e = 1 if ge < 0 else ge
    ref, cells = o.all_pairs(enc, mat, g, e, nthreads=NT)
    bad = np.nonzero(s != ref)[0]
    assert len(bad) == 0, (bad[:10], s[bad[:10]], ref[bad[:10]])
    return lim, st, cells, ref


def test_limits_of_the_default_models():
    with t.Context() as ctx:
        lim = ctx.limits()
    # BIAS 53 + 11 L + 2 delta (L + 1) + 16 <= 65535 with the 64-residue padding margin (tsq_api.cpp: fits16)
    assert lim["delta"] == 2 and lim["max_len_packed"] == 4300 and lim["max_len_inter"] == 4300 and not lim["inter_is_32bit"]
    assert lim["wave_packed"] and lim["wave_window"] <= lim["wave_window_max"] == 30000
    with t.Context(alphabet=1) as ctx:
        lim = ctx.limits()
    assert 7100 < lim["max_len_packed"] < 7300 and lim["wave_packed"]


def test_poly_w_at_the_longest_packed_length_and_one_below():
    """Self alignments of poly-W climb 11 per cell: H reaches BIAS + 11 L + 2 delta L, the top of the range."""
    rng = np.random.default_rng(601)
    with t.Context() as ctx:
        L = ctx.limits()["max_len_packed"]
    seqs = (["W" * L, "W" * L, "W" * (L - 1), "W" * (L - 1), "W" * (L - 2)] +
            [rand(rng, L), rand(rng, L - 1), "C" * L, "A" * L, "WC" * (L // 2)] +      # unrelated / low-scoring at the limit
            ["W" * 37, "W", rand(rng, 300), rand(rng, 3000)])                            # very unequal lengths: long end gaps
    lim, st, cells, ref = check(seqs)
    assert st["cells_s16"] == cells and st["cells_s32"] == 0                           # everything ran on the packed kernel
    assert ref[0] == 11 * L and ref.max() == 11 * L
    # one residue more and the sequence must leave the packed kernel -- and still be exact
    seqs2 = ["W" * (L + 1), "W" * (L + 1), "W" * L, rand(rng, 200)]
    lim, st2, cells2, ref2 = check(seqs2)
    assert st2["cells_s32"] > 0 and ref2[0] == 11 * (L + 1)


def test_poly_a_nucleotide_at_its_limit():
    rng = np.random.default_rng(602)
    with t.Context(alphabet=1) as ctx:
        L = ctx.limits()["max_len_packed"]
    seqs = ["A" * L, "A" * L, "A" * (L - 1), "C" * L, rand(rng, L, "ACGT"), rand(rng, L - 3, "ACGT"), "N" * L, "ACGT", "A" * 999]
    lim, st, cells, ref = check(seqs, alphabet=1)
    assert st["cells_s16"] == cells and ref[0] == 5 * L


@pytest.mark.parametrize("go,ge", [(200, 5), (4096, 3), (40, 30), (0, 64)])
def test_most_negative_paths_large_gap_costs_at_the_limit(go, ge):
    """ge > delta makes the skewed values FALL along a gap run: the BIAS has to hold -(go + L ge) above zero.
    Unrelated, unequal and near-empty sequences at the admitted limit drive H to its lowest values."""
    rng = np.random.default_rng(603 + go + ge)
    with t.Context(gap_open=go, gap_extend=ge) as ctx:
        lim = ctx.limits()
    L = lim["max_len_inter"]
    assert L >= 64
    if lim["inter_is_32bit"]:
        L = min(L, 1500)
    seqs = (["W" * L, "A" * L, "C" * (L - 1), rand(rng, L), rand(rng, L), "W" * (L // 2), "P", "PG", rand(rng, 11)] +
            [rand(rng, int(l)) for l in rng.integers(1, L, 12)])
    lim, st, cells, ref = check(seqs, go=go, ge=ge)
    if not lim["inter_is_32bit"]:
        assert st["cells_s16"] == cells
    assert ref.min() < -(go + ge)                       # gap-dominated alignments are in the set


@pytest.mark.parametrize("hi,lo", [(127, -128), (127, 0), (0, -128), (3, -1)])
def test_matrices_over_the_whole_int8_range(hi, lo):
    rng = np.random.default_rng(604 + hi - lo)
    m = rng.integers(lo, hi + 1, (23, 23))
    m = np.triu(m) + np.triu(m, 1).T
    m[17, 17] = hi                                      # W against W: the largest value on the diagonal
    m[0, 17] = m[17, 0] = lo                            # A against W: the smallest off it
    m = m.astype(np.int8)
    with t.Context(matrix=m, gap_open=9, gap_extend=2) as ctx:
        lim = ctx.limits()
    L = max(lim["max_len_inter"], 8)
    L = min(L, 2000)
    seqs = (["W" * L, "W" * L, "A" * L, "W" * (L - 1), rand(rng, L), rand(rng, max(1, L - 1)), "AW" * (L // 2), "W", "A"] +
            [rand(rng, int(l)) for l in rng.integers(1, L + 1, 14)])
    lim, st, cells, ref = check(seqs, go=9, ge=2, matrix=m)
    assert ref[0] == hi * L
    if lim["max_len_packed"] >= 64:
        assert not lim["inter_is_32bit"] and st["cells_s16"] == cells


def test_a_matrix_too_wide_for_16_bits_runs_on_the_32_bit_inter_task_kernel():
    m = np.full((23, 23), -128, dtype=np.int8)
    np.fill_diagonal(m, 127)
    with t.Context(matrix=m, gap_open=4000, gap_extend=1000) as ctx:
        lim = ctx.limits()
    assert lim["inter_is_32bit"] and lim["max_len_packed"] < 64
    rng = np.random.default_rng(605)
    seqs = ["W" * 900, "W" * 900, "A" * 700, rand(rng, 800), rand(rng, 64), "W"]
    check(seqs, go=4000, ge=1000, matrix=m)


# ---- packed wavefront kernel: the moving 16-bit window ---------------------------------------------------------
def _long_set(rng, alphabet_letters, lens):
    return [rand(rng, l, alphabet_letters) for l in lens]


@pytest.mark.parametrize("go,packed,cols", [(20, True, 24), (21, True, 16), (31, True, 16), (32, False, 16)])
def test_wave16_window_just_under_and_just_over_the_limit(go, packed, cols):
    """Nucleotides: window = (32 * columns per lane + 204) * (5 + go + 1 + 4).  24 columns per lane while that fits
    30 000 (go = 20: 972 * 30 = 29 160), 16 columns beyond (go = 21: 716 * 31; go = 31: 716 * 41 = 29 356), and from
    go = 32 (716 * 42 = 30 072) the host must take the 32-bit kernel.  Sequences beyond the inter-task limit;
    poly-A pairs (steepest climb), unrelated ones (steepest fall) and very unequal lengths."""
    rng = np.random.default_rng(606)
    with t.Context(alphabet=1, gap_open=go, gap_extend=1) as ctx:
        lim = ctx.limits()
    assert bool(lim["wave_packed"]) == packed
    assert (lim["wave_window"] <= lim["wave_window_max"]) == packed
    assert lim["wave_window"] == (32 * cols + 204) * (5 + go + 1 + 4)
    Lw = lim["max_len_inter"] + 1
    seqs = (["A" * (Lw + 700), "A" * (Lw + 300), "C" * (Lw + 10), rand(rng, Lw + 1200, "ACGT"), rand(rng, Lw, "ACGT")] +
            ["A" * 1500, rand(rng, 2200, "ACGT"), "ACGT" * 300, "G"])
    lim, st, cells, ref = check(seqs, alphabet=1, go=go, ge=1)
    assert st["cells_s32"] > 0.8 * cells
    assert ref[0] == 5 * (Lw + 300) - (go + 400)       # poly-A against poly-A: all matches and one end gap


def test_wave16_protein_window_and_extreme_rows():
    """Proteins, 8 columns per lane: window = 460 * (max|S| + go + ge + 2 delta)."""
    rng = np.random.default_rng(607)
    with t.Context(gap_open=35, gap_extend=2) as ctx:
        lim = ctx.limits()
    assert lim["wave_packed"] and lim["wave_window"] == 460 * (11 + 35 + 2 + 4)
    with t.Context(gap_open=49, gap_extend=2) as ctx:
        assert not ctx.limits()["wave_packed"]            # 460 * 66 = 30 360
    Lw = lim["max_len_inter"] + 1
    seqs = ["W" * (Lw + 50), "W" * Lw, rand(rng, Lw + 20), "C" * 3000, rand(rng, 900), "W" * 5]
    check(seqs, go=35, ge=2)


def test_a_tma_barrier_that_never_completes_is_an_error_not_a_score(monkeypatch):
    """tma_stage.cuh: the warp gives up after 2^29 clocks, raises the context's device fault word and abandons
    its task; tsq_download must return TSQ_ERR_CUDA.  TSQ_FAULT_INJECT=tma makes the first task arm a tile
    barrier without issuing its copy."""
    seqs = synth.nucleotide(6, 9000, 9500, 11)
    monkeypatch.setenv("TSQ_FAULT_INJECT", "tma")
    with t.Context(alphabet=1) as ctx:
        ctx.set_sequences(seqs)
        with pytest.raises(t.TsqError) as e:
            ctx.run()
        assert e.value.status == -3 and "TMA" in str(e.value)
        with pytest.raises(t.TsqError):
            ctx.scores()                                   # nothing is handed out
        monkeypatch.delenv("TSQ_FAULT_INJECT")
        ctx.run()                                          # the context is usable again
        s = ctx.scores()
    enc = [o.encode(x, 1) for x in seqs]
    ref, _ = o.all_pairs(enc, o.matrix(1), 10, 1, nthreads=NT)
    assert (s == ref).all()
