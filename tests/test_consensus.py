"""consensus annotation.  The oracle restates Consensus::calculate
(tweakseq/Core/Annotations/Consensus.cpp:80-161) loop for loop; the CUDA path computes the same thing
from per-column class histograms and must agree character for character."""
import numpy as np
import pytest

from oracle import pyoracle as o


def test_consensus_oracle_known_columns():
    rows = ["AWC-a", "AWC-A", "AYD-A", "RWC-x"]
    # col0: A,A,A,R -> A (3 partners score > 0 for the first A: matches 2 >= 2)      -> 'A'
    # col1: W,W,Y,W -> first W: W/W 11, W/W 11, W/Y 2 -> matches 3                    -> 'W'
    # col2: C,C,D,C -> first C: matches 2 (C/D = -3)                                   -> 'C'
    # col3: all '-' -> class 99 everywhere: score 3, matches 3                          -> '-'
    # col4: a,A,A,x -> lower case is "not a residue" (no case folding in the reference):
    #        row 1 ('A'): A/A +4, two 99s -4 each = -4; row 0 ('a'): 99/99 +1, -4, -4 = -7  -> row 1 wins, matches 1 < 2 -> '?'
    assert o.consensus(rows) == "AWC-?"
    assert o.consensus(rows, plurality=1.0) == "AWC-A"
    assert o.consensus(["ACD"]) == "???"                      # one row: no partners, 0 < 0.5
    assert o.consensus(["ACD"], plurality=0.0) == "ACD"
    assert o.consensus([], None) == ""


def test_consensus_oracle_first_row_wins_ties():
    # I and V: I/V = 3 both ways, I/I = V/V = 4 -> rows 0 (I) and 1 (V) tie at 3; the first row is reported
    assert o.consensus(["I", "V"], plurality=1.0) == "I"
    assert o.consensus(["V", "I"], plurality=1.0) == "V"


def _random_alignment(rng, nrows, ncols):
    alphabet = list("ARNDCQEGHILKMFPSTWYVBZXJOU-") + list("acd.*")
    base = rng.choice(list("ARNDCQEGHILKMFPSTWYV"), ncols)
    rows = []
    for _ in range(nrows):
        r = base.copy()
        mut = rng.random(ncols) < 0.35
        r[mut] = rng.choice(alphabet, int(mut.sum()))
        rows.append("".join(r))
    return rows


@pytest.mark.gpu
@pytest.mark.parametrize("nrows,ncols", [(1, 7), (2, 33), (17, 64), (100, 301), (400, 95)])
def test_gpu_consensus_matches_the_reference_restatement(nrows, ncols):
    import tweakseq_b200 as t
    rng = np.random.default_rng(nrows * 1000 + ncols)
    rows = _random_alignment(rng, nrows, ncols)
    with t.Context() as ctx:
        assert ctx.consensus(rows) == o.consensus(rows)
        for pl in (0.0, 1.0, nrows * 0.8, nrows + 5.0):
            assert ctx.consensus(rows, plurality=pl) == o.consensus(rows, plurality=pl)


@pytest.mark.gpu
def test_gpu_consensus_edge_cases():
    import tweakseq_b200 as t
    with t.Context() as ctx:
        assert ctx.consensus([]) == ""
        assert ctx.consensus(["", ""]) == ""
        assert ctx.consensus(["----", "----", "AC-D"]) == o.consensus(["----", "----", "AC-D"])
        rows = ["AWC-a", "AWC-A", "AYD-A", "RWC-x"]
        assert ctx.consensus(rows) == "AWC-?"


# ---- the reference's OWN code (oracle/_ref/libref_consensus.so = tweakseq/Core/Annotations/Consensus.cpp
# compiled where it lies by oracle/Makefile) beside the restatement and beside the CUDA path ------------

needs_ref = pytest.mark.skipif(not o.ref_consensus_available(), reason="oracle/_ref/libref_consensus.so not built "
                                                                       "(needs /root/reference at build time)")


def _cells(rng, rows):
    """The rows as 16-bit residue cells with tweakseq's flag bits (Sequence.h:36-39) sprinkled in: bits >= 0x100
    vanish under Consensus.cpp's `& 0xff`; EXCLUDE_CELL (0x80) survives it and turns the cell into a non-residue."""
    out = []
    for r in rows:
        cells = []
        for ch in r:
            v = ord(ch)
            u = rng.random()
            if u < 0.10:
                v |= 0x0100            # HIGHLIGHT_CELL
            elif u < 0.15:
                v |= 0x0080            # EXCLUDE_CELL
            cells.append(v)
        out.append(cells)
    return out


@needs_ref
def test_restatement_equals_the_references_own_consensus_code():
    rng = np.random.default_rng(2026)
    assert o.ref_consensus(["AWC-a", "AWC-A", "AYD-A", "RWC-x"]) == "AWC-?"
    for trial in range(60):
        nrows, ncols = int(rng.integers(1, 40)), int(rng.integers(1, 80))
        rows = _random_alignment(rng, nrows, ncols)
        cells = _cells(rng, rows)
        seen = ["".join(chr(v & 0xff) for v in r) for r in cells]        # what `residues[c].unicode() & 0xff` yields
        assert o.ref_consensus(cells) == o.consensus(seen), (trial, nrows, ncols)       # default plurality rows / 2
        for pl in (0.0, 1.0, nrows * 0.8, nrows + 5.0):
            assert o.ref_consensus(cells, pl) == o.consensus(seen, plurality=pl)
