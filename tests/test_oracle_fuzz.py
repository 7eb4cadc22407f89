"""Property-based pinning of the oracle (hypothesis): three independent formulations agree on
arbitrary small inputs, scores are symmetric, and the documented bounds hold."""
import numpy as np
from hypothesis import given, settings, strategies as st

from np_gotoh import gotoh_np, gotoh_py
from oracle import pyoracle as o

MAT = o.matrix(o.PROTEIN)
DNA = o.matrix(o.NUCLEOTIDE)
seq23 = st.lists(st.integers(0, 22), min_size=0, max_size=40)
seq5 = st.lists(st.integers(0, 4), min_size=0, max_size=60)
gaps = st.tuples(st.integers(0, 20), st.integers(0, 6))


@settings(max_examples=300, deadline=None)
@given(seq23, seq23, gaps)
def test_three_formulations_agree_protein(a, b, g):
    go, ge = g
    s = o.gotoh(np.array(a, np.uint8), np.array(b, np.uint8), MAT, go, ge)
    assert s == gotoh_np(a, b, MAT, go, ge) == gotoh_py(a, b, MAT.tolist(), go, ge)
    assert s == o.gotoh(np.array(b, np.uint8), np.array(a, np.uint8), MAT, go, ge)      # symmetric matrix
    if a and b:
        lo = -(2 * go + (len(a) + len(b)) * ge)       # all-gap path
        hi = 11 * min(len(a), len(b))                  # every column a W/W match
        assert lo <= s <= hi


@settings(max_examples=200, deadline=None)
@given(seq5, seq5, gaps)
def test_three_formulations_agree_nucleotide(a, b, g):
    go, ge = g
    s = o.gotoh(np.array(a, np.uint8), np.array(b, np.uint8), DNA, go, ge)
    assert s == gotoh_np(a, b, DNA, go, ge) == gotoh_py(a, b, DNA.tolist(), go, ge)


@settings(max_examples=100, deadline=None)
@given(st.lists(st.integers(0, 19), min_size=0, max_size=40), gaps)   # the 20 standard residues: S(a,a) >= S(a,b) > -inf
def test_self_alignment_is_the_diagonal_sum(a, g):
    go, ge = g
    arr = np.array(a, np.uint8)
    assert o.gotoh(arr, arr, MAT, go, ge) == (o.self_score(arr, MAT) if a else 0)
