"""Property-based pinning of the progressive alignment (hypothesis): on arbitrary small sequence sets, merge
orders and gap costs the oracle (tsq_oracle_msa), the independent Python statement (tests/np_msa.py) and the
product's own plan + kernel phase code run on the CPU (tests/msa_emul.cpp) produce the same rows and the same
score for every merge; the rows always are a valid alignment of the inputs."""
import numpy as np
from hypothesis import given, settings, strategies as st

import np_msa
from oracle import pyoracle as o
from test_msa import PROT, check_rows, emul, smat_dict  # noqa: F401  (emul: the module-scoped fixture)

MAT = o.matrix(o.PROTEIN)
S = smat_dict(o.PROTEIN)


@st.composite
def jobs(draw):
    n = draw(st.integers(1, 6))
    seqs = [draw(st.lists(st.integers(0, 22), min_size=0, max_size=9)) for _ in range(n)]
    # a merge order: repeatedly join two of the live nodes
    live, left, right = list(range(n)), [], []
    for t in range(n - 1):
        i = draw(st.integers(0, len(live) - 1))
        a = live.pop(i)
        j = draw(st.integers(0, len(live) - 1))
        b = live.pop(j)
        left.append(a); right.append(b); live.append(n + t)
    go, ge = draw(st.integers(0, 20)), draw(st.integers(0, 5))
    knobs = draw(st.integers(0, 7))
    return seqs, left, right, go, ge, knobs


@settings(max_examples=int(__import__("os").environ.get("TSQ_FUZZ_EXAMPLES", "150")), deadline=None)
@given(jobs())
def test_three_statements_of_the_alignment_agree(emul, job):
    seqs, left, right, go, ge, knobs = job
    enc = [np.array(s, np.uint8) for s in seqs]
    rows, sc = o.msa(enc, MAT, go, ge, left, right)
    canon = ["".join(PROT[v] for v in s) for s in seqs]
    prow, psc = np_msa.progressive(canon, left, right, S, go, ge)
    assert rows == prow and sc.tolist() == psc
    # bit 0: rolling diagonals in global scratch; bit 1: int64 sweep; bit 2: ascending thread order
    threads = [1, 32, 64, 1024][knobs & 3] | ((knobs & 1) << 16) | (((knobs >> 1) & 1) << 17)
    got, gsc, order, *_ = emul(enc, MAT, go, ge, left, right, threads=threads, ascending=(knobs >> 2) & 1)
    assert got == rows and gsc.tolist() == psc
    assert sorted(order.tolist()) == list(range(len(seqs)))
    check_rows(rows, canon, o.PROTEIN)
