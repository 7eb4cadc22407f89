"""identity-aware scoring.  CPU leg pins the oracle's rule (best score first, then
most identities) against a brute-force enumeration of every alignment of tiny sequences; the gpu leg
checks the CUDA path (32-bit inter-task kernel and 32-bit wavefront kernel) against the oracle."""
import itertools

import numpy as np
import pytest

from oracle import pyoracle as o

MAT = o.matrix(0)


def brute_force(a, b, mat, go, ge):
    """Best (score, identities) over ALL global alignments, lexicographic."""
    best = None
    m, n = len(a), len(b)

    def rec(i, j, score, nid, state):       # state: 0 none/diag, 1 gap in a (consuming b), 2 gap in b
        nonlocal best
        if i == m and j == n:
            cand = (score, nid)
            if best is None or cand > best:
                best = cand
            return
        if i < m and j < n:
            rec(i + 1, j + 1, score + int(mat[a[i]][b[j]]), nid + (a[i] == b[j]), 0)
        if j < n:
            rec(i, j + 1, score - ge - (go if state != 1 else 0), nid, 1)
        if i < m:
            rec(i + 1, j, score - ge - (go if state != 2 else 0), nid, 2)
    rec(0, 0, 0, 0, 0)
    return best


def test_identity_oracle_matches_brute_force_on_tiny_inputs():
    rng = np.random.default_rng(8)
    for _ in range(250):
        a = rng.integers(0, 6, rng.integers(1, 6)).tolist()      # small alphabet: many ties and identities
        b = rng.integers(0, 6, rng.integers(1, 6)).tolist()
        go, ge = int(rng.integers(0, 4)), int(rng.integers(0, 3))
        s, k = o.gotoh_id(np.array(a), np.array(b), MAT, go, ge)
        assert (s, k) == brute_force(a, b, MAT.tolist(), go, ge), (a, b, go, ge)
        assert s == o.gotoh(np.array(a, np.uint8), np.array(b, np.uint8), MAT, go, ge)     # the score itself is unchanged


def test_identity_oracle_edge_cases():
    assert o.gotoh_id(o.encode("ACDEF"), o.encode("ACDEF"), MAT, 11, 1) == (o.score_str("ACDEF", "ACDEF"), 5)
    assert o.gotoh_id(o.encode(""), o.encode("ACD"), MAT, 11, 1) == (-14, 0)
    assert o.gotoh_id(o.encode("WWWW"), o.encode("WWCWW"), MAT, 11, 1)[1] == 4
    # co-optimal alignments with different identity counts: go = ge = 0 makes many paths tie
    s, k = o.gotoh_id(o.encode("AXA"), o.encode("AA"), MAT, 0, 0)
    assert (s, k) == (8, 2)


@pytest.mark.gpu
@pytest.mark.parametrize("flags_extra", [0, 1], ids=["gotoh32", "wave32"])      # 1 = TSQ_FLAG_FORCE_S32
def test_gpu_identity_mode_matches_oracle(flags_extra):
    import tweakseq_b200 as t
    from tweakseq_b200 import synth
    rng = np.random.default_rng(70 + flags_extra)
    fam = synth.protein(18, (60, 200, 120, 40), 1, family=True)
    rag = ["".join(rng.choice(list("ARNDCQEGHILKMFPSTWYVBZX"), int(l))) for l in rng.integers(0, 180, 30)]
    seqs = fam + rag + ["", "A", "acdefg"]
    for go, ge in ((11, 1), (0, 0), (4, 3)):
        with t.Context(gap_open=go, gap_extend=ge, flags=t.FLAG_IDENTITY | flags_extra) as ctx:
            ctx.set_sequences(seqs)
            ctx.run()
            s, k, d, st = ctx.scores(), ctx.identities(), ctx.distances(), ctx.stats()
        enc = [o.encode(x) for x in seqs]
        rs, rk, rd = o.all_pairs_id(enc, MAT, go, ge)
        assert (s == rs).all() and (k == rk).all()
        assert d.tobytes() == rd.tobytes()
        assert st["cells_s16"] == 0 and st["cells_s32"] > 0
        plain, _ = o.all_pairs(enc, MAT, go, ge, nthreads=4)
        assert (s == plain).all()              # identity mode never changes the scores


@pytest.mark.gpu
def test_gpu_identity_mode_nucleotide_and_long():
    import tweakseq_b200 as t
    rng = np.random.default_rng(72)
    base = "".join(rng.choice(list("ACGT"), 9000))
    seqs = [base, base[:4000] + base[4100:], "".join(rng.choice(list("ACGT"), 8500)), "ACGT" * 50, "ACGTN"]
    with t.Context(alphabet=1, flags=t.FLAG_IDENTITY) as ctx:      # 9 kb > 8192: wavefront path with wide keys
        ctx.set_sequences(seqs)
        ctx.run()
        s, k = ctx.scores(), ctx.identities()
    enc = [o.encode(x, 1) for x in seqs]
    rs, rk, _ = o.all_pairs_id(enc, o.matrix(1), 10, 1)
    assert (s == rs).all() and (k == rk).all()


@pytest.mark.gpu
def test_wide_gap_penalties_use_the_32_bit_inter_task_kernel():
    import tweakseq_b200 as t
    rng = np.random.default_rng(73)
    seqs = ["".join(rng.choice(list("ARNDCQEGHILKMFPSTWYV"), int(l))) for l in rng.integers(1, 150, 60)]
    with t.Context(gap_open=4000, gap_extend=900) as ctx:           # no room for 16 bits, sequences short
        ctx.set_sequences(seqs)
        ctx.run()
        s, st = ctx.scores(), ctx.stats()
    enc = [o.encode(x) for x in seqs]
    ref, cells = o.all_pairs(enc, MAT, 4000, 900, nthreads=4)
    assert (s == ref).all() and st["cells_s32"] == cells and st["cells_s16"] == 0


@pytest.mark.gpu
@pytest.mark.parametrize("K", [30, 32, 36, 40, 44, 48, 50, 52, 56])
def test_every_strip_width_of_the_32_bit_inter_task_kernel(K, monkeypatch):
    """gotoh32_kernel is instantiated for the same strip widths as the packed kernel; identity mode
    routes every short sequence through it."""
    import tweakseq_b200 as t
    monkeypatch.setenv("TSQ_FORCE_K", str(K))
    rng = np.random.default_rng(500 + K)
    lens = list(rng.integers(1, 3 * K + 7, 40)) + [K - 1, K, K + 1, 2 * K, 2 * K + 1]
    seqs = ["".join(rng.choice(list("ARNDCQEGHILKMFPSTWYV"), int(l))) for l in lens]
    with t.Context(flags=t.FLAG_IDENTITY) as ctx:
        ctx.set_sequences(seqs)
        ctx.run()
        s, k, st = ctx.scores(), ctx.identities(), ctx.stats()
    assert st["strip_width"] == K and st["cells_s16"] == 0
    rs, rk, _ = o.all_pairs_id([o.encode(x) for x in seqs], MAT, 11, 1)
    assert (s == rs).all() and (k == rk).all()


# ---- Kimura-corrected identity distance (SURVEY 8f-2 "+ Kimura") ---------------------------------------------------
def test_the_specs_logarithm_is_a_logarithm():
    """tsq_oracle_ln: the fixed sequence of IEEE operations both sides evaluate; within 2 ulp of libm on the range the
    correction uses (arguments 0.1375 .. 1) and beyond."""
    import math
    rng = np.random.default_rng(5)
    xs = np.concatenate([np.linspace(0.1375, 1.0, 5001), rng.uniform(1e-3, 1.0, 5000), rng.uniform(1.0, 1e6, 2000)])
    for x in xs:
        a, b = o.ln(float(x)), math.log(float(x))
        assert abs(a - b) <= 2 * np.spacing(abs(b)) + 1e-300, x
    assert o.ln(1.0) == 0.0
    d, ok = o.kimura(100, 100)
    assert ok and d == 0.0
    d, ok = o.kimura(60, 100)                          # D = 0.4: -ln(1 - 0.4 - 0.032)
    assert ok and abs(d - (-math.log(1 - 0.4 - 0.4 * 0.4 / 5))) < 1e-15
    assert o.kimura(25, 100) == (0.75, False) and o.kimura(26, 100)[1] and not o.kimura(0, 0)[1]


@pytest.mark.gpu
def test_gpu_kimura_distances_equal_the_oracles_bit_for_bit():
    import tweakseq_b200 as t
    from tweakseq_b200 import synth
    fam = [s for s in synth.protein(60, (150, 260, 200, 25), 11, family=True)]
    enc = [o.encode(s) for s in fam]
    mat = o.matrix(0)
    # a family mutated at 10-60 %: keep the members whose every pair stays below D = 0.75
    with t.Context(flags=t.FLAG_IDENTITY) as ctx:
        ctx.set_sequences(fam)
        ctx.run()
        nid, d0 = ctx.identities(), ctx.distances()
    n = len(fam)
    keep = [i for i in range(n) if all(d0[t.pair_index(min(i, j), max(i, j), n)] < 0.7 for j in range(n) if j != i)][:24]
    if len(keep) < 8:                                  # fall back to close copies of one member
        keep = list(range(8))
        fam = [fam[0][:k] + fam[0][k + 1:] for k in range(3, 83, 10)]
    else:
        fam = [fam[i] for i in keep]
    with t.Context(flags=t.FLAG_IDENTITY | t.FLAG_KIMURA) as ctx:
        ctx.set_sequences(fam)
        ctx.run()
        s, nid, d = ctx.scores(), ctx.identities(), ctx.distances()
    m = len(fam)
    want = []
    for i in range(m):
        for j in range(i + 1, m):
            k = nid[t.pair_index(i, j, m)]
            dd, ok = o.kimura(int(k), min(len(fam[i]), len(fam[j])))
            assert ok
            want.append(dd)
    assert d.tobytes() == np.array(want, dtype=np.float64).tobytes()
    assert (d >= 0).all() and d.max() > 0.05


@pytest.mark.gpu
def test_gpu_kimura_beyond_its_range_is_an_error_not_a_guess():
    import tweakseq_b200 as t
    from tweakseq_b200 import synth
    seqs = synth.protein(12, 120, 12)                   # unrelated: D ~ 0.9
    with t.Context(flags=t.FLAG_IDENTITY | t.FLAG_KIMURA) as ctx:
        ctx.set_sequences(seqs)
        with pytest.raises(t.TsqError) as e:
            ctx.run()
        assert e.value.status == -9 and "0.75" in str(e.value)
        with pytest.raises(t.TsqError):
            ctx.distances()
    with pytest.raises(t.TsqError) as e:
        t.Context(flags=t.FLAG_KIMURA)                  # corrects the identity distance: needs TSQ_FLAG_IDENTITY
    assert e.value.status == -1
