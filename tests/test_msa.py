"""Progressive multiple alignment along the guide tree (tweakseq_b200/csrc/msa.cuh, msa_host.h).

CPU tier: the oracle (tsq_oracle_msa) against an independent pure-Python statement of the spec and
against tsq_oracle_traceback for n = 2; then the product's own planning code and kernel phase functions,
run thread by thread on the CPU by tests/msa_emul.cpp, against the oracle.
GPU tier: tsq_msa / tsq_run_fasta through the C ABI against the oracle, byte for byte."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import pyoracle as o
import np_msa

HERE = os.path.dirname(os.path.abspath(__file__))
PROT = "ARNDCQEGHILKMFPSTWYVBZX"
NUC = "ACGTN"


def smat_dict(alphabet):
    letters = NUC if alphabet == o.NUCLEOTIDE else PROT
    m = o.matrix(alphabet)
    return {a: {b: int(m[i, j]) for j, b in enumerate(letters)} for i, a in enumerate(letters)}


def family(rng, n, length, alphabet=o.PROTEIN, mut=0.25, indel=0.08):
    letters = "ACGT" if alphabet == o.NUCLEOTIDE else PROT[:20]
    root = rng.choice(list(letters), size=length)
    out = []
    for _ in range(n):
        s = []
        for ch in root:
            u = rng.random()
            if u < indel / 2:
                continue
            if u < indel:
                s.append(rng.choice(list(letters)))
            s.append(rng.choice(list(letters)) if rng.random() < mut else ch)
        out.append("".join(s))
    return out


def random_tree(rng, n):
    """A random binary merge order (not necessarily UPGMA): exercises arbitrary level structures."""
    alive = list(range(n))
    left, right = [], []
    for t in range(n - 1):
        i, j = sorted(rng.choice(len(alive), size=2, replace=False))
        a, b = alive[i], alive[j]
        if rng.random() < 0.5:
            a, b = b, a
        left.append(a); right.append(b)
        alive = [x for k, x in enumerate(alive) if k not in (i, j)] + [n + t]
    return np.array(left, np.uint32), np.array(right, np.uint32)


def caterpillar(n):
    left = [0] + [n + t - 1 for t in range(1, n - 1)]
    right = list(range(1, n))
    return np.array(left, np.uint32), np.array(right, np.uint32)


def upgma_tree(seqs, alphabet, go, ge):
    enc = [o.encode(s, alphabet) for s in seqs]
    mat = o.matrix(alphabet)
    sc, _ = o.all_pairs(enc, mat, go, ge)
    selfs = np.array([o.self_score(e, mat) for e in enc], dtype=np.int32)
    left, right, _ = o.upgma(o.distances(sc, selfs), len(seqs))
    return left, right


def check_rows(rows, seqs, alphabet):
    """Properties every multiple alignment must have, whatever the scores."""
    canon = lambda s: "".join((NUC if alphabet == o.NUCLEOTIDE else PROT)[v] for v in o.encode(s, alphabet))
    assert len(rows) == len(seqs)
    assert len({len(r) for r in rows}) <= 1
    for r, s in zip(rows, seqs):
        assert r.replace("-", "") == canon(s)
    if rows and len(seqs) > 0:
        for c in range(len(rows[0])):
            assert any(r[c] != "-" for r in rows), f"column {c} is all gaps"


# ---------------------------------------------------------------- oracle -------------------------

def test_oracle_two_sequences_is_the_pairwise_traceback():
    rng = np.random.default_rng(5)
    mat = o.matrix(o.PROTEIN)
    for _ in range(40):
        a, b = family(rng, 2, int(rng.integers(0, 40)), mut=0.4, indel=0.2)
        ea, eb = o.encode(a), o.encode(b)
        go, ge = int(rng.integers(0, 14)), int(rng.integers(0, 4))
        rows, sc = o.msa([ea, eb], mat, go, ge, [0], [1])
        ra, rb, s = o.traceback(ea, eb, mat, go, ge)
        assert (rows[0], rows[1], int(sc[0])) == (ra, rb, s)
        assert s == o.gotoh(ea, eb, mat, go, ge)


@pytest.mark.parametrize("alphabet", [o.PROTEIN, o.NUCLEOTIDE])
def test_oracle_equals_independent_python_statement(alphabet):
    rng = np.random.default_rng(11 + alphabet)
    S = smat_dict(alphabet)
    mat = o.matrix(alphabet)
    letters = NUC if alphabet == o.NUCLEOTIDE else PROT
    for trial in range(30):
        n = int(rng.integers(1, 7))
        seqs = family(rng, n, int(rng.integers(0, 14)), alphabet, mut=0.35, indel=0.25)
        if trial % 5 == 0 and n > 1:
            seqs[int(rng.integers(0, n))] = ""
        go, ge = int(rng.integers(0, 13)), int(rng.integers(0, 3))
        left, right = random_tree(rng, n) if n > 1 else (np.zeros(0, np.uint32),) * 2
        enc = [o.encode(s, alphabet) for s in seqs]
        rows, sc = o.msa(enc, mat, go, ge, left, right, alphabet)
        canon = ["".join(letters[v] for v in e) for e in enc]
        prow, psc = np_msa.progressive(canon, left.tolist(), right.tolist(), S, go, ge)
        assert rows == prow
        assert sc.tolist() == psc
        check_rows(rows, seqs, alphabet)


def test_oracle_identical_sequences_align_without_gaps():
    s = "MKTAYIAKQRQISFVKSHFSRQLEERLGLIEVQ"
    enc = [o.encode(s)] * 5
    left, right = caterpillar(5)
    rows, sc = o.msa(enc, o.matrix(0), 11, 1, left, right)
    assert rows == [s] * 5
    selfs = o.self_score(enc[0], o.matrix(0))
    assert sc.tolist() == [selfs * k for k in (1, 2, 3, 4)]   # |X| |Y| pair sums, no gap


# ---------------------------------------------------------------- emulator -----------------------

@pytest.fixture(scope="module")
def emul():
    so = os.path.join(HERE, "_msa_emul.so")
    src = [os.path.join(HERE, "msa_emul.cpp"), os.path.join(HERE, "..", "tweakseq_b200", "csrc", "msa.cuh"),
           os.path.join(HERE, "..", "tweakseq_b200", "csrc", "msa_host.h")]
    if not os.path.exists(so) or any(os.path.getmtime(so) < os.path.getmtime(f) for f in src):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wall", "-Wextra", "-shared", "-fPIC", "-o", so, src[0]])
    L = C.CDLL(so)
    u8p, u32p, u64p = (C.POINTER(t) for t in (C.c_uint8, C.c_uint32, C.c_uint64))
    L.msa_emul.argtypes = [u8p, u64p, u32p, C.c_uint32, C.POINTER(C.c_int8), C.c_int, C.c_int, C.c_int, u32p, u32p,
                           C.c_char_p, u8p, C.c_uint64, u32p, C.POINTER(C.c_longlong), u32p, C.c_uint64, C.c_uint32,
                           C.c_int, u32p, u32p]

    def run(enc, mat, go, ge, left, right, alphabet=o.PROTEIN, budget=0, threads=0, ascending=0):
        n = len(enc)
        flat, offs, lens = o._pack(enc)
        m8 = np.ascontiguousarray(mat, dtype=np.int8)
        left = np.ascontiguousarray(left, dtype=np.uint32) if n > 1 else np.zeros(1, np.uint32)
        right = np.ascontiguousarray(right, dtype=np.uint32) if n > 1 else np.zeros(1, np.uint32)
        cap = n * (int(lens.sum()) + 1)
        rows = np.zeros(cap, np.uint8)
        sc = np.zeros(max(n, 1), np.int64)
        order = np.zeros(max(n, 1), np.uint32)
        ncols, launches, levels = C.c_uint32(), C.c_uint32(), C.c_uint32()
        p = lambda a, t: a.ctypes.data_as(C.POINTER(t))
        letters = (NUC if alphabet == o.NUCLEOTIDE else PROT).encode()
        rc = L.msa_emul(p(flat, C.c_uint8), p(offs, C.c_uint64), p(lens, C.c_uint32), n, p(m8, C.c_int8), m8.shape[0],
                        go, ge, p(left, C.c_uint32), p(right, C.c_uint32), letters, p(rows, C.c_uint8), cap,
                        C.byref(ncols), p(sc, C.c_longlong), p(order, C.c_uint32), budget, threads, ascending,
                        C.byref(launches), C.byref(levels))
        if rc == 4:
            return "cancelled"
        if rc == 3:
            return "bad tree"
        assert rc == 0, rc
        k = ncols.value
        out = [rows[r * k:(r + 1) * k].tobytes().decode("ascii") for r in range(n)]
        return out, sc[:max(n - 1, 0)].copy(), order[:n].copy(), launches.value, levels.value
    return run


@pytest.mark.parametrize("alphabet", [o.PROTEIN, o.NUCLEOTIDE])
def test_kernel_phases_on_cpu_equal_oracle_random_trees(emul, alphabet):
    rng = np.random.default_rng(101 + alphabet)
    mat = o.matrix(alphabet)
    for trial in range(40):
        n = int(rng.integers(1, 14))
        seqs = family(rng, n, int(rng.integers(0, 70)), alphabet, mut=0.3, indel=0.15)
        if trial % 4 == 0 and n > 2:
            seqs[int(rng.integers(0, n))] = ""
        go, ge = int(rng.integers(0, 14)), int(rng.integers(0, 4))
        left, right = random_tree(rng, n) if n > 1 else (np.zeros(0, np.uint32),) * 2
        enc = [o.encode(s, alphabet) for s in seqs]
        want, wsc = o.msa(enc, mat, go, ge, left, right, alphabet)
        # bit 16: rolling diagonals in global scratch; bit 17: int64 sweep even where int32 would do
        threads = [0, 1, 32, 96, 1024][trial % 5] | ((trial // 5) & 1) << 16 | ((trial // 3) & 1) << 17
        budget = [0, 1, 1 << 16][trial % 3]   # 1 byte: one merge per launch
        got, gsc, order, launches, levels = emul(enc, mat, go, ge, left, right, alphabet, budget, threads, trial & 1)
        assert got == want
        assert gsc.tolist() == wsc.tolist()
        assert sorted(order.tolist()) == list(range(n))
        assert launches >= levels + 1
        check_rows(got, seqs, alphabet)


def test_kernel_phases_on_cpu_upgma_family_and_caterpillar(emul):
    rng = np.random.default_rng(7)
    mat = o.matrix(o.PROTEIN)
    seqs = family(rng, 24, 120, mut=0.2, indel=0.06)
    enc = [o.encode(s) for s in seqs]
    for left, right in (upgma_tree(seqs, o.PROTEIN, 11, 1), caterpillar(len(seqs))):
        want, wsc = o.msa(enc, mat, 11, 1, left, right)
        got, gsc, order, launches, levels = emul(enc, mat, 11, 1, left, right)
        assert got == want and gsc.tolist() == wsc.tolist()
        check_rows(got, seqs, o.PROTEIN)
    # leaves left to right of the caterpillar ((((0,1),2),3)...) are 0, 1, 2, ...
    assert order.tolist() == list(range(len(seqs)))
    assert levels == len(seqs) - 1


def test_kernel_phases_on_cpu_custom_matrix_and_wide_counts(emul):
    rng = np.random.default_rng(9)
    m = rng.integers(-8, 12, size=(23, 23))
    m = ((m + m.T) // 2).astype(np.int8)
    seqs = family(rng, 40, 30, mut=0.5, indel=0.2)
    enc = [o.encode(s) for s in seqs]
    left, right = random_tree(rng, len(seqs))
    want, wsc = o.msa(enc, m, 3, 2, left, right)
    got, gsc, *_ = emul(enc, m, 3, 2, left, right)
    assert got == want and gsc.tolist() == wsc.tolist()


# ---------------------------------------------------------------- GPU, through the C ABI ---------

def _gpu_msa(seqs, alphabet=o.PROTEIN, go=-1, ge=-1, matrix=None, flags=0):
    """(rows, tree order, merges) from libtsqb200.so."""
    import tweakseq_b200 as t
    with t.Context(alphabet=alphabet, gap_open=go, gap_extend=ge, matrix=matrix, flags=flags) as ctx:
        ctx.set_sequences(seqs)
        ctx.run()
        left, right, _ = ctx.guide_tree()
        rows, order = ctx.msa()
        again, _ = ctx.msa()          # cached result, same answer
        assert again == rows
        st = ctx.stats()
    return rows, order, left, right, st


def _tree_order(left, right, n):
    if n == 0:
        return []
    out, st = [], [0 if n == 1 else 2 * n - 2]
    while st:
        i = st.pop()
        if i < n:
            out.append(i)
        else:
            st.append(int(right[i - n])); st.append(int(left[i - n]))
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("n,length,alphabet", [(1, 30, 0), (2, 50, 0), (3, 40, 0), (17, 90, 0), (64, 150, 0),
                                               (33, 200, 1), (150, 60, 0)])
def test_gpu_msa_equals_oracle(n, length, alphabet):
    rng = np.random.default_rng(1000 + n)
    seqs = family(rng, n, length, alphabet, mut=0.25, indel=0.08)
    if n >= 17:
        seqs[5] = ""                                      # an empty sequence becomes an all-gap row
        seqs[7] = seqs[7][:3]
    go, ge = (11, 1) if alphabet == o.PROTEIN else (10, 1)
    rows, order, left, right, st = _gpu_msa(seqs, alphabet)
    enc = [o.encode(s, alphabet) for s in seqs]
    want, _ = o.msa(enc, o.matrix(alphabet), go, ge, left, right, alphabet)
    assert rows == want
    assert order == _tree_order(left, right, n)
    check_rows(rows, seqs, alphabet)
    assert st["msa_ms"] > 0 or n == 0


@pytest.mark.gpu
def test_gpu_msa_of_two_sequences_is_align_pair():
    import tweakseq_b200 as t
    rng = np.random.default_rng(77)
    a, b = family(rng, 2, 300, mut=0.3, indel=0.1)
    with t.Context() as ctx:
        ctx.set_sequences([a, b])
        ctx.run()
        rows, order = ctx.msa()
        ra, rb, _ = ctx.align_pair(0, 1)
    assert rows == [ra, rb] and order == [0, 1]


@pytest.mark.gpu
def test_gpu_msa_custom_matrix_gaps_and_identity_tree():
    import tweakseq_b200 as t
    rng = np.random.default_rng(78)
    m = rng.integers(-6, 10, size=(23, 23))
    m = ((m + m.T) // 2).astype(np.int8)
    seqs = family(rng, 40, 80, mut=0.4, indel=0.15)
    for flags in (0, t.FLAG_IDENTITY):
        rows, order, left, right, _ = _gpu_msa(seqs, go=4, ge=2, matrix=m, flags=flags)
        want, _ = o.msa([o.encode(s) for s in seqs], m, 4, 2, left, right)
        assert rows == want


@pytest.mark.gpu
def test_gpu_msa_long_profiles_span_several_sweeps():
    # diagonals longer than a CTA (1 024 threads): every thread takes several cells per diagonal
    rng = np.random.default_rng(79)
    seqs = family(rng, 6, 2600, o.NUCLEOTIDE, mut=0.1, indel=0.03)
    rows, order, left, right, _ = _gpu_msa(seqs, o.NUCLEOTIDE)
    want, _ = o.msa([o.encode(s, o.NUCLEOTIDE) for s in seqs], o.matrix(o.NUCLEOTIDE), 10, 1, left, right, o.NUCLEOTIDE)
    assert rows == want


@pytest.mark.gpu
def test_gpu_run_fasta_writes_the_alignment_readNewAlignment_expects(tmp_path):
    import tweakseq_b200 as t
    from tweakseq_b200.fasta import read_fasta
    rng = np.random.default_rng(80)
    seqs = family(rng, 20, 130, mut=0.2, indel=0.06)
    seqs[4] = seqs[4].lower()                              # the input's own spelling comes back
    seqs[6] = seqs[6][:40] + "JOU" + seqs[6][40:]          # letters the matrix folds into X stay as typed
    labels = [f"seq{k}" for k in range(len(seqs))]
    fin, fout = str(tmp_path / "in.fa"), str(tmp_path / "out.fa")
    with open(fin, "w") as f:
        for l, s in zip(labels, seqs):
            f.write(f">{l} some description\n")
            for at in range(0, len(s), 70):
                f.write(s[at:at + 70] + "\n")
    tool = t.B200Gotoh()
    assert tool.align                                      # the default: run() hands back what the editor ingests
    log = []
    assert tool.run(fin, fout, log=log.append) == 0
    assert any("progressive alignment" in m for m in log)
    got_labels, got_rows, _ = read_fasta(fout)
    # rows come in tree order; every label of the input is there exactly once (Project.cpp:908-1032 matches by label)
    assert sorted(got_labels) == sorted(labels)
    by_label = dict(zip(got_labels, got_rows))
    assert len({len(r) for r in got_rows}) == 1
    for l, s in zip(labels, seqs):
        assert by_label[l].replace("-", "") == s
    # the same rows, canonical spelling, from the in-memory API and from the oracle
    rows, order, left, right, _ = _gpu_msa(seqs)
    assert [got_labels.index(labels[r]) for r in order] == list(range(len(seqs)))
    for l, r in zip(labels, rows):
        assert len(by_label[l]) == len(r) and all((a == "-") == (b == "-") for a, b in zip(by_label[l], r))
    want, _ = o.msa([o.encode(s) for s in seqs], o.matrix(0), 11, 1, left, right)
    assert rows == want
    # nothing else appears next to the editor's temporary file: tree and matrix (n^2 numbers of text) on request only
    assert not os.path.exists(fout + ".dnd") and not os.path.exists(fout + ".distmat")
    tool.keep_distmat = tool.keep_tree = True
    assert tool.run(fin, fout) == 0
    assert os.path.exists(fout + ".dnd")
    from tweakseq_b200.fasta import read_distmat
    lab, mat = read_distmat(fout + ".distmat")
    assert lab == labels and len(mat) == len(seqs)


def test_kernel_phases_on_cpu_degenerate_inputs(emul):
    mat = o.matrix(o.PROTEIN)
    none = np.zeros(0, np.uint32)
    assert emul([], mat, 11, 1, none, none)[0] == []
    empties = [o.encode(""), o.encode("--"), o.encode(" ")]
    rows, sc, order, _, _ = emul(empties, mat, 11, 1, [0, 3], [1, 2])
    assert rows == ["", "", ""] and sc.tolist() == [0, 0] and order.tolist() == [0, 1, 2]
    one = [o.encode("MKV")]
    assert emul(one, mat, 11, 1, none, none)[0] == ["MKV"]
    # an empty cluster against a real one: every column is a gap column of the empty side
    rows, sc, _, _, _ = emul([o.encode(""), o.encode("WW")], mat, 11, 1, [0], [1])
    assert rows == ["--", "WW"] and sc.tolist() == [-(11 + 2)]
    assert o.msa([o.encode(""), o.encode("WW")], mat, 11, 1, [0], [1])[0] == ["--", "WW"]


def balanced(n):
    """Pairs neighbours level by level: ((0,1),(2,3)),... -- the widest clusters meet at the root."""
    left, right, alive, nxt = [], [], list(range(n)), n
    while len(alive) > 1:
        new = []
        for a in range(0, len(alive) - 1, 2):
            left.append(alive[a]); right.append(alive[a + 1]); new.append(nxt); nxt += 1
        if len(alive) & 1:
            new.append(alive[-1])
        alive = new
    return np.array(left, np.uint32), np.array(right, np.uint32)


def test_int32_sweep_is_used_only_inside_its_range_bound(emul):
    """Boundary cells reach -|X||Y| (go + L ge): with the largest gap costs the ABI admits, a 47 x 47 root
    merge of ~100-column clusters sits just inside the +-2^29 bound of the int32 sweep (msa_fits_narrow) and
    a 256 x 256 one would wrap 32 bits.  The oracle is int64 throughout, so a wrapped int32 anywhere shows
    as different rows or scores."""
    rng = np.random.default_rng(21)
    mat = o.matrix(o.PROTEIN)
    assert 47 * 47 * (4096 + 95 * 1024) > (1 << 27) and 256 * 256 * (4096 + 95 * 1024) > (1 << 31)
    for n in (94, 512):
        seqs = ["".join(rng.choice(list(PROT[:20]), int(rng.integers(95, 100)))) for _ in range(n)]
        enc = [o.encode(s) for s in seqs]
        left, right = balanced(n)
        want, wsc = o.msa(enc, mat, 4096, 1024, left, right)
        got, gsc, *_ = emul(enc, mat, 4096, 1024, left, right)
        assert got == want and gsc.tolist() == wsc.tolist()


@pytest.mark.gpu
def test_gpu_msa_int32_and_int64_sweeps_at_the_range_boundary():
    rng = np.random.default_rng(22)
    for n in (94, 512):
        seqs = ["".join(rng.choice(list(PROT[:20]), int(rng.integers(95, 100)))) for _ in range(n)]
        rows, order, left, right, _ = _gpu_msa(seqs, go=4096, ge=1024)
        want, wsc = o.msa([o.encode(s) for s in seqs], o.matrix(0), 4096, 1024, left, right)
        assert rows == want


# ---------------------------------------------------------------- committed golden vectors --------

def _golden():
    import json
    return json.load(open(os.path.join(HERE, "golden", "msa_small.json")))


def test_golden_alignments_oracle_and_kernel_phases(emul):
    for g in _golden():
        enc = [o.encode(s, g["alphabet"]) for s in g["seqs"]]
        mat = o.matrix(g["alphabet"])
        rows, sc = o.msa(enc, mat, g["go"], g["ge"], g["left"], g["right"], g["alphabet"])
        assert rows == g["rows"] and sc.tolist() == g["merge_scores"]
        rows, sc, *_ = emul(enc, mat, g["go"], g["ge"], g["left"], g["right"], g["alphabet"])
        assert rows == g["rows"] and sc.tolist() == g["merge_scores"]


@pytest.mark.gpu
def test_gpu_golden_alignments():
    for g in _golden():
        rows, order, left, right, _ = _gpu_msa(g["seqs"], g["alphabet"], g["go"], g["ge"])
        assert left.tolist() == g["left"] and right.tolist() == g["right"]     # the same tree ...
        assert rows == g["rows"]                                               # ... and the same rows


def test_plan_stops_at_the_cancel_flag(emul):
    enc = [o.encode("MKTAYIAK"), o.encode("MKTAIAK"), o.encode("MKAYIAK")]
    assert emul(enc, o.matrix(0), 11, 1, [0, 3], [1, 2], threads=1 << 18) == "cancelled"


def test_plan_rejects_merges_that_are_not_a_tree(emul):
    enc = [o.encode(x) for x in ("MKTAYIAK", "MKTAIAK", "MKAYIAK", "MKAYIA")]
    mat = o.matrix(0)
    assert emul(enc, mat, 11, 1, [0, 0, 4], [1, 2, 5]) == "bad tree"        # leaf 0 merged twice
    assert emul(enc, mat, 11, 1, [0, 5, 4], [1, 2, 3]) == "bad tree"        # node 5 used before it exists
    assert emul(enc, mat, 11, 1, [0, 2, 4], [0, 3, 5]) == "bad tree"        # a node merged with itself
    assert emul(enc, mat, 11, 1, [0, 2, 4], [1, 3, 5])[0][0].replace("-", "") == "MKTAYIAK"


# ---------------------------------------------------------------- the definition, by enumeration ---

def _profile_paths(m, q):
    """Every alignment of m columns against q columns as moves 'M' (both), 'D' (X only), 'I' (Y only)."""
    out = []

    def rec(i, j, acc):
        if i == m and j == q:
            out.append(acc)
            return
        if i < m and j < q:
            rec(i + 1, j + 1, acc + "M")
        if i < m:
            rec(i + 1, j, acc + "D")
        if j < q:
            rec(i, j + 1, acc + "I")

    rec(0, 0, "")
    return out


def test_every_merge_score_is_the_best_over_all_column_alignments():
    """Pins the merge recurrence to its DEFINITION rather than to another DP: for tiny clusters every
    alignment of the two column sets is enumerated and scored directly -- residue pairs across the two
    clusters summed column by column, every maximal run of k columns taken from one cluster only costs
    |X||Y| (go + k ge) -- and the best of them must be the score the oracle reports for that merge."""
    rng = np.random.default_rng(314)
    S = smat_dict(o.PROTEIN)
    mat = o.matrix(o.PROTEIN)
    for trial in range(25):
        n = int(rng.integers(2, 5))
        seqs = family(rng, n, int(rng.integers(1, 5)), mut=0.4, indel=0.3)
        seqs = [s[:4] if s else PROT[int(rng.integers(0, 20))] for s in seqs]
        go, ge = int(rng.integers(0, 13)), int(rng.integers(0, 3))
        left, right = random_tree(rng, n)
        enc = [o.encode(s) for s in seqs]
        _, scores = o.msa(enc, mat, go, ge, left, right)
        # rebuild every intermediate cluster with the independent Python statement to get its rows
        canon = ["".join(PROT[v] for v in e) for e in enc]
        rows = {r: [canon[r]] for r in range(n)}
        for t, (l, r) in enumerate(zip(left.tolist(), right.tolist())):
            X, Y = rows[l], rows[r]
            m, q, w = len(X[0]), len(Y[0]), len(X) * len(Y)
            best = None
            for path in _profile_paths(m, q):
                i = j = 0
                total, prev = 0, ""
                for mv in path:
                    if mv == "M":
                        total += np_msa.sub_score([x[i] for x in X], [y[j] for y in Y], S)
                        i += 1; j += 1
                    else:
                        total -= w * (ge + (go if mv != prev else 0))
                        if mv == "D":
                            i += 1
                        else:
                            j += 1
                    prev = mv
                best = total if best is None or total > best else best
            assert int(scores[t]) == best, (trial, t, seqs)
            rows[n + t], _ = np_msa.align_profiles(X, Y, S, go, ge)


@pytest.mark.gpu
def test_gpu_msa_profiles_beyond_shared_memory():
    """7 rolling int32 diagonals of a 7 600-column profile need 213 KB: past the 200 KB the plan grants a CTA,
    so this merge sweeps through its global (L2) scratch -- the one path the smaller tests never take."""
    rng = np.random.default_rng(81)
    seqs = family(rng, 3, 7600, o.NUCLEOTIDE, mut=0.05, indel=0.01)
    assert min(len(s) for s in seqs) > 7400
    rows, order, left, right, _ = _gpu_msa(seqs, o.NUCLEOTIDE)
    want, _ = o.msa([o.encode(s, o.NUCLEOTIDE) for s in seqs], o.matrix(o.NUCLEOTIDE), 10, 1, left, right, o.NUCLEOTIDE)
    assert rows == want
