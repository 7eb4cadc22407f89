"""Parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle on the
same seeded inputs, against the committed golden vectors, and -- at BASELINE.json's full sizes --
through size-independent properties.  Bit-exact: integer scores identical, fp64 distances
identical bit for bit."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import tweakseq_b200 as t
from tweakseq_b200 import capi, synth
from tweakseq_b200.fasta import read_distmat, write_fasta
from oracle import pyoracle as o

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
AA = "ARNDCQEGHILKMFPSTWYVBZX"
NT = os.cpu_count() or 1


def gpu_run(seqs, **kw):
    with t.Context(**kw) as ctx:
        ctx.set_sequences(seqs)
        ctx.run()
        d = None if kw.get("flags", 0) & t.FLAG_NO_DISTANCES else ctx.distances()
        return ctx.scores(), d, ctx.self_scores(), ctx.stats()


def oracle_run(seqs, alphabet=0, go=None, ge=1):
    go = (11 if alphabet == 0 else 10) if go is None else go
    enc = [o.encode(s, alphabet) for s in seqs]
    mat = o.matrix(alphabet)
    s, cells = o.all_pairs(enc, mat, go, ge, nthreads=NT)
    selfs = np.array([o.self_score(e, mat) for e in enc], dtype=np.int32)
    return s, o.distances(s, selfs), selfs, cells


def assert_same(seqs, alphabet=0, go=None, ge=1, **kw):
    s, d, selfs, st = gpu_run(seqs, alphabet=alphabet, gap_open=-1 if go is None else go, gap_extend=ge, **kw)
    rs, rd, rselfs, cells = oracle_run(seqs, alphabet, go, ge)
    assert (s == rs).all(), np.nonzero(s != rs)[0][:10]
    assert (selfs == rselfs).all()
    assert d.tobytes() == rd.tobytes()
    return st, cells


def ragged(rng, n, lo, hi, letters=AA):
    return ["".join(rng.choice(list(letters), int(l))) for l in rng.integers(lo, hi, n)]


def test_device_is_b200_class():
    assert t.load_library().tsq_device_count() >= 1


def test_golden_allpairs():
    g = json.load(open(os.path.join(HERE, "golden", "allpairs_small.json")))
    s, d, selfs, _ = gpu_run(g["seqs"], gap_open=g["go"], gap_extend=g["ge"])
    assert s.tolist() == g["scores"]
    assert selfs.tolist() == g["self"]
    assert [float(x).hex() for x in d] == g["distances_hex"]


def test_golden_pairs_each_as_a_two_sequence_job():
    pairs = json.load(open(os.path.join(HERE, "golden", "pairs.json")))
    for p in pairs:
        s, _, _, _ = gpu_run([p["a"], p["b"]], alphabet=p["alphabet"], gap_open=p["go"], gap_extend=p["ge"])
        assert int(s[0]) == p["score"], p


@pytest.mark.parametrize("go,ge", [(None, 1), (5, 2), (0, 0), (0, 3), (30, 0), (1, 7)])
def test_ragged_protein_all_gap_models(go, ge):
    rng = np.random.default_rng(31)
    seqs = ragged(rng, 77, 0, 140)
    seqs[4] = ""; seqs[40] = ""; seqs[9] = "acdefg-hik.lmn pq"; seqs[11] = "W"; seqs[12] = "*1?JOU"
    assert_same(seqs, 0, go, ge)


@pytest.mark.parametrize("n", [1, 2, 3, 31, 32, 33, 64, 65, 129])
def test_group_and_chunk_edges(n):
    rng = np.random.default_rng(n)
    seqs = ragged(rng, n, 1, 90)
    if n == 1:
        s, d, selfs, _ = gpu_run(seqs)
        assert len(s) == 0 and len(d) == 0 and len(selfs) == 1
    else:
        assert_same(seqs)


def test_zero_sequences_and_all_empty():
    s, d, selfs, _ = gpu_run([])
    assert len(s) == 0 and len(selfs) == 0
    s, d, _, _ = gpu_run(["", "--", " "])
    assert s.tolist() == [0, 0, 0] and d.tolist() == [1.0, 1.0, 1.0]


@pytest.mark.parametrize("lens", [(49, 50, 51), (59, 60, 61, 120, 121), (1, 299, 300, 301), (31, 32, 33, 64)])
def test_strip_width_boundaries(lens):
    rng = np.random.default_rng(sum(lens))
    seqs = []
    for l in lens:
        seqs += ragged(rng, 9, l, l + 1, AA[:20])
    assert_same(seqs)


def test_config1_full():
    _, seqs = synth.config(1)
    st, cells = assert_same(seqs)
    assert st["cells"] == cells == synth.total_cells(seqs)
    fam = synth.protein(100, (200, 400, 300, 30), 1, family=True)
    assert_same(fam)


def test_config2_full_1000x300():
    _, seqs = synth.config(2)
    st, cells = assert_same(seqs)
    assert st["n_pairs"] == 499500 and st["cells"] == cells == 44955000000


def test_longer_proteins_and_mixed_lengths():
    rng = np.random.default_rng(77)
    seqs = ragged(rng, 40, 600, 1500, AA[:20]) + ragged(rng, 30, 1, 50, AA[:20])
    assert_same(seqs)


def test_nucleotide_short_reads():
    rng = np.random.default_rng(78)
    seqs = ragged(rng, 50, 0, 400, "ACGTN") + ["acgu-nn", "RYKM"]
    assert_same(seqs, alphabet=1)
    assert_same(seqs, alphabet=1, go=4, ge=4)


def test_custom_symmetric_matrix():
    rng = np.random.default_rng(79)
    m = rng.integers(-7, 9, (23, 23))
    m = np.triu(m) + np.triu(m, 1).T
    seqs = ragged(rng, 45, 1, 120)
    with t.Context(matrix=m.astype(np.int8), gap_open=6, gap_extend=2) as ctx:
        ctx.set_sequences(seqs)
        ctx.run()
        s = ctx.scores()
    enc = [o.encode(x) for x in seqs]
    ref, _ = o.all_pairs(enc, m.astype(np.int8), 6, 2, nthreads=NT)
    assert (s == ref).all()


def test_input_order_does_not_change_scores():
    rng = np.random.default_rng(80)
    seqs = ragged(rng, 60, 5, 200)
    s1, d1, _, _ = gpu_run(seqs)
    perm = rng.permutation(len(seqs))
    s2, d2, _, _ = gpu_run([seqs[k] for k in perm])
    n = len(seqs)
    for a in range(n):
        for b in range(a + 1, n):
            i, j = sorted((int(perm[a]), int(perm[b])))
            assert s2[t.pair_index(a, b, n)] == s1[t.pair_index(i, j, n)]
            assert d2[t.pair_index(a, b, n)] == d1[t.pair_index(i, j, n)]


def test_properties_identical_and_gapped_sequences():
    rng = np.random.default_rng(81)
    base = ragged(rng, 20, 50, 300, AA[:20])
    gapped = [s[:10] + "-" * 3 + s[10:] + ".." for s in base]
    s, d, selfs, _ = gpu_run(base + base + gapped)
    n = 60
    for k in range(20):
        assert s[t.pair_index(k, k + 20, n)] == selfs[k]          # S(x, x) = sum of diagonal
        assert d[t.pair_index(k, k + 20, n)] == 0.0
        assert s[t.pair_index(k, k + 40, n)] == selfs[k]          # gaps are stripped before DP
    s2, _, _, _ = gpu_run(base + base + gapped)                    # idempotent
    assert (s == s2).all()


def test_staged_api_and_repeat_compute_is_stable():
    _, seqs = synth.config(2, 0.2)
    with t.Context() as ctx:
        ctx.set_sequences(seqs)
        ctx.upload()
        ctx.compute(); ctx.download()
        a = ctx.scores()
        ctx.compute(); ctx.compute(); ctx.download()
        b = ctx.scores()
        st = ctx.stats()
    assert (a == b).all() and st["kernel_ms"] > 0 and st["launches"] >= 1


def test_partition_slabs_tile_the_matrix():
    rng = np.random.default_rng(82)
    seqs = ragged(rng, 150, 30, 260, AA[:20])
    full, _, _, _ = gpu_run(seqs, flags=t.FLAG_NO_DISTANCES)
    import torch
    world = 3
    bufs, ranges = [], []
    ctxs = []
    for r in range(world):
        ctx = t.Context(part_rank=r, part_world=world, flags=t.FLAG_NO_DISTANCES)
        ctx.set_sequences(seqs); ctx.upload(); ctx.compute(); ctx.synchronize()
        ranges.append(ctx.partition())
        arr, first = ctx.device_slab()                 # rank 0: the whole triangle; the others: their slab only
        assert first == (0 if r == 0 else ranges[r][0])
        bufs.append(torch.as_tensor(arr, device="cuda"))
        ctxs.append(ctx)
    assert ranges == capi.plan_partition([len(s) for s in seqs], world)
    assert ranges[0][0] == 0 and ranges[-1][1] == len(full)
    root = bufs[0]
    assert len(root) == len(full)
    for r in range(1, world):
        b, e = ranges[r]
        assert len(bufs[r]) == e - b
        root[b:e] = bufs[r]                            # what the NCCL gather does across ranks
    ctxs[0].finalize(); ctxs[0].download()
    assert (ctxs[0].scores() == full).all()
    for c in ctxs:
        c.close()


def test_run_fasta_writes_clustalo_distmat(tmp_path):
    rng = np.random.default_rng(83)
    seqs = ragged(rng, 12, 20, 90, AA[:20])
    labels = [f"seq{k}" for k in range(12)]
    fin, fout = str(tmp_path / "in.fa"), str(tmp_path / "out.mat")
    write_fasta(fin, labels, seqs, [f">{l} some description" for l in labels])
    log = []
    tool = t.B200Gotoh()
    tool.align = False                       # fout = the distance matrix (clustalo hand-off), not the alignment
    assert tool.run(fin, fout, log=log.append) == 0
    lab, rows = read_distmat(fout)
    assert lab == labels and len(rows) == 12 and any("GCUPS" in m for m in log)
    _, rd, _, _ = oracle_run(seqs)
    for i in range(12):
        assert rows[i][i] == 0.0
        for j in range(i + 1, 12):
            assert abs(rows[i][j] - rd[t.pair_index(i, j, 12)]) < 5e-7 and rows[j][i] == rows[i][j]
    assert tool.run(str(tmp_path / "missing.fa"), fout) == -7


def test_cancel_and_state_errors():
    with t.Context() as ctx:
        with pytest.raises(t.TsqError) as e:
            ctx.upload()
        assert e.value.status == -6
        ctx.set_sequences(["ACD", "ACE"])
        flag = C.c_int(1)
        with pytest.raises(t.TsqError) as e:
            ctx.run(cancel=flag)
        assert e.value.status == -5
        with pytest.raises(t.TsqError):
            ctx.scores()
        ctx.run()
        assert ctx.scores().tolist() == [o.score_str("ACD", "ACE")]


def test_backend_object_and_progress_callback():
    tool = t.B200Gotoh()
    seen = []
    _, seqs = synth.config(1)
    s, d = tool.distance_matrix(seqs[:30], progress=lambda f, m: seen.append(f))
    rs, rd, _, _ = oracle_run(seqs[:30])
    assert (s == rs).all() and d.tobytes() == rd.tobytes()
    assert seen[0] == 0.0 and seen[-1] == 1.0 and tool.last_stats["cells"] > 0
    cells = [[ord(c) for c in q] for q in seqs[:5]]
    cells[0][3] |= 0x80            # an excluded residue cell (Sequence.h:36)
    s2, _ = tool.distance_matrix_from_cells(cells)
    rs2, _, _, _ = oracle_run([seqs[0][:3] + seqs[0][4:]] + seqs[1:5])
    assert (s2 == rs2).all()


def test_dpx_probe_reports_the_alu_rate():
    with t.Context() as ctx:
        ops, mhz = ctx.measure_dpx_rate()
    assert 50 < ops < 80 and 1000 < mhz < 2200


@pytest.mark.parametrize("K", [30, 32, 36, 40, 44, 48, 50, 52, 56])
def test_every_strip_width_variant(K, monkeypatch):
    """Each instantiated kernel variant (strip width K), odd and even row counts, both gap-model
    specialisations (immediate -ge' for the defaults, runtime for the rest)."""
    monkeypatch.setenv("TSQ_FORCE_K", str(K))
    rng = np.random.default_rng(K)
    seqs = ragged(rng, 70, 1, 3 * K + 7, AA[:20]) + ragged(rng, 6, K, K + 1) + ragged(rng, 6, 2 * K - 1, 2 * K)
    st, _ = assert_same(seqs)
    assert st["strip_width"] == K
    assert_same(seqs, 0, 7, 3)


# ---- regime 2: the anti-diagonal wavefront kernels (packed wave16, and 32-bit wave32) ----------------
WAVE = [pytest.param(0, id="wave16"), pytest.param(4, id="wave32")]   # 4 = TSQ_FLAG_NO_WAVE16


@pytest.mark.parametrize("wf", WAVE)
def test_force_wavefront_protein_ragged_matches_oracle(wf):
    rng = np.random.default_rng(90)
    seqs = ragged(rng, 48, 0, 420) + ["W", "AC", "ACD"]
    st, cells = assert_same(seqs, flags=t.FLAG_FORCE_S32 | wf)
    assert st["cells_s32"] == cells and st["cells_s16"] == 0
    assert_same(seqs, 0, 3, 2, flags=t.FLAG_FORCE_S32 | wf)


@pytest.mark.parametrize("wf", WAVE)
def test_force_wavefront_nucleotide_ragged_matches_oracle(wf):
    rng = np.random.default_rng(91)
    seqs = ragged(rng, 40, 1, 2600, "ACGT") + ragged(rng, 6, 1020, 1030, "ACGTN") + ["A", "ACGTACGT"]
    assert_same(seqs, alphabet=1, flags=t.FLAG_FORCE_S32 | wf)
    assert_same(seqs[:20], alphabet=1, go=0, ge=2, flags=t.FLAG_FORCE_S32 | wf)


@pytest.mark.parametrize("kw,ctas", [(32, 2), (24, 3), (16, 3), (8, 2)])
def test_every_packed_wavefront_variant(kw, ctas, monkeypatch):
    """Each instantiated wave16 variant (columns per lane, CTAs per SM): sequences shorter than, equal to
    and several times one pass of 32*kw columns; default and runtime gap-model specialisations."""
    monkeypatch.setenv("TSQ_FORCE_KW16", f"{kw},{ctas}")
    rng = np.random.default_rng(kw)
    seqs = (ragged(rng, 14, 1, 3 * 32 * kw + 50, "ACGT") + ragged(rng, 3, 32 * kw - 1, 32 * kw + 2, "ACGTN") +
            ragged(rng, 2, 1023, 1026, "ACGT"))
    st, cells = assert_same(seqs, alphabet=1, flags=t.FLAG_FORCE_S32)
    assert st["cells_s32"] == cells
    assert_same(seqs[:12], alphabet=1, go=4, ge=2, flags=t.FLAG_FORCE_S32)


@pytest.mark.parametrize("kw,ctas", [(32, 2), (24, 3), (16, 4), (16, 3), (8, 4), (8, 2)])
def test_every_32_bit_wavefront_variant(kw, ctas, monkeypatch):
    monkeypatch.setenv("TSQ_FORCE_KW", f"{kw},{ctas}")
    rng = np.random.default_rng(100 + kw + ctas)
    seqs = ragged(rng, 12, 1, 3 * 32 * kw + 50, "ACGT") + ragged(rng, 3, 32 * kw - 1, 32 * kw + 2, "ACGTN")
    st, cells = assert_same(seqs, alphabet=1, flags=t.FLAG_FORCE_S32 | t.FLAG_NO_WAVE16)
    assert st["cells_s32"] == cells


@pytest.mark.parametrize("wf", WAVE)
def test_long_nucleotide_sequences_use_both_regimes(wf):
    rng = np.random.default_rng(92)
    seqs = (ragged(rng, 24, 10, 400, "ACGT") + ragged(rng, 5, 7300, 9500, "ACGT") + ragged(rng, 2, 12000, 12600, "ACGT")
            + [""])
    fam = synth.nucleotide(3, 8000, 9000, 4, family=True)
    st, cells = assert_same(seqs + fam, alphabet=1, flags=wf)
    assert st["cells_s16"] > 0 and st["cells_s32"] > 0 and st["cells_s16"] + st["cells_s32"] == cells


@pytest.mark.parametrize("wf", WAVE)
def test_long_protein_beyond_the_16_bit_range(wf):
    rng = np.random.default_rng(93)
    seqs = ragged(rng, 10, 50, 300, AA[:20]) + ragged(rng, 2, 4700, 5200, AA[:20])
    st, _ = assert_same(seqs, flags=wf)
    assert st["cells_s32"] > 0


def test_wave16_extreme_scores_walk_the_base():
    """Identical / poly-W long sequences drive H far from 0 (up to +11 per cell) and unrelated ones far
    below: the moving base of the packed wavefront kernel has to follow both."""
    rng = np.random.default_rng(96)
    w = "W" * 6000
    c = "C" * 5500
    r1 = "".join(rng.choice(list(AA[:20]), 5200))
    seqs = [w, w[:5900], c, r1, r1[:2000] + r1[2100:], "ACDEFGHIKL" * 480, "A" * 30]
    assert_same(seqs)
    g1 = "".join(rng.choice(list("ACGT"), 16000))
    nts = ["A" * 15000, "A" * 14000 + "C" * 900, g1, g1[:9000] + g1[9500:], "ACGT" * 2500, "T" * 8000, "ACGTN"]
    assert_same(nts, alphabet=1)


def test_large_gap_penalties_fall_back_to_the_32_bit_wavefront():
    rng = np.random.default_rng(97)
    seqs = ragged(rng, 12, 100, 900, "ACGT")
    assert_same(seqs, alphabet=1, go=120, ge=9, flags=t.FLAG_FORCE_S32)   # window too wide for wave16


def test_partition_slabs_with_wavefront_rows():
    rng = np.random.default_rng(94)
    seqs = ragged(rng, 30, 20, 200, "ACGT") + ragged(rng, 6, 7300, 8000, "ACGT")
    full, _, _, _ = gpu_run(seqs, alphabet=1, flags=t.FLAG_NO_DISTANCES)
    import torch
    world, ctxs, bufs, ranges = 2, [], [], []
    for r in range(world):
        ctx = t.Context(alphabet=1, part_rank=r, part_world=world, flags=t.FLAG_NO_DISTANCES)
        ctx.set_sequences(seqs); ctx.upload(); ctx.compute(); ctx.synchronize()
        ranges.append(ctx.partition()); bufs.append(torch.as_tensor(ctx.device_slab()[0], device="cuda")); ctxs.append(ctx)
    assert ranges == capi.plan_partition([len(s) for s in seqs], world, alphabet=1)
    b, e = ranges[1]
    bufs[0][b:e] = bufs[1]                              # rank 1 holds its slab only
    ctxs[0].finalize(); ctxs[0].download()
    assert (ctxs[0].scores() == full).all()
    rs, _, _, _ = oracle_run(seqs, 1)
    assert (full == rs).all()
    for c in ctxs:
        c.close()


def test_flat_input_form_equals_pointer_form():
    rng = np.random.default_rng(95)
    seqs = ragged(rng, 40, 0, 150) + ["ac-d e", ""]
    a, _, _, _ = gpu_run(seqs)
    buf, offs = capi.flatten(seqs)
    with t.Context() as ctx:
        ctx.set_sequences_flat(buf, offs)
        ctx.run()
        b = ctx.scores()
    assert (a == b).all()


# ---- BASELINE.json's sizes: sampled oracle comparison + size-independent properties ------------------
def _sampled_check(seqs, alphabet, nsample, seed, go=None, ge=1, flags=0):
    n = len(seqs)
    with t.Context(alphabet=alphabet, flags=flags | t.FLAG_NO_DISTANCES) as ctx:
        buf, offs = capi.flatten(seqs)
        ctx.set_sequences_flat(buf, offs)
        ctx.run()
        s = ctx.scores()
        st = ctx.stats()
    assert st["cells"] == synth.total_cells(seqs)
    rng = np.random.default_rng(seed)
    pi = rng.integers(0, n - 1, nsample)
    pj = rng.integers(1, n, nsample)
    keep = pi < pj
    pi, pj = pi[keep], pj[keep]
    # plus the complete first and last rows of the triangle
    pi = np.concatenate([pi, np.zeros(n - 1, dtype=pi.dtype), np.arange(0, n - 1)])
    pj = np.concatenate([pj, np.arange(1, n), np.full(n - 1, n - 1)])
    enc = [o.encode(x, alphabet) for x in seqs]
    g = (11 if alphabet == 0 else 10) if go is None else go
    ref, _ = o.pair_list(enc, pi, pj, o.matrix(alphabet), g, ge, nthreads=NT)
    idx = pi.astype(np.int64) * n - pi.astype(np.int64) * (pi.astype(np.int64) + 1) // 2 + (pj.astype(np.int64) - pi - 1)
    assert (s[idx] == ref).all()
    return s, st


def test_config3_full_size_sampled_and_properties():
    """10,000 x 400 aa (49,995,000 pairs, 8.0e12 cells): >= 2e4 sampled pairs + first/last rows vs
    the oracle; planted duplicates must score their self score (S(x,x) = sum of the diagonal)."""
    _, seqs = synth.config(3)
    seqs[9999] = seqs[0]
    seqs[5000] = seqs[17]
    s, st = _sampled_check(seqs, 0, 30000, 3)
    n = len(seqs)
    mat = o.matrix(0)
    assert s[t.pair_index(0, 9999, n)] == o.self_score(o.encode(seqs[0]), mat)
    assert s[t.pair_index(17, 5000, n)] == o.self_score(o.encode(seqs[17]), mat)
    assert st["n_pairs"] == 49995000 and st["cells_s32"] == 0


def test_config5_scaled_twin_sampled():
    """configs[4] scaled to 20,000 x 150 aa (2.0e8 pairs, 4.5e12 cells), sampled."""
    _, seqs = synth.config(5, 0.2)
    _sampled_check(seqs, 0, 60000, 5)


def test_config4_scaled_twin_wavefront_sampled():
    """configs[3] scaled to 20 x 10-30 kb nucleotide genomes: the 32-bit wavefront kernel."""
    _, seqs = synth.config(4, 0.04)
    n = len(seqs)
    with t.Context(alphabet=1, flags=t.FLAG_NO_DISTANCES) as ctx:
        ctx.set_sequences(seqs)
        ctx.run()
        s, st = ctx.scores(), ctx.stats()
    assert st["cells_s32"] == synth.total_cells(seqs) and st["cells_s16"] == 0
    rng = np.random.default_rng(4)
    sel = rng.choice(n * (n - 1) // 2, 40, replace=False)
    iu, ju = np.triu_indices(n, 1)
    enc = [o.encode(x, 1) for x in seqs]
    ref, _ = o.pair_list(enc, iu[sel], ju[sel], o.matrix(1), 10, 1, nthreads=NT)
    assert (s[sel] == ref).all()
    fam = synth.nucleotide(6, 10000, 12000, 4, family=True)     # related genomes: long diagonal runs
    assert_same(fam, alphabet=1)
    assert_same(fam, alphabet=1, flags=t.FLAG_NO_WAVE16)


def test_cancel_while_the_kernels_run_stops_early():
    """'Stop' (SeqEditMainWin.cpp:803-812): the cancel flag is polled by the host while the stream runs
    and by every warp at each task fetch, so a long job drains within one task instead of finishing."""
    import threading, time
    _, seqs = synth.config(3, 0.8)            # 8,000 x 400 aa: ~0.6 s of kernel time
    flag = C.c_int(0)
    out = {}
    with t.Context(flags=t.FLAG_NO_DISTANCES) as ctx:
        ctx.set_sequences(seqs)
        ctx.upload()
        ctx.compute(); ctx.synchronize()
        full_ms = ctx.stats()["kernel_ms"]

        def work():
            t0 = time.perf_counter()
            try:
                ctx.run(cancel=flag)
                out["status"] = 0
            except t.TsqError as e:
                out["status"] = e.status
            out["s"] = time.perf_counter() - t0

        th = threading.Thread(target=work)
        th.start()
        time.sleep(0.15)
        flag.value = 1
        th.join(timeout=60)
        assert out["status"] == -5
        assert out["s"] < 0.15 + 0.6 * full_ms / 1e3, (out, full_ms)
        with pytest.raises(t.TsqError):
            ctx.scores()
        flag.value = 0
        ctx.run(cancel=flag)                      # the context is reusable after a cancel
        assert len(ctx.scores()) == len(seqs) * (len(seqs) - 1) // 2


def test_randomised_gap_models_all_three_kernels_agree():
    """Random gap penalties and ragged lengths: the packed inter-task kernel, the packed wavefront
    kernel and the 32-bit wavefront kernel must all reproduce the oracle."""
    rng = np.random.default_rng(2026)
    for trial in range(14):
        alphabet = int(rng.integers(0, 2))
        letters = "ACGTN" if alphabet else AA
        go, ge = int(rng.integers(0, 21)), int(rng.integers(0, 6))
        n = int(rng.integers(3, 40))
        hi = 1500 if alphabet else 600
        seqs = ragged(rng, n, 0, hi, letters)
        rs, _, _, _ = oracle_run(seqs, alphabet, go, ge)
        for flags in (0, t.FLAG_FORCE_S32, t.FLAG_FORCE_S32 | t.FLAG_NO_WAVE16):
            s, _, _, _ = gpu_run(seqs, alphabet=alphabet, gap_open=go, gap_extend=ge, flags=flags | t.FLAG_NO_DISTANCES)
            assert (s == rs).all(), (trial, alphabet, go, ge, flags, np.nonzero(s != rs)[0][:5])


@pytest.mark.parametrize("flags,alphabet", [(8, 0), (1, 1), (8 | 1, 0), (4, 1)])
def test_host_planner_matches_the_context_in_every_mode(flags, alphabet):
    """tsq_plan_partition (host only) and tsq_upload must cut the rows identically whatever kernels the
    flags select (identity keys, forced wavefront, no packed wavefront)."""
    rng = np.random.default_rng(flags * 10 + alphabet)
    letters = "ACGT" if alphabet else AA[:20]
    seqs = ragged(rng, 90, 0, 300, letters) + ragged(rng, 3, 7400 if alphabet else 4600, 7600 if alphabet else 4700, letters)
    planned = capi.plan_partition([len(s) for s in seqs], 3, alphabet=alphabet, flags=flags)
    got = []
    for r in range(3):
        with t.Context(alphabet=alphabet, part_rank=r, part_world=3, flags=flags | t.FLAG_NO_DISTANCES) as ctx:
            ctx.set_sequences(seqs)
            ctx.upload()
            got.append(ctx.partition())
    assert got == planned


def test_backend_object_identity_tree_and_consensus(tmp_path):
    tool = t.B200Gotoh()
    _, seqs = synth.config(1, 0.25)
    tool.identity = True
    s, d = tool.distance_matrix(seqs)
    enc = [o.encode(x) for x in seqs]
    rs, rk, rd = o.all_pairs_id(enc, o.matrix(0), 11, 1)
    assert (s == rs).all() and d.tobytes() == rd.tobytes()
    l, r, h = tool.guide_tree(seqs, labels=[f"s{k}" for k in range(len(seqs))], newick_path=str(tmp_path / "g.dnd"))
    ol, orr, oh = o.upgma(rd, len(seqs))
    assert (l == ol).all() and (r == orr).all() and h.tobytes() == oh.tobytes()
    assert open(tmp_path / "g.dnd").read().strip().endswith(";")
    rows = [x[:200].ljust(200, "-") for x in seqs]
    assert tool.consensus(rows) == o.consensus(rows)


# ---- published optima (tests/published_vectors.py): the anchors outside this repository, through the library ----------
def test_published_alignment_scores_on_the_gpu():
    """Durbin et al. 1998 fig. 2.5 (BLOSUM50, linear gap 8: score 1) and the Needleman-Wunsch worked example (+1/-1/-1:
    score 0), as custom matrices through the C ABI; both argument orders, and in the same job as unrelated filler so that
    the pairs sit in full 32-subject chunks of the packed kernel."""
    from published_vectors import VECTORS
    rng = np.random.default_rng(77)
    for name, alphabet, a, b, mat, go, ge, published in VECTORS:
        letters = ("AEGHPW" if mat is not None else "ARNDCQEGHILKMFPSTWYV") if alphabet == 0 else "ACGT"
        filler = ["".join(rng.choice(list(letters), int(l))) for l in rng.integers(1, 40, 70)]
        seqs = [a, b] + filler + [b, a]
        with t.Context(alphabet=alphabet, matrix=mat, gap_open=go, gap_extend=ge, flags=t.FLAG_NO_DISTANCES) as ctx:
            ctx.set_sequences(seqs)
            ctx.run()
            s = ctx.scores()
        n = len(seqs)
        assert s[t.pair_index(0, 1, n)] == published, name
        assert s[t.pair_index(n - 2, n - 1, n)] == published, name
        assert s[t.pair_index(0, n - 2, n)] == published and s[t.pair_index(1, n - 1, n)] == published, name   # (a, b) again, across the job
        enc = [o.encode(x, alphabet) for x in seqs]
        ref, _ = o.all_pairs(enc, o.matrix(alphabet) if mat is None else mat, go, ge, nthreads=4)
        assert (s == ref).all()
