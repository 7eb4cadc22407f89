"""Several devices behind ONE context (tsq_params.n_devices) and the sharded
result path of partitioned jobs (tsq_results_sharded / tsq_set_result_buffers), through the C ABI,
against the CPU oracle.

Every test runs twice: on real devices when the box has at least two (gpurun --gpus 2), and with
TSQ_MULTI_SAME_DEVICE=1 -- all children of the context on the one device, the same host logic
(per-device planning, slab finalize, peer-store finalize, own-stream downloads, collected
distances) -- so that a one-GPU box still covers it."""
import ctypes as C
import os

import numpy as np
import pytest

import tweakseq_b200 as t
from tweakseq_b200 import capi, synth
from tweakseq_b200.fasta import write_fasta
from oracle import pyoracle as o

pytestmark = pytest.mark.gpu
AA = "ARNDCQEGHILKMFPSTWYVBZX"
NT = os.cpu_count() or 1


def ragged(rng, n, lo, hi, letters=AA):
    return ["".join(rng.choice(list(letters), int(l))) for l in rng.integers(lo, hi, n)]


def oracle_run(seqs, alphabet=0, go=None, ge=1):
    go = (11 if alphabet == 0 else 10) if go is None else go
    enc = [o.encode(s, alphabet) for s in seqs]
    mat = o.matrix(alphabet)
    s, cells = o.all_pairs(enc, mat, go, ge, nthreads=NT)
    selfs = np.array([o.self_score(e, mat) for e in enc], dtype=np.int32)
    return s, o.distances(s, selfs), selfs, cells


@pytest.fixture(params=["devices", "same-device"])
def ndev(request, monkeypatch):
    """Number of devices a multi-device context of this flavour can take (>= 2)."""
    have = t.load_library().tsq_device_count()
    if request.param == "devices":
        if have < 2:
            pytest.skip("one device here: the real multi-device flavour runs under gpurun --gpus 2")
        return min(have, 4)
    monkeypatch.setenv("TSQ_MULTI_SAME_DEVICE", "1")
    return 3


def multi_run(seqs, n_devices, **kw):
    with t.Context(n_devices=n_devices, **kw) as ctx:
        ctx.set_sequences(seqs)
        ctx.run()
        d = None if kw.get("flags", 0) & t.FLAG_NO_DISTANCES else ctx.distances()
        return ctx.scores(), d, ctx.stats(), [ctx.device_stats(k) for k in range(n_devices)], ctx.results_sharded()


def test_fixed_length_results_stay_sharded_and_match_the_oracle(ndev):
    _, seqs = synth.config(2, 0.4)            # 400 x 300 aa
    s, d, st, per, sharded = multi_run(seqs, ndev)
    rs, rd, _, cells = oracle_run(seqs)
    assert sharded
    assert (s == rs).all() and d.tobytes() == rd.tobytes()
    assert st["cells"] == cells == sum(p["cells"] for p in per) and st["n_pairs"] == len(rs)
    assert all(p["cells"] > 0 and p["kernel_ms"] > 0 for p in per)
    assert max(p["cells"] for p in per) / (cells / ndev) < 1.05      # the planner balances DP cells
    assert st["kernel_ms"] == max(p["kernel_ms"] for p in per)


def test_ragged_input_is_unsorted_into_the_first_device_by_peer_stores(ndev):
    rng = np.random.default_rng(41)
    seqs = ragged(rng, 190, 0, 380) + ["", "W", "acd-ef"]
    s, d, st, per, sharded = multi_run(seqs, ndev)
    rs, rd, _, cells = oracle_run(seqs)
    assert not sharded
    assert (s == rs).all(), np.nonzero(s != rs)[0][:10]
    assert d.tobytes() == rd.tobytes()
    assert st["cells"] == cells
    s2, d2, _, _, _ = multi_run(seqs, ndev, gap_open=4, gap_extend=3)
    rs2, rd2, _, _ = oracle_run(seqs, 0, 4, 3)
    assert (s2 == rs2).all() and d2.tobytes() == rd2.tobytes()


def test_mixed_regimes_nucleotide(ndev):
    rng = np.random.default_rng(42)
    seqs = ragged(rng, 40, 10, 900, "ACGT") + synth.nucleotide(5, 7600, 9200, 3) + ["ACGTN"]
    s, d, st, per, _ = multi_run(seqs, ndev, alphabet=1)
    rs, rd, _, cells = oracle_run(seqs, 1)
    assert (s == rs).all() and d.tobytes() == rd.tobytes()
    assert st["cells_s16"] > 0 and st["cells_s32"] > 0 and st["cells"] == cells


def test_fixed_length_wavefront_and_scores_only(ndev):
    seqs = synth.nucleotide(14, 8000, 8000, 6)
    s, d, st, _, sharded = multi_run(seqs, ndev, alphabet=1, flags=t.FLAG_NO_DISTANCES)
    rs, _, _, _ = oracle_run(seqs, 1)
    assert sharded and d is None and (s == rs).all() and st["cells_s32"] > 0


def test_identity_mode_on_several_devices(ndev):
    rng = np.random.default_rng(43)
    seqs = ragged(rng, 60, 20, 200, AA[:20])
    with t.Context(n_devices=ndev, flags=t.FLAG_IDENTITY) as ctx:
        ctx.set_sequences(seqs)
        ctx.run()
        s, nid, d = ctx.scores(), ctx.identities(), ctx.distances()
    with t.Context(flags=t.FLAG_IDENTITY) as ctx:
        ctx.set_sequences(seqs)
        ctx.run()
        s1, nid1, d1 = ctx.scores(), ctx.identities(), ctx.distances()
    assert (s == s1).all() and (nid == nid1).all() and d.tobytes() == d1.tobytes()
    rs, _, _, _ = oracle_run(seqs)
    assert (s == rs).all()


@pytest.mark.parametrize("fixed", [True, False])
def test_tree_and_alignment_on_the_first_device_equal_the_single_device_ones(ndev, fixed):
    rng = np.random.default_rng(44)
    seqs = synth.protein(48, 120, 8, family=True) if fixed else ragged(rng, 40, 30, 160, AA[:20])
    with t.Context(n_devices=ndev) as ctx:
        ctx.set_sequences(seqs)
        ctx.run()
        tree = ctx.guide_tree()
        rows, order = ctx.msa()
        a = ctx.align_pair(3, 17)
    with t.Context() as ctx:
        ctx.set_sequences(seqs)
        ctx.run()
        tree1 = ctx.guide_tree()
        rows1, order1 = ctx.msa()
        a1 = ctx.align_pair(3, 17)
    assert all((x == y).all() for x, y in zip(tree, tree1))
    assert rows == rows1 and order == order1 and a == a1


def test_run_fasta_with_n_devices(ndev, tmp_path):
    rng = np.random.default_rng(45)
    seqs = ragged(rng, 30, 40, 150, AA[:20])
    labels = [f"s{k}" for k in range(len(seqs))]
    fin = str(tmp_path / "in.fa")
    write_fasta(fin, labels, seqs, [f">{l}" for l in labels])
    out1, outn = str(tmp_path / "one.fa"), str(tmp_path / "many.fa")
    log = []
    assert capi.run_fasta(fin, out1, flags=t.FLAG_MSA_OUT) == 0
    assert capi.run_fasta(fin, outn, log=log.append, flags=t.FLAG_MSA_OUT, n_devices=ndev) == 0
    assert open(out1).read() == open(outn).read()
    assert any(f"on {ndev} devices" in m for m in log), log
    assert not os.path.exists(outn + ".dnd")          # the tree file is opt-in next to an alignment
    assert capi.run_fasta(fin, outn, flags=t.FLAG_MSA_OUT | capi.FLAG_KEEP_TREE, n_devices=ndev) == 0
    assert os.path.exists(outn + ".dnd")


def test_caller_owned_result_buffers_and_repeat_runs(ndev):
    _, seqs = synth.config(2, 0.3)
    n = len(seqs)
    scores = np.full(n * (n - 1) // 2, -7, dtype=np.int32)
    dist = np.full(n * (n - 1) // 2, -7.0, dtype=np.float64)
    rs, rd, _, _ = oracle_run(seqs)
    with t.Context(n_devices=ndev) as ctx:
        ctx.set_result_buffers(scores, dist)
        ctx.set_sequences(seqs)
        for _ in range(2):
            scores[:] = -7
            ctx.run()
            assert (scores == rs).all() and dist.tobytes() == rd.tobytes()
            assert ctx.scores(copy=False).ctypes.data == scores.ctypes.data
        ctx.set_result_buffers(None)
        scores[:] = -7
        ctx.run()
        assert (ctx.scores() == rs).all() and (scores == -7).all()


def test_cancel_and_errors_on_a_multi_device_context(ndev):
    with t.Context(n_devices=ndev) as ctx:
        ctx.set_sequences(["ACD", "ACE", "WWW"])
        with pytest.raises(t.TsqError) as e:
            ctx.run(cancel=C.c_int(1))
        assert e.value.status == -5
        with pytest.raises(t.TsqError):
            ctx.scores()
        with pytest.raises(t.TsqError):
            ctx.set_stream(1234)
        ctx.run()
        assert ctx.scores().tolist() == [o.score_str("ACD", "ACE"), o.score_str("ACD", "WWW"), o.score_str("ACE", "WWW")]
    have = t.load_library().tsq_device_count()
    if "TSQ_MULTI_SAME_DEVICE" not in os.environ:
        with pytest.raises(t.TsqError) as e:
            t.Context(n_devices=have + 1)
        assert e.value.status == -2
        with t.Context(n_devices=-1) as ctx:          # every usable device of the box
            ctx.set_sequences(synth.protein(64, 100, 3))
            ctx.run()
            assert len(ctx.scores()) == 64 * 63 // 2
    with pytest.raises(t.TsqError):
        t.Context(n_devices=2, part_world=2)


# ---- partitioned ranks (the torchrun flavour), emulated on one device ----------------------------------------
@pytest.mark.parametrize("world", [2, 3])
def test_ranks_of_a_fixed_length_job_download_their_own_slabs_into_one_result(world):
    """What tweakseq_b200/distributed.py does across processes: every rank finalizes its slab and copies it
    into ONE host result (here a numpy array shared by the contexts; across processes a /dev/shm segment)."""
    _, seqs = synth.config(2, 0.35)
    n = len(seqs)
    scores = np.full(n * (n - 1) // 2, -7, dtype=np.int32)
    dist = np.full(n * (n - 1) // 2, -7.0, dtype=np.float64)
    ctxs = []
    for r in range(world):
        ctx = t.Context(part_rank=r, part_world=world)
        ctx.set_result_buffers(scores, dist)
        ctx.set_sequences(seqs)
        ctx.upload()
        assert ctx.results_sharded()
        arr, first = ctx.device_slab()
        b, e = ctx.partition()
        assert first == b and arr.__cuda_array_interface__["shape"][0] == e - b     # a rank holds its slab only
        ctx.compute()
        ctx.download()
        ctxs.append(ctx)
    rs, rd, _, _ = oracle_run(seqs)
    assert (scores == rs).all() and dist.tobytes() == rd.tobytes()
    with pytest.raises(t.TsqError):
        ctxs[1].guide_tree()              # a rank holds a slab of the matrix only
    for c in ctxs:
        c.close()


def test_a_lone_rank_without_result_buffers_keeps_its_slab_only():
    _, seqs = synth.config(2, 0.2)
    rs, _, _, _ = oracle_run(seqs)
    with t.Context(part_rank=1, part_world=2, flags=t.FLAG_NO_DISTANCES) as ctx:
        ctx.set_sequences(seqs)
        ctx.run()
        b, e = ctx.partition()
        s = ctx.scores()
    assert len(s) == e - b and (s == rs[b:e]).all()


def test_ragged_ranks_still_gather_into_rank_0():
    rng = np.random.default_rng(46)
    seqs = ragged(rng, 90, 10, 200, AA[:20])
    with t.Context(part_rank=0, part_world=2) as ctx:
        ctx.set_sequences(seqs)
        ctx.upload()
        assert not ctx.results_sharded()
        arr, first = ctx.device_slab()
        assert first == 0 and arr.__cuda_array_interface__["shape"][0] == len(seqs) * (len(seqs) - 1) // 2
    with t.Context(part_rank=1, part_world=2) as ctx:
        ctx.set_sequences(seqs)
        ctx.upload()
        ctx.compute()
        with pytest.raises(t.TsqError) as e:
            ctx.finalize()
        assert e.value.status == -6


# ---- streamed results: row ranges leave for the host behind the launch that computed them --------------------------
@pytest.mark.parametrize("chunks", ["1", "3", "8", "16"])
@pytest.mark.parametrize("dist", [True, False])
@pytest.mark.parametrize("by", ["counters", "launches"])
def test_streamed_results_equal_the_plain_ones(chunks, dist, by, monkeypatch):
    """tsq_stream_results, both flavours.  "counters" (the default): ONE launch; the kernel writes the distances itself
    and ticks a counter per row range at the end of every task, the copy stream waits for a range's count
    (cuStreamWaitValue32) and copies it out while the launch runs on.  "launches" (TSQ_STREAM_LAUNCHES, the fallback
    without stream memory operations): the tasks in `chunks` launches over consecutive row ranges, each followed on the
    side stream by finalize + copy-out of those rows.  TSQ_STREAM_CHUNKS forces the split on a small job."""
    monkeypatch.setenv("TSQ_STREAM_CHUNKS", chunks)
    if by == "launches":
        monkeypatch.setenv("TSQ_STREAM_LAUNCHES", "1")
    _, seqs = synth.config(2, 0.45)                     # 450 x 300 aa; also an odd row count below
    seqs = seqs[:449]
    rs, rd, _, _ = oracle_run(seqs)
    flags = 0 if dist else t.FLAG_NO_DISTANCES
    with t.Context(flags=flags) as ctx:
        ctx.set_sequences(seqs)
        ctx.stream_results(True)
        ctx.upload()
        for _ in range(2):                              # staged, repeated: the side stream is reused
            ctx.compute()
            ctx.download()
            assert (ctx.scores() == rs).all()
            if dist:
                assert ctx.distances().tobytes() == rd.tobytes()
        st = ctx.stats()
        assert st["d2h_bytes"] == len(rs) * (12 if dist else 4)
        assert st["launches"] == 1 if by == "counters" else st["launches"] >= min(int(chunks), 8)
        if dist:
            tree = ctx.guide_tree()                     # the main stream sees the side stream's distances
            assert len(tree[0]) == len(seqs) - 1
        ctx.stream_results(False)
        ctx.compute(); ctx.download()
        assert (ctx.scores() == rs).all()


def test_streaming_is_skipped_where_the_results_need_an_unsort(monkeypatch):
    monkeypatch.setenv("TSQ_STREAM_CHUNKS", "4")
    rng = np.random.default_rng(47)
    seqs = ragged(rng, 120, 0, 200, AA[:20])            # ragged: the un-sort scatters rows, nothing can leave early
    rs, rd, _, _ = oracle_run(seqs)
    with t.Context() as ctx:
        ctx.set_sequences(seqs)
        ctx.stream_results(True)
        ctx.run()
        assert (ctx.scores() == rs).all() and ctx.distances().tobytes() == rd.tobytes()


def test_streamed_results_on_several_devices_into_caller_buffers(ndev, monkeypatch):
    monkeypatch.setenv("TSQ_STREAM_CHUNKS", "3")
    _, seqs = synth.config(2, 0.5)
    n = len(seqs)
    scores = np.full(n * (n - 1) // 2, -7, dtype=np.int32)
    dist = np.full(n * (n - 1) // 2, -7.0, dtype=np.float64)
    rs, rd, _, _ = oracle_run(seqs)
    with t.Context(n_devices=ndev) as ctx:
        ctx.set_result_buffers(scores, dist)
        ctx.set_sequences(seqs)
        ctx.run()                                       # tsq_run streams by itself
        assert (scores == rs).all() and dist.tobytes() == rd.tobytes()
        tree = ctx.guide_tree()
    with t.Context() as ctx:
        ctx.set_sequences(seqs)
        ctx.run()
        tree1 = ctx.guide_tree()
    assert all((x == y).all() for x, y in zip(tree, tree1))
