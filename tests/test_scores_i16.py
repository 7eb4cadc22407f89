"""TSQ_FLAG_SCORES_I16: the scores as int16 (10 GB instead of 20 GB at configs[4]), on every route
a result takes to the host -- whole-triangle download, a device's own slab, rows streamed out of the running launch
(per-range counters) and out of one launch per range, several devices behind one context -- always equal to the
oracle's scores, with the distances untouched; refused where a score could leave int16."""
import numpy as np
import pytest

import tweakseq_b200 as t
from tweakseq_b200 import synth
from oracle import pyoracle as o

pytestmark = pytest.mark.gpu
NT = 8
AA = "ARNDCQEGHILKMFPSTWYV"


def oracle_run(seqs, alphabet=0):
    enc = [o.encode(s, alphabet) for s in seqs]
    mat = o.matrix(alphabet)
    s, _ = o.all_pairs(enc, mat, 11 if alphabet == 0 else 10, 1, nthreads=NT)
    selfs = np.array([o.self_score(e, mat) for e in enc], dtype=np.int32)
    return s, o.distances(s, selfs)


def ragged(rng, n, lo, hi):
    return ["".join(rng.choice(list(AA), int(l))) for l in rng.integers(lo, hi, n)]


def check(ctx, rs, rd, dist=True):
    s16 = ctx.scores16()
    assert s16.dtype == np.int16 and (s16.astype(np.int32) == rs).all()
    if dist:
        assert ctx.distances().tobytes() == rd.tobytes()
    with pytest.raises(t.TsqError) as e:
        ctx.scores()
    assert e.value.status == -6
    assert ctx.stats()["d2h_bytes"] == len(rs) * (10 if dist else 2)


@pytest.mark.parametrize("dist", [True, False])
def test_ragged_input_whole_triangle(dist):
    rng = np.random.default_rng(811)
    seqs = ragged(rng, 150, 0, 260)
    rs, rd = oracle_run(seqs)
    with t.Context(flags=t.FLAG_SCORES_I16 | (0 if dist else t.FLAG_NO_DISTANCES)) as ctx:
        ctx.set_sequences(seqs)
        ctx.run()
        check(ctx, rs, rd, dist)
        if dist:
            assert len(ctx.guide_tree()[0]) == len(seqs) - 1      # what follows the matrix is unaffected


@pytest.mark.parametrize("route", ["run", "launches", "staged"])
def test_fixed_length_input_streamed_and_staged(route, monkeypatch):
    """run: rows leave the ONE running launch (the kernel writes the int16 copy itself); launches: one launch per row
    range, narrowed behind each; staged: no streaming, narrowed before the download."""
    _, seqs = synth.config(2, 0.4)
    rs, rd = oracle_run(seqs)
    monkeypatch.setenv("TSQ_STREAM_CHUNKS", "5")
    if route == "launches":
        monkeypatch.setenv("TSQ_STREAM_LAUNCHES", "1")
    with t.Context(flags=t.FLAG_SCORES_I16) as ctx:
        ctx.set_sequences(seqs)
        for _ in range(2):
            if route == "staged":
                ctx.upload(); ctx.compute(); ctx.download()
            else:
                ctx.run()
            check(ctx, rs, rd)


@pytest.mark.parametrize("kind", ["fixed", "ragged"])
def test_several_devices_behind_one_context(kind, monkeypatch):
    monkeypatch.setenv("TSQ_MULTI_SAME_DEVICE", "1")       # children on one device: the host logic is the same
    rng = np.random.default_rng(812)
    seqs = synth.config(2, 0.3)[1] if kind == "fixed" else ragged(rng, 140, 1, 220)
    rs, rd = oracle_run(seqs)
    with t.Context(n_devices=3, flags=t.FLAG_SCORES_I16) as ctx:
        ctx.set_sequences(seqs)
        ctx.run()
        s16 = ctx.scores16()
        assert (s16.astype(np.int32) == rs).all() and ctx.distances().tobytes() == rd.tobytes()


def test_nucleotides_and_long_sequences_inside_the_range():
    seqs = synth.nucleotide(5, 5200, 6400, 5)              # 6 400 x 5 = 32 000 < 32 767: the longest admitted under +5/-4
    rs, rd = oracle_run(seqs, alphabet=1)
    with t.Context(alphabet=1, flags=t.FLAG_SCORES_I16) as ctx:
        ctx.set_sequences(seqs)
        ctx.run()
        check(ctx, rs, rd)


def test_refused_where_a_score_could_leave_int16():
    with t.Context(flags=t.FLAG_SCORES_I16) as ctx:
        ctx.set_sequences(["W" * 3000, "W" * 3000, "A" * 10])      # 11 x 3 000 = 33 000
        with pytest.raises(t.TsqError) as e:
            ctx.run()
        assert e.value.status == -9 and "int16" in str(e.value)
        ctx.set_sequences(["W" * 2900, "W" * 2900, "A" * 10])      # 31 900 (+ the bound's margin) still fits
        ctx.run()
        assert int(ctx.scores16()[0]) == 11 * 2900


def test_flag_combinations_that_are_refused():
    with pytest.raises(t.TsqError):
        t.Context(flags=t.FLAG_SCORES_I16 | t.FLAG_IDENTITY)
    with t.Context(flags=t.FLAG_SCORES_I16) as ctx:
        buf = np.zeros(3, np.int32)
        with pytest.raises(t.TsqError):
            ctx.set_result_buffers(buf, None)
    with t.Context() as ctx:
        ctx.set_sequences(["ACD", "ACE"]); ctx.run()
        with pytest.raises(t.TsqError):
            ctx.scores16()
