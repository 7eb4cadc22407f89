"""Randomised soak of the CUDA path against the CPU oracle: random alphabets, matrices, gap models,
length mixes (empties, singletons, strip/pass boundaries, long tails), kernel-selection flags,
identity mode (with and without the Kimura correction), partitions, several devices behind one context (real ones, or
all children on one device), streamed results in 1-8 launches, caller-owned result buffers, guide tree, progressive
alignment and traceback.  Runs until --seconds elapse; prints one line
per trial class and a final summary; exits non-zero at the first mismatch (with the seed)."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import tweakseq_b200 as t
from oracle import pyoracle as o

AA = "ARNDCQEGHILKMFPSTWYVBZX"
NT = os.cpu_count() or 1


def make_seqs(rng, alphabet):
    letters = "ACGTN" if alphabet else AA
    style = rng.integers(0, 6)
    n = int(rng.integers(2, 90))
    if style == 0:      # short ragged, many empties
        lens = rng.integers(0, 40, n)
    elif style == 1:    # around strip widths
        base = int(rng.choice([30, 32, 36, 40, 44, 48, 50, 52, 56]))
        lens = base * rng.integers(1, 5, n) + rng.integers(-1, 2, n)
    elif style == 2:    # typical protein / read lengths
        lens = rng.integers(50, 700 if not alphabet else 2500, n)
    elif style == 3:    # a few long ones among short ones (both regimes in one job)
        n = int(rng.integers(4, 30))
        lens = np.concatenate([rng.integers(1, 300, n), rng.integers(3000, 9000, int(rng.integers(1, 4)))])
    elif style == 4:    # fixed length (sort is the identity)
        lens = np.full(n, int(rng.integers(1, 400)))
    else:               # family: mutated copies of one root (non-trivial distances, many ties)
        L = int(rng.integers(20, 500))
        root = rng.choice(list(letters[:20] if not alphabet else "ACGT"), L)
        out = []
        for _ in range(n):
            s = root.copy()
            mut = rng.random(L) < rng.uniform(0.0, 0.5)
            s[mut] = rng.choice(list(letters), int(mut.sum()))
            keep = rng.random(L) > rng.uniform(0.0, 0.15)
            out.append("".join(s[keep]))
        return out
    lens = np.maximum(lens, 0)
    return ["".join(rng.choice(list(letters), int(l))) for l in lens]


def trial(seed, counts):
    rng = np.random.default_rng(seed)
    alphabet = int(rng.integers(0, 2))
    nsym = 5 if alphabet else 23
    seqs = make_seqs(rng, alphabet)
    go, ge = (None, 1) if rng.random() < 0.3 else (int(rng.integers(0, 40)), int(rng.integers(0, 8)))
    gov = (10 if alphabet else 11) if go is None else go
    matrix = None
    if rng.random() < 0.25:
        m = rng.integers(-9, 12, (nsym, nsym))
        matrix = (np.triu(m) + np.triu(m, 1).T).astype(np.int8)
    mat = o.matrix(alphabet) if matrix is None else matrix
    flags = int(rng.choice([0, 0, 0, t.FLAG_FORCE_S32, t.FLAG_FORCE_S32 | t.FLAG_NO_WAVE16, t.FLAG_NO_WAVE16]))
    identity = rng.random() < 0.25
    kimura = identity and rng.random() < 0.4
    if identity:
        flags |= t.FLAG_IDENTITY
    if kimura:
        flags |= t.FLAG_KIMURA
    # several devices behind one context, streamed results, caller-owned result buffers
    ndev = int(rng.choice([1, 1, 2, 3]))
    if ndev > 1 and t.load_library().tsq_device_count() < ndev:
        os.environ["TSQ_MULTI_SAME_DEVICE"] = "1"
    chunks = rng.choice(["", "1", "3", "8"])
    if chunks:
        os.environ["TSQ_STREAM_CHUNKS"] = str(chunks)
    else:
        os.environ.pop("TSQ_STREAM_CHUNKS", None)
    own_buffers = rng.random() < 0.3
    enc = [o.encode(s, alphabet) for s in seqs]
    n = len(seqs)
    tag = (f"seed={seed} alphabet={alphabet} n={n} go={gov} ge={ge} flags={flags} custom={matrix is not None} maxlen={max(map(len, seqs))} "
           f"ndev={ndev} chunks={chunks!r} own_buffers={own_buffers}")
    try:
        with t.Context(alphabet=alphabet, gap_open=-1 if go is None else go, gap_extend=ge, matrix=matrix, flags=flags, n_devices=ndev) as ctx:
            npairs = n * (n - 1) // 2
            if own_buffers and npairs:
                bs, bd = np.full(npairs, -7, dtype=np.int32), np.full(npairs, -7.0, dtype=np.float64)
                ctx.set_result_buffers(bs, bd)
            ctx.set_sequences(seqs)
            ctx.run()
            s, d, st = ctx.scores(), ctx.distances(), ctx.stats()
            if own_buffers and npairs:
                assert (bs == s).all() and bd.tobytes() == d.tobytes(), "caller buffers " + tag
            if ndev > 1:
                counts["multi_device"] += 1
            if identity:
                rs, rk, rd = o.all_pairs_id(enc, mat, gov, ge)
                assert (ctx.identities() == rk).all(), "identities " + tag
                if kimura:      # reached only if every pair is below D = 0.75 (else TsqError -9 above)
                    rd = np.array([o.kimura(int(k), min(len(enc[i]), len(enc[j])))[0]
                                   for k, (i, j) in zip(rk, ((i, j) for i in range(n) for j in range(i + 1, n)))], dtype=np.float64)
                    counts["kimura"] += 1
            else:
                rs, _ = o.all_pairs(enc, mat, gov, ge, nthreads=NT)
                selfs = np.array([o.self_score(e, mat) for e in enc], dtype=np.int32)
                rd = o.distances(rs, selfs)
            assert (s == rs).all(), "scores " + tag + f" first={np.nonzero(s != rs)[0][:5]}"
            assert d.tobytes() == rd.tobytes(), "distances " + tag
            counts["pairs"] += len(rs)
            import torch
            ref_sorted = None if identity or n < 2 or ndev > 1 else torch.as_tensor(ctx.device_scores(), device="cuda").cpu().numpy().copy()
            counts["cells"] += st["cells"]
            if n >= 2 and rng.random() < 0.5:
                left, right, height = ctx.guide_tree()
                rl, rr, rh = o.upgma(rd, n)
                assert (left == rl).all() and (right == rr).all() and height.tobytes() == rh.tobytes(), "tree " + tag
                counts["trees"] += 1
                if n <= 48 and max(map(len, seqs)) <= 700:       # progressive alignment along that tree
                    rows, order = ctx.msa()
                    want, _ = o.msa(enc, mat, gov, ge, left, right, alphabet)
                    assert rows == want, "msa " + tag
                    assert sorted(order) == list(range(n)), "msa order " + tag
                    counts["alignments"] += 1
            for _ in range(3):
                i, j = int(rng.integers(0, n)), int(rng.integers(0, n))
                if len(enc[i]) * len(enc[j]) > 4_000_000:
                    continue
                got = ctx.align_pair(i, j)
                assert got == o.traceback(enc[i], enc[j], mat, gov, ge, alphabet), f"traceback ({i},{j}) " + tag
                counts["tracebacks"] += 1
        if n >= 4 and ref_sorted is not None and rng.random() < 0.3:      # slabs of a 3-way partition tile the sorted triangle
            world = 3
            got = np.zeros(n * (n - 1) // 2, dtype=np.int64)
            cover = np.zeros(n * (n - 1) // 2, dtype=np.int32)
            for r in range(world):
                with t.Context(alphabet=alphabet, gap_open=-1 if go is None else go, gap_extend=ge, matrix=matrix,
                               flags=(flags & ~t.FLAG_IDENTITY) | t.FLAG_NO_DISTANCES, part_rank=r, part_world=world) as ctx:
                    ctx.set_sequences(seqs)
                    ctx.upload(); ctx.compute(); ctx.synchronize()
                    b, e = ctx.partition()
                    arr, first = ctx.device_slab()          # rank 0 may hold the whole triangle, the others their slab only
                    slab = torch.as_tensor(arr, device="cuda")[b - first:e - first].cpu().numpy() if e > b else np.zeros(0, np.int32)
                    got[b:e] = slab
                    cover[b:e] += 1
            assert (cover == 1).all(), "partition cover " + tag
            assert (got == ref_sorted).all(), "partition values " + tag
            counts["partitions"] += 1
    except t.TsqError as e:
        if e.status == -9:                       # documented range refusal (e.g. identity keys of huge penalties)
            counts["range_refusals"] += 1
            return
        raise
    counts["trials"] += 1


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=120)
    ap.add_argument("--seed", type=int, default=1)
    a = ap.parse_args()
    counts = dict(trials=0, pairs=0, cells=0, trees=0, alignments=0, tracebacks=0, partitions=0, range_refusals=0, multi_device=0, kimura=0)
    t0 = time.time()
    seed = a.seed
    while time.time() - t0 < a.seconds:
        trial(seed, counts)
        seed += 1
    counts["seconds"] = round(time.time() - t0, 1)
    counts["first_seed"], counts["last_seed"] = a.seed, seed - 1
    print("soak ok", counts, flush=True)


if __name__ == "__main__":
    main()
