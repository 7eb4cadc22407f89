"""Small jobs through every kernel, for compute-sanitizer (memcheck / racecheck / initcheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import tweakseq_b200 as t
from tweakseq_b200 import synth
from oracle import pyoracle as o

def job(seqs, alphabet=0, flags=0, tree=False):
    with t.Context(alphabet=alphabet, flags=flags) as ctx:
        ctx.set_sequences(seqs); ctx.run()
        s = ctx.scores()
        if tree: ctx.guide_tree()
    enc = [o.encode(x, alphabet) for x in seqs]
    ref, _ = o.all_pairs(enc, o.matrix(alphabet), 11 if alphabet == 0 else 10, 1, nthreads=4)
    assert (s == ref).all()

rng = np.random.default_rng(0)
prot = ["".join(rng.choice(list("ARNDCQEGHILKMFPSTWYV"), int(l))) for l in rng.integers(0, 130, 70)]
job(prot, tree=True)                                            # gotoh16 + finalize + subject_db + upgma
nt = ["".join(rng.choice(list("ACGT"), int(l))) for l in rng.integers(1, 1300, 12)]
job(nt, alphabet=1, flags=t.FLAG_FORCE_S32)                      # wave16 (TMA ring)
job(nt[:8], alphabet=1, flags=t.FLAG_FORCE_S32 | t.FLAG_NO_WAVE16)  # wave32
with t.Context(alphabet=1) as ctx:                               # traceback: single-CTA and 8-CTA cluster variants
    ctx.set_sequences([nt[0][:300], nt[1][:200], "ACGT" * 600, "ACGA" * 560])
    ctx.upload()
    enc = [o.encode(x, 1) for x in (nt[0][:300], nt[1][:200], "ACGT" * 600, "ACGA" * 560)]
    for i, j in ((0, 1), (2, 3)):
        assert ctx.align_pair(i, j) == o.traceback(enc[i], enc[j], o.matrix(1), 10, 1, 1)
# round 2: several devices behind one context (children on one device here), ragged (peer-store finalize) and
# fixed-length (slab finalize, streamed out in three launches), caller-owned result buffers
os.environ["TSQ_MULTI_SAME_DEVICE"] = "1"
os.environ["TSQ_STREAM_CHUNKS"] = "3"
for seqs in (prot, synth.protein(96, 60, 3)):
    n = len(seqs)
    bs, bd = np.zeros(n * (n - 1) // 2, np.int32), np.zeros(n * (n - 1) // 2, np.float64)
    with t.Context(n_devices=3) as ctx:
        ctx.set_result_buffers(bs, bd)
        ctx.set_sequences(seqs); ctx.run(); ctx.guide_tree()
    enc = [o.encode(x) for x in seqs]
    ref, _ = o.all_pairs(enc, o.matrix(0), 11, 1, nthreads=4)
    assert (bs == ref).all()
fam = ["ACDEFGHIKLMNPQRSTVWY" * 3, "ACDEFGHIKLMNPQRSTVWY" * 3, "ACDEFGHIKLMNPQRSTVWA" * 3, "ACDEFGHIKLMNPQRSTVWY" * 2 + "ACDEFGHIKL"]
with t.Context(flags=t.FLAG_IDENTITY | t.FLAG_KIMURA) as ctx:      # Kimura branch of finalize
    ctx.set_sequences(fam); ctx.run()
    assert (ctx.distances() >= 0).all()
print("sanitize_run ok")
