"""Short driver for ncu captures: a few tsq_compute passes of one workload."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tweakseq_b200 as t
from tweakseq_b200 import synth
which = sys.argv[1] if len(sys.argv) > 1 else "c2"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
alpha, seqs = {"c2": lambda: synth.config(2), "c3s": lambda: synth.config(3, 0.3), "c1": lambda: synth.config(1),
               "c4s": lambda: synth.config(4, 0.06), "c4m": lambda: synth.config(4, 0.24), "c4": lambda: synth.config(4), "c3": lambda: synth.config(3), "c5s": lambda: synth.config(5, 0.2), "c5t": lambda: synth.config(5, 0.1)}[which]()
flags = int(os.environ.get("TSQ_FLAGS", "0"))
with t.Context(alphabet=alpha, flags=flags | t.FLAG_NO_DISTANCES) as ctx:
    ctx.set_sequences(seqs); ctx.upload()
    for _ in range(reps):
        ctx.compute(); ctx.synchronize()
        st = ctx.stats()
        print(which, "kernel_ms", round(st["kernel_ms"], 4), "GCUPS", round(st["gcups_kernel"], 1), "K", st["strip_width"], flush=True)
