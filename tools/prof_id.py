"""Timing of identity mode (32-bit inter-task kernel) on configs[1]."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tweakseq_b200 as t
from tweakseq_b200 import synth
_, seqs = synth.config(2)
for flags, name in ((t.FLAG_IDENTITY, "identity (gotoh32)"), (0, "plain (gotoh16)")):
    with t.Context(flags=flags) as ctx:
        ctx.set_sequences(seqs); ctx.upload()
        for _ in range(3):
            ctx.compute(); ctx.synchronize()
        st = ctx.stats()
        print(name, "kernel_ms", round(st["kernel_ms"], 3), "GCUPS", round(st["gcups_kernel"], 1), "K", st["strip_width"], flush=True)
