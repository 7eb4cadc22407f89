"""Where the end-to-end milliseconds go (host buffers -> results on the host), C2."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tweakseq_b200 as t
from tweakseq_b200 import synth
_, seqs = synth.config(2)
ctx = t.Context()
for it in range(6):
    t0 = time.perf_counter(); ctx.set_sequences(seqs)
    t1 = time.perf_counter(); ctx.upload()
    t2 = time.perf_counter(); ctx.compute(); ctx.synchronize()
    t3 = time.perf_counter(); ctx.finalize(); ctx.synchronize()
    t4 = time.perf_counter(); ctx.download()
    t5 = time.perf_counter(); s = ctx.scores(); d = ctx.distances()
    t6 = time.perf_counter()
    st = ctx.stats()
    print(f"set_sequences {1e3*(t1-t0):.3f}  upload {1e3*(t2-t1):.3f} (lib {st['upload_ms']:.3f})  compute {1e3*(t3-t2):.3f} (kernel {st['kernel_ms']:.3f})  "
          f"finalize {1e3*(t4-t3):.3f}  download {1e3*(t5-t4):.3f} (lib {st['download_ms']:.3f})  numpy-copy {1e3*(t6-t5):.3f}  total {1e3*(t5-t0):.3f} ms")
