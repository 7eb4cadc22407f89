#!/bin/bash
# Round 2 (1 GPU): progressive alignment, A/B inside one call: the library as it was before the tiled sweep
# (tools/ab/libtsqb200_base.so: one cell per thread and diagonal, serial walk-back) against the current one, best of three.
mkdir -p gpurun_out
cp tweakseq_b200/libtsqb200.so /tmp/new.so
for rep in 1 2; do
  cp tools/ab/libtsqb200_base.so tweakseq_b200/libtsqb200.so
  echo "base (diagonal sweep):" >> gpurun_out/r3b_msa_ab.log; timeout 600 python tools/prof_msa3.py >> gpurun_out/r3b_msa_ab.log 2>&1
  cp /tmp/new.so tweakseq_b200/libtsqb200.so
  echo "new (tiled sweep, worker warps, warp walk-back):" >> gpurun_out/r3b_msa_ab.log; timeout 600 python tools/prof_msa3.py >> gpurun_out/r3b_msa_ab.log 2>&1
done
cat gpurun_out/r3b_msa_ab.log
