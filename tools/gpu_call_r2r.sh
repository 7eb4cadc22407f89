#!/bin/bash
# Round 2 (1 GPU): wave16 step-overhead variants, parity of each, then A/B/C on configs[3] in full inside this call:
#   base = before the work, cb = one rare test per step + records, main = + lane 31 preloads the pass boundary (rotating shuffle)
mkdir -p gpurun_out
cp tweakseq_b200/libtsqb200.so /tmp/main.so
for v in main cb; do
  if [ $v = cb ]; then cp tools/ab/libtsqb200_cb.so tweakseq_b200/libtsqb200.so; else cp /tmp/main.so tweakseq_b200/libtsqb200.so; fi
  ( timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_range_edges.py -m gpu -q --timeout 600 -x ; echo "exit $?" ) > gpurun_out/r2r_tests_$v.log 2>&1
  echo $v; tail -n 3 gpurun_out/r2r_tests_$v.log
done
for rep in 1 2; do
  for v in base cb main; do
    if [ $v = main ]; then cp /tmp/main.so tweakseq_b200/libtsqb200.so; else cp tools/ab/libtsqb200_$v.so tweakseq_b200/libtsqb200.so; fi
    echo "$v:" >> gpurun_out/r2r_ab.log; timeout 300 python tools/prof_run.py c4 1 >> gpurun_out/r2r_ab.log 2>&1
  done
done
cp /tmp/main.so tweakseq_b200/libtsqb200.so
cat gpurun_out/r2r_ab.log
