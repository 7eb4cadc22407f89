#!/bin/bash
# Round 2 (1 GPU): wave16 with the result extraction moved out of both step bodies -- parity, then A/B against the
# committed build (tools/ab/libtsqb200_cb.so) on configs[3] in full, alternating, one call.
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_range_edges.py -m gpu -q --timeout 600 -x ; echo "exit $?" ) > gpurun_out/r3i_tests.log 2>&1
tail -n 3 gpurun_out/r3i_tests.log
cp tweakseq_b200/libtsqb200.so /tmp/main.so
rm -f gpurun_out/r3i_ab.log
for rep in 1 2; do
  for v in cb main; do
    if [ $v = main ]; then cp /tmp/main.so tweakseq_b200/libtsqb200.so; else cp tools/ab/libtsqb200_$v.so tweakseq_b200/libtsqb200.so; fi
    echo "$v:" >> gpurun_out/r3i_ab.log; timeout 300 python tools/prof_run.py c4 1 >> gpurun_out/r3i_ab.log 2>&1
  done
done
cp /tmp/main.so tweakseq_b200/libtsqb200.so
cat gpurun_out/r3i_ab.log
