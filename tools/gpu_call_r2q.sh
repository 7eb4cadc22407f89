#!/bin/bash
# Round 2 (1 GPU): wave16 per-step overhead (ring test 1 in 32, running ring offset / step position, 48-byte boundary
# records on one running pointer, PRMT+IMAD profile addresses).  Parity first, then A/B on configs[3] in full against
# the library built before the change (tools/ab/libtsqb200_base.so), alternating, inside this one call.
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_range_edges.py -m gpu -q --timeout 600 -x ; echo "exit $?" ) > gpurun_out/r2q_tests.log 2>&1
tail -n 4 gpurun_out/r2q_tests.log
cp tweakseq_b200/libtsqb200.so /tmp/new.so
for rep in 1 2; do
  cp tools/ab/libtsqb200_base.so tweakseq_b200/libtsqb200.so
  echo "base:" >> gpurun_out/r2q_ab.log; timeout 300 python tools/prof_run.py c4 2 >> gpurun_out/r2q_ab.log 2>&1
  cp /tmp/new.so tweakseq_b200/libtsqb200.so
  echo "new:" >> gpurun_out/r2q_ab.log; timeout 300 python tools/prof_run.py c4 2 >> gpurun_out/r2q_ab.log 2>&1
done
cat gpurun_out/r2q_ab.log
