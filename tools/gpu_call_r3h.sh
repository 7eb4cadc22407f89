#!/bin/bash
# Round 2 (1 GPU): progressive alignment -- parity, then base (before the tiled sweep) against the current build, one call.
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_msa.py tests/test_zz_aligner_cli.py -m gpu -q --timeout 600 -x ; echo "exit $?" ) > gpurun_out/r3h_tests.log 2>&1
tail -n 3 gpurun_out/r3h_tests.log
cp tweakseq_b200/libtsqb200.so /tmp/new.so
rm -f gpurun_out/r3h_msa_ab.log
for rep in 1 2; do
  cp tools/ab/libtsqb200_base.so tweakseq_b200/libtsqb200.so
  echo "base (diagonal sweep):" >> gpurun_out/r3h_msa_ab.log; timeout 600 python tools/prof_msa3.py >> gpurun_out/r3h_msa_ab.log 2>&1
  cp /tmp/new.so tweakseq_b200/libtsqb200.so
  echo "new:" >> gpurun_out/r3h_msa_ab.log; timeout 600 python tools/prof_msa3.py >> gpurun_out/r3h_msa_ab.log 2>&1
done
cat gpurun_out/r3h_msa_ab.log
