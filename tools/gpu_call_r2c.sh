#!/bin/bash
# Round 2, third call (1 GPU): four-row wavefront body -- parity tests, variant sweep on 120 genomes, full configs[3], ncu at full occupancy.
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q --timeout 400 -x ; echo "exit $?" ) > gpurun_out/r2c_tests.log 2>&1
tail -n 6 gpurun_out/r2c_tests.log
for v in "" "32,2" "16,3"; do
  ( TSQ_FORCE_KW16=$v timeout 120 python tools/prof_run.py c4m 3 ; echo "exit $?" ) > gpurun_out/r2c_w16_kw_${v/,/_}.log 2>&1
  echo "KW16=$v"; tail -n 2 gpurun_out/r2c_w16_kw_${v/,/_}.log
done
( timeout 600 python bench.py --workload c4 --no-cpu ; echo "exit $?" ) > gpurun_out/r2c_bench_c4.log 2>&1
grep '^{"metric"' gpurun_out/r2c_bench_c4.log | cut -c1-900; tail -n 1 gpurun_out/r2c_bench_c4.log
( timeout 500 ncu --set full --clock-control none --import-source on -k regex:wave16 -c 1 -f -o gpurun_out/r2c_wave16_c4m python tools/prof_run.py c4m 1 ; echo "exit $?" ) > gpurun_out/r2c_ncu_w16.log 2>&1
tail -n 2 gpurun_out/r2c_ncu_w16.log
