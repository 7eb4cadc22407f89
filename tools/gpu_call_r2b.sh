#!/bin/bash
# Round 2, second call (1 GPU): GPU test tier again (after the fixes + the LDS.128 wavefront profile), full configs[3]
# with its parity block, ncu launch list of the bench command and --set full captures of the two packed kernels.
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q --timeout 400 ; echo "exit $?" ) > gpurun_out/r2b_tests.log 2>&1
tail -n 8 gpurun_out/r2b_tests.log
( timeout 600 python bench.py --workload c4 --no-cpu ; echo "exit $?" ) > gpurun_out/r2b_bench_c4.log 2>&1
grep '^{"metric"' gpurun_out/r2b_bench_c4.log | cut -c1-1800; tail -n 1 gpurun_out/r2b_bench_c4.log
( timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b_launches_bench_c2.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-c3 --no-plugin ; echo "exit $?" ) > gpurun_out/r2b_launches.log 2>&1
tail -n 1 gpurun_out/r2b_launches.log
( timeout 300 ncu --set full --clock-control none --import-source on -k regex:gotoh16 -c 1 -f -o gpurun_out/r2b_gotoh16_c2 python tools/prof_run.py c2 2 ; echo "exit $?" ) > gpurun_out/r2b_ncu_g16.log 2>&1
tail -n 2 gpurun_out/r2b_ncu_g16.log
( timeout 400 ncu --set full --clock-control none --import-source on -k regex:wave16 -c 1 -f -o gpurun_out/r2b_wave16_c4s python tools/prof_run.py c4s 1 ; echo "exit $?" ) > gpurun_out/r2b_ncu_w16.log 2>&1
tail -n 2 gpurun_out/r2b_ncu_w16.log
ls -la gpurun_out/*.ncu-rep
