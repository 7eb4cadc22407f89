#!/bin/bash
# Round 2 (1 GPU): evidence pass after the streaming and wave16 step work: full bench line (all legs), the CPU arm,
# configs[3] in full with parity, ncu launch list of the bench command, ncu --set full of wave16 at full occupancy.
mkdir -p gpurun_out
( timeout 900 python bench.py ; echo "exit $?" ) > gpurun_out/r2s_bench_1gpu.log 2>&1
tail -n 2 gpurun_out/r2s_bench_1gpu.log | cut -c1-1500
( timeout 600 python bench.py --impl reference ; echo "exit $?" ) > gpurun_out/r2s_bench_reference.log 2>&1
tail -n 2 gpurun_out/r2s_bench_reference.log | cut -c1-600
( timeout 900 python bench.py --workload c4 --no-cpu ; echo "exit $?" ) > gpurun_out/r2s_bench_c4.log 2>&1
tail -n 2 gpurun_out/r2s_bench_c4.log | cut -c1-1500
( timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2s_launches_bench_c2.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-c3 --no-plugin ; echo "exit $?" ) > gpurun_out/r2s_launches.log 2>&1
tail -n 1 gpurun_out/r2s_launches.log
( timeout 500 ncu --set full --clock-control none --import-source on -k regex:wave16 -c 1 -f -o gpurun_out/r2s_wave16_c4m python tools/prof_run.py c4m 1 ; echo "exit $?" ) > gpurun_out/r2s_ncu_w16.log 2>&1
tail -n 2 gpurun_out/r2s_ncu_w16.log
