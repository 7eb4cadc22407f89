#!/bin/bash
# Round 2 (1 GPU): the default bench line and the CPU arm with the round's final build
mkdir -p gpurun_out
( timeout 600 python bench.py ; echo "exit $?" ) > gpurun_out/r3p_bench_1gpu.log 2>&1
grep '^{"metric"' gpurun_out/r3p_bench_1gpu.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value', round(d['value']), 'e2e', round(d['e2e']['value']), d['e2e']['stages_ms_rank0'], 'frac', round(d['roofline']['frac'],3), round(d['roofline']['frac_dpx_issue'],3), round(d['roofline']['frac_mix'],3), 'traffic', d['roofline']['traffic'], 'cpu', round(d['cpu_baseline']['value'],1), 'c3', round(d['c3']['value']), round(d['c3']['e2e']['value']), 'plugin c2', round(d['e2e_plugin']['c2']['wall_ms'],1))"
tail -n 1 gpurun_out/r3p_bench_1gpu.log
