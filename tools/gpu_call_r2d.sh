#!/bin/bash
# Round 2, fourth call (1 GPU): adaptive wavefront variant (32 columns per lane), whole GPU tier, full configs[3].
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q --timeout 400 ; echo "exit $?" ) > gpurun_out/r2d_tests.log 2>&1
tail -n 8 gpurun_out/r2d_tests.log
( timeout 600 python bench.py --workload c4 --no-cpu ; echo "exit $?" ) > gpurun_out/r2d_bench_c4.log 2>&1
grep '^{"metric"' gpurun_out/r2d_bench_c4.log | cut -c1-900; tail -n 1 gpurun_out/r2d_bench_c4.log
( timeout 300 python bench.py ; echo "exit $?" ) > gpurun_out/r2d_bench_1gpu.log 2>&1
grep '^{"metric"' gpurun_out/r2d_bench_1gpu.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('c2', round(d['value']), round(d['e2e']['value']), 'c3', round(d['c3']['value']), round(d['c3']['e2e']['value']))
print(json.dumps(d['e2e_plugin'])[:1500])"
