#!/bin/bash
# Round 2, sixth call (1 GPU): whole GPU tier after streaming / Kimura / MSA tables / wave16 24-16 fallback, bench N=1.
mkdir -p gpurun_out
df -h /dev/shm /tmp | tail -n 2; free -g | head -2; nproc
( timeout 900 python -m pytest tests -m gpu -q --timeout 400 ; echo "exit $?" ) > gpurun_out/r2f_tests.log 2>&1
tail -n 12 gpurun_out/r2f_tests.log
( timeout 400 python bench.py ; echo "exit $?" ) > gpurun_out/r2f_bench_1gpu.log 2>&1
grep '^{"metric"' gpurun_out/r2f_bench_1gpu.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('c2', round(d['value']), round(d['e2e']['value']), d['e2e']['ms_per_step'], 'c3', round(d['c3']['value']), round(d['c3']['e2e']['value']), d['parity']['mismatches'], d['c3']['parity']['mismatches'])
print(json.dumps(d['e2e_plugin'])[:1600])
print(json.dumps(d['cpu_baseline'])[:900])"
tail -n 1 gpurun_out/r2f_bench_1gpu.log
( timeout 300 python bench.py --impl reference --steps 3 --warmup 1 ; echo "exit $?" ) > gpurun_out/r2f_bench_ref.log 2>&1
tail -n 2 gpurun_out/r2f_bench_ref.log | cut -c1-700
