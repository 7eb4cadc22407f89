#!/bin/bash
# Round 2 (1 GPU): results streamed out of ONE running launch (per-range counters + cuStreamWaitValue32),
# vector encoder, upload without its host wait.  A/B against the launch-per-range flavour inside one call.
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q --timeout 400 -x ; echo "exit $?" ) > gpurun_out/r2n_tests.log 2>&1
tail -n 5 gpurun_out/r2n_tests.log
show() { grep '^{"metric"' "$1" | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$2', 'value', round(d['value']), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],3), d['e2e']['stages_ms_rank0'], 'parity', d['parity']['mismatches'], '| c3', round(d['c3']['value']), round(d['c3']['e2e']['value']), d['c3']['e2e']['stages_ms_rank0'], d['c3']['parity']['mismatches'])"; }
for rep in 1 2; do
  ( timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-plugin ) > gpurun_out/r2n_bench_counters_$rep.log 2>&1
  show gpurun_out/r2n_bench_counters_$rep.log counters
  ( TSQ_STREAM_LAUNCHES=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-plugin ) > gpurun_out/r2n_bench_launches_$rep.log 2>&1
  show gpurun_out/r2n_bench_launches_$rep.log launches
done
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
