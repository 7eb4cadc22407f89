#!/bin/bash
# Round 2 (1 GPU): after moving the kernel-start event to the first launch -- GPU test tier, smoke, one bench line
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -q --timeout 600 ; echo "exit $?" ) > gpurun_out/r3n_tests.log 2>&1
tail -n 3 gpurun_out/r3n_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
( timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-plugin --no-c3 ) > gpurun_out/r3n_bench.log 2>&1
grep '^{"metric"' gpurun_out/r3n_bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value', round(d['value']), 'e2e', round(d['e2e']['value']), d['e2e']['stages_ms_rank0'], d['kernel_ms_per_rank'], d['roofline']['frac_mix'])"
