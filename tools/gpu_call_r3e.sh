#!/bin/bash
# Round 2 (1 GPU): the whole GPU test tier and the default bench line after the alignment / reader work
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -q --timeout 600 ; echo "exit $?" ) > gpurun_out/r3e_tests.log 2>&1
tail -n 4 gpurun_out/r3e_tests.log
( timeout 900 python bench.py ; echo "exit $?" ) > gpurun_out/r3e_bench_1gpu.log 2>&1
grep '^{"metric"' gpurun_out/r3e_bench_1gpu.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value', round(d['value']), 'e2e', round(d['e2e']['value']), d['e2e']['stages_ms_rank0'], 'c3', round(d['c3']['value']), round(d['c3']['e2e']['value']))
for k,v in d['e2e_plugin'].items():
    print(k, v if isinstance(v,str) else (round(v['wall_ms'],2), v['stages_ms']))"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
