#!/bin/bash
# Round 2 (2 GPUs): asynchronous bounce copies for the partial pages of caller-owned result buffers.
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q --timeout 400 ; echo "exit $?" ) > gpurun_out/r2m_tests.log 2>&1
tail -n 4 gpurun_out/r2m_tests.log
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_sharded.py ; echo "exit $?" ) > gpurun_out/r2m_check_sharded_2gpu.log 2>&1
grep -E "sharded ok|Error|exit" gpurun_out/r2m_check_sharded_2gpu.log | cut -c1-120 | tail -n 7
( timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 ; echo "exit $?" ) > gpurun_out/r2m_bench_2gpu.log 2>&1
grep '^{"metric"' gpurun_out/r2m_bench_2gpu.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('N', d['n_gpus'], 'value', round(d['value']), 'e2e', round(d['e2e']['value']), d['e2e']['ms_per_step'], d['e2e']['stages_ms_rank0'], 'parity', d['parity']['mismatches'], '| c3', round(d['c3']['value']), round(d['c3']['e2e']['value']), d['c3']['e2e']['stages_ms_rank0'], d['c3']['parity']['mismatches'])"
tail -n 1 gpurun_out/r2m_bench_2gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
