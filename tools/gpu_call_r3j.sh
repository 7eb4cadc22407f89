#!/bin/bash
# Round 2 (2 GPUs): end-of-round validation -- the whole GPU test tier on real devices, sharded check under torchrun,
# the bench at N = 1 and N = 2 (range boundaries by task count with a shrinking tail).
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -q --timeout 600 ; echo "exit $?" ) > gpurun_out/r3j_tests.log 2>&1
tail -n 3 gpurun_out/r3j_tests.log
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_sharded.py ; echo "exit $?" ) > gpurun_out/r3j_check_sharded_2gpu.log 2>&1
grep -E "sharded ok|Error|exit" gpurun_out/r3j_check_sharded_2gpu.log | cut -c1-100 | tail -n 6
show() { grep '^{"metric"' "$1" | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('N', d['n_gpus'], 'value', round(d['value']), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],3), d['e2e']['stages_ms_rank0'], 'parity', d['parity']['mismatches'], '| c3', round(d['c3']['value']), round(d['c3']['e2e']['value']), d['c3']['parity']['mismatches'])"; }
( timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-plugin ) > gpurun_out/r3j_bench_1gpu.log 2>&1; show gpurun_out/r3j_bench_1gpu.log
( timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 ) > gpurun_out/r3j_bench_2gpu.log 2>&1; show gpurun_out/r3j_bench_2gpu.log
( timeout 400 python bench.py --gpus 2 --inprocess --steps 20 --warmup 5 ) > gpurun_out/r3j_bench_2gpu_inprocess.log 2>&1; show gpurun_out/r3j_bench_2gpu_inprocess.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
