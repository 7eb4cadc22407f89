#!/bin/bash
# Round 2, eleventh call (2 GPUs): per-device host threads (KidPool) -- multi-device tests on real devices, short soak,
# in-process bench at N=2.
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_multi_device.py tests/test_gpu_parity.py -m gpu -q --timeout 300 ; echo "exit $?" ) > gpurun_out/r2k_tests.log 2>&1
tail -n 4 gpurun_out/r2k_tests.log
( timeout 200 python tools/soak.py --seconds 60 --seed 30000 ; echo "exit $?" ) > gpurun_out/r2k_soak.log 2>&1
tail -n 2 gpurun_out/r2k_soak.log | cut -c1-400
( timeout 400 python bench.py --gpus 2 --inprocess --steps 20 --warmup 5 ; echo "exit $?" ) > gpurun_out/r2k_bench_2gpu_inprocess.log 2>&1
grep '^{"metric"' gpurun_out/r2k_bench_2gpu_inprocess.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('N', d['n_gpus'], 'value', round(d['value']), 'e2e', round(d['e2e']['value']), d['e2e']['ms_per_step'], d['e2e']['stages_ms_rank0'], 'parity', d['parity']['mismatches'], '| c3', round(d['c3']['value']), round(d['c3']['e2e']['value']), d['c3']['parity']['mismatches'])"
tail -n 1 gpurun_out/r2k_bench_2gpu_inprocess.log
