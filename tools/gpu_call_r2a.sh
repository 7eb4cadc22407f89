#!/bin/bash
# Round 2, first call (2 GPUs): the whole GPU test tier incl. the multi-device and range-edge tests, the torchrun
# sharded check, bench at N=1, N=2 (torchrun) and N=2 in one process, the pipe probe.
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2a_gpus.txt 2>&1
( timeout 900 python -m pytest tests -m gpu -q --timeout 400 ; echo "exit $?" ) > gpurun_out/r2a_tests.log 2>&1
tail -n 25 gpurun_out/r2a_tests.log
( timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_sharded.py ; echo "exit $?" ) > gpurun_out/r2a_check_sharded_2gpu.log 2>&1
grep -E "sharded ok|Error|error|exit" gpurun_out/r2a_check_sharded_2gpu.log | tail -n 12
( timeout 400 python bench.py ; echo "exit $?" ) > gpurun_out/r2a_bench_1gpu.log 2>&1
tail -n 2 gpurun_out/r2a_bench_1gpu.log | cut -c1-3000
( timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 ; echo "exit $?" ) > gpurun_out/r2a_bench_2gpu.log 2>&1
grep '^{"metric"' gpurun_out/r2a_bench_2gpu.log | cut -c1-3000; tail -n 1 gpurun_out/r2a_bench_2gpu.log
( timeout 400 python bench.py --gpus 2 --inprocess --steps 20 --warmup 5 ; echo "exit $?" ) > gpurun_out/r2a_bench_2gpu_inprocess.log 2>&1
grep '^{"metric"' gpurun_out/r2a_bench_2gpu_inprocess.log | cut -c1-3000; tail -n 1 gpurun_out/r2a_bench_2gpu_inprocess.log
( timeout 60 tools/pipe_probe2 ; echo "exit $?" ) > gpurun_out/r2a_pipe_probe2.jsonl 2>&1
grep "1024\|512" gpurun_out/r2a_pipe_probe2.jsonl | grep -E "V2|V5|V6|V7"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_smoke.log 2>&1; tail -n 1 gpurun_out/r2a_smoke.log
