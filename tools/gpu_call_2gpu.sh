#!/bin/bash
mkdir -p gpurun_out
( timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_sharded.py ; echo "exit $?" ) > gpurun_out/check_sharded_2gpu.log 2>&1
tail -n 4 gpurun_out/check_sharded_2gpu.log
( timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 ; echo "exit $?" ) > gpurun_out/bench_2gpu.log 2>&1
grep '^{"metric"' gpurun_out/bench_2gpu.log | cut -c1-400; tail -n 1 gpurun_out/bench_2gpu.log
