#!/bin/bash
# Round 2, seventh call (8 GPUs): multi-device tests on real devices, torchrun sharded check at N=8, the bench at N=8
# (weak configs[1] + strong configs[2]) under torchrun and in one process (tsq_params.n_devices), full configs[4].
mkdir -p gpurun_out
nvidia-smi -L | wc -l; df -h /dev/shm /tmp | tail -n 2; free -g | head -2 | tail -1
( timeout 600 python -m pytest tests/test_multi_device.py -m gpu -q --timeout 300 ; echo "exit $?" ) > gpurun_out/r2l_tests_multi.log 2>&1
tail -n 5 gpurun_out/r2l_tests_multi.log
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 tools/check_sharded.py ; echo "exit $?" ) > gpurun_out/r2l_check_sharded_8gpu.log 2>&1
grep -E "sharded ok|Error|exit" gpurun_out/r2l_check_sharded_8gpu.log | cut -c1-160 | tail -n 8
summ='
import json,sys
for l in sys.stdin:
    if not l.startswith("{\"metric\""): continue
    d=json.loads(l)
    print(d["config"]["launch"], "| N", d["n_gpus"], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"],3), round(d["e2e"]["ms_per_step"],3), "parity", d["parity"]["mismatches"], d["parity"]["pairs_checked"], "ranks", d["kernel_ms_per_rank"])
    c3=d.get("c3")
    if c3: print("   c3 value", round(c3["value"]), "e2e", round(c3["e2e"]["value"]), "ms", round(c3["ms_per_step"],2), round(c3["e2e"]["ms_per_step"],2), "parity", c3["parity"]["mismatches"], c3["parity"].get("distance_mismatches"), "ranks", c3["kernel_ms_per_rank"])
'
( timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 20 --warmup 5 ; echo "exit $?" ) > gpurun_out/r2l_bench_8gpu.log 2>&1
python -c "$summ" < gpurun_out/r2l_bench_8gpu.log; tail -n 1 gpurun_out/r2l_bench_8gpu.log
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 8 --workload c5 --no-cpu ; echo "exit $?" ) > gpurun_out/r2l_bench_8gpu_c5.log 2>&1
python -c "$summ" < gpurun_out/r2l_bench_8gpu_c5.log; tail -n 3 gpurun_out/r2l_bench_8gpu_c5.log | cut -c1-300
( timeout 500 python bench.py --gpus 8 --inprocess --steps 20 --warmup 5 ; echo "exit $?" ) > gpurun_out/r2l_bench_8gpu_inprocess.log 2>&1
python -c "$summ" < gpurun_out/r2l_bench_8gpu_inprocess.log; tail -n 1 gpurun_out/r2l_bench_8gpu_inprocess.log
( timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29524 bench.py --gpus 4 --steps 20 --warmup 5 ; echo "exit $?" ) > gpurun_out/r2l_bench_4gpu.log 2>&1
python -c "$summ" < gpurun_out/r2l_bench_4gpu.log; tail -n 1 gpurun_out/r2l_bench_4gpu.log
