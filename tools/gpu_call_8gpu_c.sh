#!/bin/bash
# Round 2, 8 GPUs, third pass: the scaling lines on the final code of the round (ranges by task count with a shrinking tail, int16 route present),
# the vector encoder and the arrival flags in the shared result (no NCCL barrier at the end of a step).
mkdir -p gpurun_out
nvidia-smi -L | wc -l; nproc
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 tools/check_sharded.py ; echo "exit $?" ) > gpurun_out/r3l_check_sharded_8gpu.log 2>&1
grep -E "sharded ok|Error|exit" gpurun_out/r3l_check_sharded_8gpu.log | cut -c1-160 | tail -n 8
summ='
import json,sys
for l in sys.stdin:
    if not l.startswith("{\"metric\""): continue
    d=json.loads(l)
    print(d["config"]["launch"], "| N", d["n_gpus"], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"],3), round(d["e2e"]["ms_per_step"],3), d["e2e"]["stages_ms_rank0"], "parity", d["parity"]["mismatches"], d["parity"]["pairs_checked"])
    c3=d.get("c3")
    if c3: print("   c3 value", round(c3["value"]), "e2e", round(c3["e2e"]["value"]), "ms", round(c3["ms_per_step"],2), round(c3["e2e"]["ms_per_step"],2), c3["e2e"]["stages_ms_rank0"], "parity", c3["parity"]["mismatches"], c3["parity"].get("distance_mismatches"))
'
( timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 20 --warmup 5 ; echo "exit $?" ) > gpurun_out/r3l_bench_8gpu.log 2>&1
python -c "$summ" < gpurun_out/r3l_bench_8gpu.log; tail -n 1 gpurun_out/r3l_bench_8gpu.log
( timeout 500 python bench.py --gpus 8 --inprocess --steps 20 --warmup 5 ; echo "exit $?" ) > gpurun_out/r3l_bench_8gpu_inprocess.log 2>&1
python -c "$summ" < gpurun_out/r3l_bench_8gpu_inprocess.log; tail -n 1 gpurun_out/r3l_bench_8gpu_inprocess.log
( timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29524 bench.py --gpus 4 --steps 20 --warmup 5 ; echo "exit $?" ) > gpurun_out/r3l_bench_4gpu.log 2>&1
python -c "$summ" < gpurun_out/r3l_bench_4gpu.log; tail -n 1 gpurun_out/r3l_bench_4gpu.log
( timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-plugin ; echo "exit $?" ) > gpurun_out/r3l_bench_1gpu.log 2>&1
python -c "$summ" < gpurun_out/r3l_bench_1gpu.log; tail -n 1 gpurun_out/r3l_bench_1gpu.log
