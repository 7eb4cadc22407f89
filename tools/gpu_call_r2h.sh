#!/bin/bash
# Round 2, eighth call (2 GPUs): GPU tier, bench N=1 and N=2 with the e2e stage split (threaded encode, single-write FASTA).
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q --timeout 400 ; echo "exit $?" ) > gpurun_out/r2h_tests.log 2>&1
tail -n 6 gpurun_out/r2h_tests.log
summ='
import json,sys
for l in sys.stdin:
    if not l.startswith("{\"metric\""): continue
    d=json.loads(l)
    print(d["config"]["launch"], "| N", d["n_gpus"], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"],3), round(d["e2e"]["ms_per_step"],3), d["e2e"].get("stages_ms_rank0"), "parity", d["parity"]["mismatches"])
    c3=d.get("c3")
    if c3: print("   c3 value", round(c3["value"]), "e2e", round(c3["e2e"]["value"]), "ms", round(c3["ms_per_step"],2), round(c3["e2e"]["ms_per_step"],2), c3["e2e"].get("stages_ms_rank0"), "parity", c3["parity"]["mismatches"])
    if "e2e_plugin" in d: print("   plugin", {k:(v["wall_ms"], v["stages_ms"]) for k,v in d["e2e_plugin"].items() if isinstance(v,dict) and "wall_ms" in v})
'
( timeout 400 python bench.py ; echo "exit $?" ) > gpurun_out/r2h_bench_1gpu.log 2>&1
python -c "$summ" < gpurun_out/r2h_bench_1gpu.log; tail -n 1 gpurun_out/r2h_bench_1gpu.log
( timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 ; echo "exit $?" ) > gpurun_out/r2h_bench_2gpu.log 2>&1
python -c "$summ" < gpurun_out/r2h_bench_2gpu.log; tail -n 1 gpurun_out/r2h_bench_2gpu.log
( timeout 400 python bench.py --gpus 2 --inprocess --steps 20 --warmup 5 ; echo "exit $?" ) > gpurun_out/r2h_bench_2gpu_inprocess.log 2>&1
python -c "$summ" < gpurun_out/r2h_bench_2gpu_inprocess.log; tail -n 1 gpurun_out/r2h_bench_2gpu_inprocess.log
