#!/bin/bash
# One gpurun call: the whole GPU test tier, the soak, the bench line (both arms) and the MSA timings.
mkdir -p gpurun_out
( timeout 500 python -m pytest tests -m gpu -x -q ; echo "exit $?" ) > gpurun_out/full_tests.log 2>&1
tail -n 3 gpurun_out/full_tests.log
( timeout 120 python tools/soak.py --seconds 60 --seed 9000 ; echo "exit $?" ) > gpurun_out/soak.log 2>&1
tail -n 2 gpurun_out/soak.log
( timeout 120 python tools/prof_msa.py ; echo "exit $?" ) > gpurun_out/msa_prof.log 2>&1
cat gpurun_out/msa_prof.log
( timeout 240 python bench.py ; echo "exit $?" ) > gpurun_out/bench_1gpu.log 2>&1
tail -n 2 gpurun_out/bench_1gpu.log | cut -c1-600
( timeout 200 python bench.py --impl reference --steps 2 --warmup 1 ; echo "exit $?" ) > gpurun_out/bench_ref.log 2>&1
tail -n 2 gpurun_out/bench_ref.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -n 1 gpurun_out/smoke.log
