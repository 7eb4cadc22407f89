#!/bin/bash
# Round 2 (1 GPU): tiled sweep of the progressive alignment (4 x 4 cells per thread and step) -- parity on the GPU,
# tsq_msa timings on the r01 workloads, the plugin call of the bench line (with the faster FASTA reader).
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_msa.py tests/test_zz_aligner_cli.py tests/test_guide_tree.py tests/test_traceback.py -m gpu -q --timeout 600 -x ; echo "exit $?" ) > gpurun_out/r2t_tests.log 2>&1
tail -n 4 gpurun_out/r2t_tests.log
( TSQ_MSA_DEBUG=1 timeout 600 python tools/prof_msa.py ; echo "exit $?" ) > gpurun_out/r2t_msa.log 2>&1
cat gpurun_out/r2t_msa.log | cut -c1-220
( timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-c3 ; echo "exit $?" ) > gpurun_out/r2t_bench.log 2>&1
grep '^{"metric"' gpurun_out/r2t_bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
for k,v in d['e2e_plugin'].items():
    print(k, v if isinstance(v,str) else (v['wall_ms'], v['stages_ms']))"
