#!/bin/bash
# Round 2 (1 GPU): progressive alignment with column-score worker warps -- parity, timings, per-launch trace on configs[1]
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_msa.py tests/test_zz_aligner_cli.py -m gpu -q --timeout 600 -x ; echo "exit $?" ) > gpurun_out/r3a_tests.log 2>&1
tail -n 3 gpurun_out/r3a_tests.log
( TSQ_MSA_DEBUG=1 timeout 600 python tools/prof_msa.py ; echo "exit $?" ) > gpurun_out/r3a_msa.log 2>&1
grep -v "^tsq_msa" gpurun_out/r3a_msa.log | cut -c1-200
( TSQ_MSA_DEBUG=2 timeout 300 python tools/prof_msa_c2.py ; echo "exit $?" ) > gpurun_out/r3a_msa_c2.log 2>&1
grep level gpurun_out/r3a_msa_c2.log | awk 'NR<=12 || NR%6==0' | cut -c1-150
( timeout 300 ncu --set full --clock-control none --import-source on -k regex:msa_merge -s 60 -c 1 -f -o gpurun_out/r3a_msa_merge python tools/prof_msa_one.py ; echo "exit $?" ) > gpurun_out/r3a_ncu.log 2>&1
tail -n 1 gpurun_out/r3a_ncu.log
