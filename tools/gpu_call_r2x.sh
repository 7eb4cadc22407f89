#!/bin/bash
# Round 2 (1 GPU): tiled sweep with the straight-line full-tile path, warp-cooperative walk-back, CTA-size policy -- parity, timings, ncu of one mid-tree merge.
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_msa.py tests/test_zz_aligner_cli.py -m gpu -q --timeout 600 -x ; echo "exit $?" ) > gpurun_out/r2x_tests.log 2>&1
tail -n 3 gpurun_out/r2x_tests.log
( TSQ_MSA_DEBUG=1 timeout 600 python tools/prof_msa.py ; echo "exit $?" ) > gpurun_out/r2x_msa.log 2>&1
grep -v "^tsq_msa" gpurun_out/r2x_msa.log | cut -c1-200
( timeout 300 ncu --set full --clock-control none --import-source on -k regex:msa_merge -s 60 -c 1 -f -o gpurun_out/r2x_msa_merge python tools/prof_msa_one.py ; echo "exit $?" ) > gpurun_out/r2x_ncu.log 2>&1
tail -n 2 gpurun_out/r2x_ncu.log
