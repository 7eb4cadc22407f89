#!/bin/bash
# Round 2 (1 GPU): the one-warp walk-back refactored into pieces the CPU emulation shares -- alignment tests, timings
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_msa.py tests/test_zz_aligner_cli.py -m gpu -q --timeout 300 -x ; echo "exit $?" ) > gpurun_out/r3s_tests.log 2>&1
tail -n 2 gpurun_out/r3s_tests.log
( timeout 300 python tools/prof_msa3.py ; echo "exit $?" ) > gpurun_out/r3s_msa.log 2>&1
cat gpurun_out/r3s_msa.log
