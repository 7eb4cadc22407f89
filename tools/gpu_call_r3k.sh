#!/bin/bash
# Round 2 (1 GPU): TSQ_FLAG_SCORES_I16 on every delivery route, and the routes it shares code with
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_scores_i16.py tests/test_multi_device.py tests/test_gpu_parity.py -m gpu -q --timeout 600 ; echo "exit $?" ) > gpurun_out/r3k_tests.log 2>&1
tail -n 25 gpurun_out/r3k_tests.log | cut -c1-220
