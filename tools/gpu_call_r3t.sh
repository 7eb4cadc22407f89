#!/bin/bash
# Round 2 (1 GPU): the published alignment scores through the library
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 200 -k "published or known or closed" ; echo "exit $?" ) > gpurun_out/r3t_tests.log 2>&1
tail -n 3 gpurun_out/r3t_tests.log
