#!/bin/bash
# Round 2 (1 GPU): randomised soak of the whole path after the late round-2 work (every trial against the oracle)
mkdir -p gpurun_out
( timeout 500 python tools/soak.py --seconds 200 --seed 31000 ; echo "exit $?" ) > gpurun_out/r3g_soak.log 2>&1
tail -n 4 gpurun_out/r3g_soak.log | cut -c1-700
