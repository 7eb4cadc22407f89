#!/bin/bash
# Round 2 (1 GPU): per-launch trace of the progressive alignment on configs[1] (where do the 30 ms go)
mkdir -p gpurun_out
( TSQ_MSA_DEBUG=2 timeout 300 python tools/prof_msa_c2.py ; echo "exit $?" ) > gpurun_out/r2y_msa_c2.log 2>&1
grep -c level gpurun_out/r2y_msa_c2.log; awk '/^run 1/{on=1} on' gpurun_out/r2y_msa_c2.log | cut -c1-160 | tail -50
