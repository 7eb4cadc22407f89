#!/bin/bash
# Round 2 (1 GPU): where a merge of the progressive alignment spends its time (ncu --set full with source counters on
# one mid-tree merge of a 100-sequence family), and the timings after removing a dynamically indexed register array.
mkdir -p gpurun_out
( TSQ_MSA_DEBUG=1 timeout 600 python tools/prof_msa.py ; echo "exit $?" ) > gpurun_out/r2u_msa.log 2>&1
grep -v "^tsq_msa" gpurun_out/r2u_msa.log | cut -c1-200
( timeout 300 ncu --set full --clock-control none --import-source on -k regex:msa_merge -s 60 -c 1 -f -o gpurun_out/r2u_msa_merge python tools/prof_msa_one.py ; echo "exit $?" ) > gpurun_out/r2u_ncu.log 2>&1
tail -n 2 gpurun_out/r2u_ncu.log
