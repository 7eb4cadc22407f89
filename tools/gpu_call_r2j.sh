#!/bin/bash
# Round 2, tenth call (1 GPU): randomised soak of the whole path incl. the round-2 routes; compute-sanitizer over every kernel.
mkdir -p gpurun_out
( timeout 400 python tools/soak.py --seconds 150 --seed 20000 ; echo "exit $?" ) > gpurun_out/r2j_soak.log 2>&1
tail -n 3 gpurun_out/r2j_soak.log | cut -c1-600
for tool in memcheck racecheck initcheck; do
  ( timeout 900 compute-sanitizer --tool $tool python tools/sanitize_run.py ; echo "exit $?" ) > gpurun_out/r2j_sanitize_$tool.log 2>&1
  echo "--tool $tool"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize_run ok|exit" gpurun_out/r2j_sanitize_$tool.log | tail -n 3
done
