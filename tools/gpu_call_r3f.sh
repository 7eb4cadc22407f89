#!/bin/bash
# Round 2 (1 GPU): compute-sanitizer over every kernel again after the round's late work (results streamed from one
# launch on per-range counters, wave16 step rework, tiled alignment sweep with worker warps and the warp walk-back).
mkdir -p gpurun_out
for tool in memcheck racecheck initcheck; do
  ( timeout 900 compute-sanitizer --tool $tool python tools/sanitize_run.py ; echo "exit $?" ) > gpurun_out/r3f_sanitize_$tool.log 2>&1
  echo "--tool $tool (every kernel)"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize_run ok|exit" gpurun_out/r3f_sanitize_$tool.log | tail -n 3
  ( timeout 300 compute-sanitizer --tool $tool python tools/sanitize_msa.py ; echo "exit $?" ) > gpurun_out/r3f_sanitize_msa_$tool.log 2>&1
  echo "--tool $tool (alignment)"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize_msa ok|exit" gpurun_out/r3f_sanitize_msa_$tool.log | tail -n 3
done
