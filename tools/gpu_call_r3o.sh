#!/bin/bash
# Round 2 (1 GPU): ncu --set full of the packed inter-task kernel on configs[1] with the round's final build
mkdir -p gpurun_out
( timeout 300 ncu --set full --clock-control none --import-source on -k regex:gotoh16 -c 1 -f -o gpurun_out/r3o_gotoh16_c2 python tools/prof_run.py c2 2 ; echo "exit $?" ) > gpurun_out/r3o_ncu_g16.log 2>&1
tail -n 2 gpurun_out/r3o_ncu_g16.log
