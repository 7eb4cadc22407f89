#!/bin/bash
# Round 2 (1 GPU): UPGMA with cluster sizes and node ids in shared memory -- guide-tree tests, timings
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_guide_tree.py tests/test_msa.py -m gpu -q --timeout 300 -x ; echo "exit $?" ) > gpurun_out/r3q_tests.log 2>&1
tail -n 2 gpurun_out/r3q_tests.log
( TSQ_UPGMA_THREADS= timeout 600 python tools/prof_tree2.py ; echo "exit $?" ) > gpurun_out/r3q_tree.log 2>&1
grep -E "1024|default|exit" gpurun_out/r3q_tree.log
