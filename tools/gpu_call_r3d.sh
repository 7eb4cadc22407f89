#!/bin/bash
# Round 2 (1 GPU): UPGMA CTA size sweep (tree identical at every size), guide-tree GPU tests
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_guide_tree.py -m gpu -q --timeout 300 -x ; echo "exit $?" ) > gpurun_out/r3d_tests.log 2>&1
tail -n 2 gpurun_out/r3d_tests.log
( timeout 600 python tools/prof_tree2.py ; echo "exit $?" ) > gpurun_out/r3d_tree.log 2>&1
cat gpurun_out/r3d_tree.log
