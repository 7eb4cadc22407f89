#!/bin/bash
# Round 2 (1 GPU): the GPU test tier and smoke on the round's last commit
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -q --timeout 600 ; echo "exit $?" ) > gpurun_out/r3r_tests.log 2>&1
tail -n 3 gpurun_out/r3r_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
