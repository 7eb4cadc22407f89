#!/bin/bash
# Round 2 (1 GPU): last validation of the round -- the whole GPU test tier, alignment timings with the staged per-level copies
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -q --timeout 600 ; echo "exit $?" ) > gpurun_out/r3m_tests.log 2>&1
tail -n 3 gpurun_out/r3m_tests.log
( timeout 300 python tools/prof_msa3.py ; echo "exit $?" ) > gpurun_out/r3m_msa.log 2>&1
cat gpurun_out/r3m_msa.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
