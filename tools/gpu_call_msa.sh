#!/bin/bash
# One gpurun call: the progressive-alignment tests first, then sanitizer, timings, and an ncu capture.
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_msa.py tests/test_host_cpp.py -m gpu -x -q ; echo "exit $?" ) > gpurun_out/msa_tests.log 2>&1
( timeout 120 compute-sanitizer --tool memcheck python tools/sanitize_msa.py ; echo "exit $?" ) > gpurun_out/msa_memcheck.log 2>&1
( timeout 120 compute-sanitizer --tool initcheck python tools/sanitize_msa.py ; echo "exit $?" ) > gpurun_out/msa_initcheck.log 2>&1
( timeout 120 compute-sanitizer --tool racecheck python tools/sanitize_msa.py ; echo "exit $?" ) > gpurun_out/msa_racecheck.log 2>&1
( TSQ_MSA_DEBUG=1 timeout 180 python tools/prof_msa.py ; echo "exit $?" ) > gpurun_out/msa_prof.log 2>&1
for f in msa_tests msa_memcheck msa_initcheck msa_racecheck; do tail -n 3 gpurun_out/$f.log; done
cat gpurun_out/msa_prof.log
( timeout 200 ncu --set full --clock-control none --import-source on -k regex:msa_merge -s 60 -c 2 -f -o gpurun_out/msa_merge python tools/prof_msa_one.py ; echo "exit $?" ) > gpurun_out/msa_ncu.log 2>&1
tail -n 2 gpurun_out/msa_ncu.log
