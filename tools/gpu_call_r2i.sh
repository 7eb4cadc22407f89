#!/bin/bash
# Round 2, ninth call (1 GPU): register-tiled MSA sweep -- parity (oracle) with it and without, timings A/B in one box.
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_msa.py tests/test_msa_fuzz.py tests/test_zz_aligner_cli.py tests/test_multi_device.py -m gpu -q --timeout 300 ; echo "exit $?" ) > gpurun_out/r2i_tests_wave.log 2>&1
tail -n 6 gpurun_out/r2i_tests_wave.log
( TSQ_MSA_NO_WAVE=1 timeout 600 python -m pytest tests/test_msa.py tests/test_msa_fuzz.py -m gpu -q --timeout 300 ; echo "exit $?" ) > gpurun_out/r2i_tests_nowave.log 2>&1
tail -n 3 gpurun_out/r2i_tests_nowave.log
echo "== with the register-tiled sweep"; ( timeout 200 python tools/prof_msa.py 2>&1 | tail -n 12 ) | tee gpurun_out/r2i_msa_wave.txt
echo "== TSQ_MSA_NO_WAVE=1"; ( TSQ_MSA_NO_WAVE=1 timeout 200 python tools/prof_msa.py 2>&1 | tail -n 12 ) | tee gpurun_out/r2i_msa_nowave.txt
