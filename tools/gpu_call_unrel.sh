#!/bin/bash
mkdir -p gpurun_out
export TSQ_MSA_DEBUG=1
python tools/prof_msa_unrel.py 2 2>&1 | tee gpurun_out/msa_unrel.log
python tools/prof_msa.py 2>&1 | tee -a gpurun_out/msa_unrel.log
