#!/bin/bash
mkdir -p gpurun_out
for c in 1 2 3 4; do echo "cells_per_thread=$c"; TSQ_MSA_CELLS_PER_THREAD=$c python tools/prof_msa_one.py; TSQ_MSA_CELLS_PER_THREAD=$c python tools/prof_msa_one.py; done 2>&1 | tee gpurun_out/msa_exp.log
