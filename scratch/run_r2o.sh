#!/bin/bash
# multi-CTA Householder QR: TDVP parity suite (QR against numpy and against the one-CTA kernels, 1TDVP / 2TDVP fixtures), then 1TDVP at chi = 256
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_tdvp_gpu.py -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r2o_pytest_tdvp.log
timeout 200 python scratch/tdvp_prof2.py 64 256 1tdvp > gpurun_out/r2o_tdvp1_prof.txt 2>&1; grep -v "Warn\|warn" gpurun_out/r2o_tdvp1_prof.txt | head -12; tail -1 gpurun_out/r2o_tdvp1_prof.txt
