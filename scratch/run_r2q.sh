#!/bin/bash
# ncu --set full of the persistent sharded tile-pass kernel (8-rank geometry, loop-back peers) on one GPU
mkdir -p gpurun_out
QCA_NCU=sharded timeout 100 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:pass_kernel_v2p --launch-skip 3 --launch-count 3 -f -o gpurun_out/r02_pass_v2p_loopback_n30_w8 python scratch/loopback_prof.py 30 8 1 > gpurun_out/r2q_ncu.log 2>&1; tail -3 gpurun_out/r2q_ncu.log | cut -c1-200
