#!/bin/bash
# records: (1) 1TDVP at chi = 256 (step time and kernel breakdown), (2) the sharded tile-pass kernels of the 2-, 4- and 8-rank
# geometries on ONE GPU with loop-back peers next to the unsharded kernel on the same slice size
mkdir -p gpurun_out
timeout 240 python scratch/tdvp_prof2.py 64 256 1tdvp > gpurun_out/r2n_tdvp1_prof.txt 2>&1; grep -v Warning gpurun_out/r2n_tdvp1_prof.txt | head -16; tail -1 gpurun_out/r2n_tdvp1_prof.txt
for w in 2 4 8; do timeout 120 python scratch/loopback_prof.py 30 $w 1 2>&1 | tail -2; done | tee gpurun_out/r2n_loopback.txt
