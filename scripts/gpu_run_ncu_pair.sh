#!/bin/bash
# ncu --set full (with source) of one wide 512->512 pair-kernel launch of the B=64 step (gate + residual variant and plain variant)
mkdir -p gpurun_out
N="ncu --set full --clock-control none --import-source on --profile-from-start off"
timeout 300 $N -k regex:igemm_pair_kernel -s 14 -c 2 -o gpurun_out/r2_pair_512 -f python scripts/profile_step.py > gpurun_out/ncu_p.log 2>&1; echo "rc=$?"
