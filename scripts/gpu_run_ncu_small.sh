#!/bin/bash
# ncu --set full of the 1x1 shortcut launches of the B=64 step (launch order: conv0_2, block1.0.conv1, block1.0 shortcut, ...)
mkdir -p gpurun_out
N="ncu --set full --clock-control none --import-source on --profile-from-start off"
timeout 300 $N -k regex:igemm_tcgen05_kernel -s 2 -c 1 -o gpurun_out/r2_conv1x1_64_128 -f python scripts/profile_step.py > gpurun_out/ncu_s2.log 2>&1; echo "rc=$?"
timeout 300 $N -k regex:igemm_pair_kernel -s 1 -c 1 -o gpurun_out/r2_conv1x1_pair -f python scripts/profile_step.py > gpurun_out/ncu_s3.log 2>&1; echo "rc=$?"
ls -la gpurun_out/*.ncu-rep | tail -4
