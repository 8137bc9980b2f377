#!/bin/bash
# ncu --set full (with source) of the thin single-CTA 3x3 launches of the B=64 step (launch order of igemm_tcgen05_kernel:
# conv0_2, block1.0.conv1 64->128, shortcut, block1.0.conv2 128->128 gate+res, block1.1.conv1 128->128 sum, ...)
mkdir -p gpurun_out
N="ncu --set full --clock-control none --import-source on --profile-from-start off"
timeout 300 $N -k regex:igemm_tcgen05_kernel -s 3 -c 2 -o gpurun_out/r2_thin_128 -f python scripts/profile_step.py > gpurun_out/ncu_t.log 2>&1; echo "rc=$?"
