#!/bin/bash
mkdir -p gpurun_out
PROFILE_B=16 timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:wgrad_tcgen05_kernel -s 6 -c 1 -o gpurun_out/r2_wgrad_512_B16 -f python scripts/profile_train_step.py > gpurun_out/ncu_wgrad.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/ncu_wgrad.log
