#!/bin/bash
# ncu --set full of the train-mode pointwise passes at 16 lines per GPU (a 512-channel h16 unit: launch index chosen inside stage 3)
mkdir -p gpurun_out
N="ncu --set full --clock-control none --import-source on --profile-from-start off"
for k in train_apply_fwd_kernel train_bwd_apply_kernel train_bwd_reduce_kernel; do
  PROFILE_B=16 timeout 300 $N -k regex:$k -s 20 -c 1 -o gpurun_out/r2_$k -f python scripts/profile_train_step.py > gpurun_out/ncu_pw_$k.log 2>&1; echo "$k rc=$?"
done
