#!/bin/bash
# training parity + the config-4 step at 2 and 16 lines per GPU (one GPU) + warm launch list of one 2-line step
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train_kernels.py tests/test_gpu_train.py -q -m gpu --timeout 300 > gpurun_out/t_pytest.log 2>&1; echo "pytest rc=$?"
tail -n 5 gpurun_out/t_pytest.log
timeout 300 python scripts/train_bench.py --lines-per-gpu 2 --steps 10 --warmup 4 > gpurun_out/t_b2.json 2> gpurun_out/t_b2.err; echo "b2 rc=$?"; cat gpurun_out/t_b2.json
timeout 300 python scripts/train_bench.py --lines-per-gpu 16 --steps 5 --warmup 3 > gpurun_out/t_b16.json 2> gpurun_out/t_b16.err; echo "b16 rc=$?"; cat gpurun_out/t_b16.json
if [ -n "$TRAIN_NCU" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/t_launches_b2.csv python scripts/profile_train_step.py > gpurun_out/t_ncu_b2.log 2>&1; echo "ncu rc=$?"
python scripts/launch_summary.py gpurun_out/t_launches_b2.csv 45
fi
