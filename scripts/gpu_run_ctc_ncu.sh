#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/r2_ncu_launches_ctc_loss_B16_relative.csv python scripts/ctc_bench.py 16 > gpurun_out/c_ncu.log 2>&1; echo rc=$?
python scripts/launch_summary.py gpurun_out/r2_ncu_launches_ctc_loss_B16_relative.csv 20
