#!/bin/bash
# ncu --set full captures of the round-2 codec kernels (one kernel each), + launch lists of the CTC loss call
# (default = split schedule at B = 16; HCTR_CTC_OVERLAP=2 = the one-pass rows kernel back to back with the scans)
mkdir -p gpurun_out
N="ncu --set full --clock-control none --import-source on"
$N -k regex:ctc_topk -s 2 -c 1 -o gpurun_out/r2_topk_chunk_f32 -f python scripts/topk_one.py > gpurun_out/c_ncu1.log 2>&1
$N -k regex:ctc_topk -s 2 -c 1 -o gpurun_out/r2_topk_chunk_bf16 -f python scripts/topk_one.py bf16 > gpurun_out/c_ncu2.log 2>&1
$N -k regex:ctc_lse_chunk_kernel -s 3 -c 1 -o gpurun_out/r2_ctc_lse_chunk_bf16_B16 -f python scripts/ctc_bench.py 16 > gpurun_out/c_ncu3.log 2>&1
$N -k regex:ctc_dense_grad_kernel -s 3 -c 1 -o gpurun_out/r2_ctc_dense_grad_bf16_B16 -f python scripts/ctc_bench.py 16 > gpurun_out/c_ncu4.log 2>&1
$N -k regex:ctc_fix_kernel -s 3 -c 1 -o gpurun_out/r2_ctc_fix_bf16_B16 -f python scripts/ctc_bench.py 16 > gpurun_out/c_ncu5.log 2>&1
$N -k regex:ctc_scan_kernel -s 3 -c 1 -o gpurun_out/r2_ctc_scan_B16 -f python scripts/ctc_bench.py 16 > gpurun_out/c_ncu6.log 2>&1
HCTR_CTC_OVERLAP=2 $N -k regex:ctc_rows_kernel -s 3 -c 1 -o gpurun_out/r2_ctc_rows_bf16_B16 -f python scripts/ctc_bench.py 16 > gpurun_out/c_ncu7.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/r2_ncu_launches_ctc_loss_B16.csv python scripts/ctc_bench.py 16 > gpurun_out/c_ncu8.log 2>&1
HCTR_CTC_OVERLAP=2 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/r2_ncu_launches_ctc_loss_B16_rows_schedule.csv python scripts/ctc_bench.py 16 > gpurun_out/c_ncu9.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/r2_ncu_launches_ctc_loss_B64.csv python scripts/ctc_bench.py 64 > gpurun_out/c_ncu10.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -8
