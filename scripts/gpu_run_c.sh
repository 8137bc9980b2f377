#!/bin/bash
# ncu --set full captures of the round-2 codec kernels (one kernel each), + launch lists
mkdir -p gpurun_out
cat > /tmp/topk_one.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import torch, hctr_b200
from hctr_b200 import native as nat
import bench_extras as bx
lib = nat.lib(); dev = torch.device("cuda", 0)
T, B, C, k = 512, 256, 7375, 10
x = bx.beam_logits_device(T, B, C, 0, dev)
if len(sys.argv) > 1: x = x.to(torch.bfloat16)
code = nat.HCTR_BF16 if len(sys.argv) > 1 else nat.HCTR_F32
ti = torch.empty((T, B, k), dtype=torch.int32, device=dev); tp = torch.empty((T, B, k), dtype=torch.float32, device=dev)
lse = torch.empty((T, B), dtype=torch.float32, device=dev)
for _ in range(3):
    nat.check(lib.hctr_ctc_topk_logsoftmax(nat.ptr(x), code, T, B, C, x.stride(0), x.stride(1), k, nat.ptr(ti), nat.ptr(tp), nat.ptr(lse), nat.stream_ptr()))
torch.cuda.synchronize()
PY
ncu --set full --clock-control none --import-source on -k regex:ctc_topk -s 2 -c 1 -o gpurun_out/r2_topk_f32 -f python /tmp/topk_one.py > gpurun_out/c_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ctc_topk -s 2 -c 1 -o gpurun_out/r2_topk_bf16 -f python /tmp/topk_one.py bf16 > gpurun_out/c_ncu2.log 2>&1
HCTR_CTC_OVERLAP=2 ncu --set full --clock-control none --import-source on -k regex:ctc_rows_kernel -s 3 -c 1 -o gpurun_out/r2_ctc_rows_bf16_B16 -f python scripts/ctc_bench.py 16 > gpurun_out/c_ncu3.log 2>&1
HCTR_CTC_OVERLAP=2 ncu --set full --clock-control none --import-source on -k regex:ctc_fix_kernel -s 3 -c 1 -o gpurun_out/r2_ctc_fix_bf16_B16 -f python scripts/ctc_bench.py 16 > gpurun_out/c_ncu4.log 2>&1
HCTR_CTC_OVERLAP=2 ncu --set full --clock-control none --import-source on -k regex:ctc_scan_kernel -s 3 -c 1 -o gpurun_out/r2_ctc_scan_B16 -f python scripts/ctc_bench.py 16 > gpurun_out/c_ncu5.log 2>&1
HCTR_CTC_OVERLAP=2 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/r2_ncu_launches_ctc_loss_B16.csv python scripts/ctc_bench.py 16 > gpurun_out/c_ncu6.log 2>&1
HCTR_CTC_OVERLAP=2 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/r2_ncu_launches_ctc_loss_B64.csv python scripts/ctc_bench.py 64 > gpurun_out/c_ncu7.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -8
