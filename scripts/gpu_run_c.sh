#!/bin/bash
mkdir -p gpurun_out
python scripts/codec_micro.py > gpurun_out/b_micro.log 2>&1
python - <<'PY'
import json
d=json.load(open("gpurun_out/codec_micro.json"))
for k,v in d.items():
    print(k, {a:(round(b["ms"],3), round(b.get("frac_hbm",b.get("frac",0)),3)) for a,b in v.items()})
PY
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
for _ in range(2):
    nat.check(lib.hctr_ctc_topk_logsoftmax(nat.ptr(x), code, T, B, C, x.stride(0), x.stride(1), k, nat.ptr(ti), nat.ptr(tp), nat.ptr(lse), nat.stream_ptr()))
torch.cuda.synchronize()
PY
ncu --set full --clock-control none --import-source on -k regex:ctc_topk_reg -c 1 -o gpurun_out/c_topk_reg -f python /tmp/topk_one.py > gpurun_out/c_ncu_topk.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ctc_topk_reg -c 1 -o gpurun_out/c_topk_reg_bf16 -f python /tmp/topk_one.py bf16 > gpurun_out/c_ncu_topk2.log 2>&1
HCTR_CTC_OVERLAP=2 ncu --set full --clock-control none --import-source on -k regex:ctc_rows_kernel -s 3 -c 1 -o gpurun_out/c_ctc_rows -f python scripts/ctc_bench.py 16 > gpurun_out/c_ncu_rows.log 2>&1
ls -la gpurun_out/*.ncu-rep
