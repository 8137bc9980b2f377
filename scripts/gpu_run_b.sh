#!/bin/bash
mkdir -p gpurun_out
python scripts/codec_micro.py > gpurun_out/b_micro.log 2>&1
timeout 600 python -m pytest tests/test_gpu_codec.py -q -m gpu --timeout 300 -k topk > gpurun_out/b_pytest_topk.log 2>&1; echo "topk rc=$?"
timeout 600 python -m pytest tests/test_gpu_ctc_loss.py -q -m gpu --timeout 300 > gpurun_out/b_pytest_ctc.log 2>&1; echo "ctc rc=$?"
tail -5 gpurun_out/b_pytest_topk.log; tail -5 gpurun_out/b_pytest_ctc.log
HCTR_CTC_OVERLAP=2 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/b_ctc_launches.csv python scripts/ctc_bench.py 16 > gpurun_out/b_ncu.log 2>&1
python scripts/launch_summary.py gpurun_out/b_ctc_launches.csv 20
python - <<'PY'
import json
d=json.load(open("gpurun_out/codec_micro.json"))
for k,v in d.items():
    print(k, {a:(round(b["ms"],3), round(b.get("frac_hbm",b.get("frac",0)),3)) for a,b in v.items()})
PY
