#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_backbone.py -q -m gpu --timeout 300 -k "graph or greedy or logits" > gpurun_out/g_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 15 gpurun_out/g_pytest.log | cut -c1-250
timeout 300 python - <<'PY'
import sys, json, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import hctr_b200, bench_extras as bx
from hctr_b200.models.handwritten_ctr_model import hctr_model
from hctr_b200.utils.ctc_codec import ctc_codec
import synth
dev = torch.device("cuda:0")
torch.manual_seed(1234)
m = hctr_model(7375).to(dev).eval(); m.logits_dtype = torch.bfloat16
codec = ctc_codec(synth.charset(7373))
print(json.dumps(bx.b1_latency_leg(m, codec, dev), indent=1))
PY
