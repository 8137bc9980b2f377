#!/bin/bash
# round-2 GPU batch A: parity (each file in its own process), micro A/B of the codec passes, then the bench line
mkdir -p gpurun_out; rm -f gpurun_out/a_rc.txt
timeout 600 python -m pytest tests/test_gpu_ctc_loss.py -q -m gpu --timeout 300 > gpurun_out/a_pytest_ctc.log 2>&1; echo "ctc rc=$?" >> gpurun_out/a_rc.txt
timeout 600 python -m pytest tests/test_gpu_codec.py -q -m gpu --timeout 300 > gpurun_out/a_pytest_codec.log 2>&1; echo "codec rc=$?" >> gpurun_out/a_rc.txt
timeout 900 python -m pytest tests/test_gpu_backbone.py tests/test_gpu_train_kernels.py tests/test_gpu_train.py tests/test_pipeline.py tests/test_resize.py -q -m gpu --timeout 300 > gpurun_out/a_pytest_rest.log 2>&1; echo "rest rc=$?" >> gpurun_out/a_rc.txt
python scripts/codec_micro.py > gpurun_out/a_micro.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; echo "bench rc=$?" >> gpurun_out/a_rc.txt
tail -c 600 gpurun_out/a_pytest_ctc.log; tail -c 400 gpurun_out/a_pytest_codec.log; tail -c 800 gpurun_out/a_pytest_rest.log; cat gpurun_out/a_rc.txt
python - <<'PY'
import json
d=json.load(open("gpurun_out/codec_micro.json"))
for k,v in d.items():
    print(k, {a:(round(b["ms"],3), round(b.get("frac_hbm",b.get("frac",0)),3)) for a,b in v.items()})
PY
