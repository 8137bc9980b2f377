#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/g2_gpus.txt
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_backbone.py -q -m gpu --timeout 600 -k "two_gpus or two_devices" > gpurun_out/g2_pytest.log 2>&1; echo "pytest rc=$?"
tail -c 1500 gpurun_out/g2_pytest.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/g2_bench.json 2> gpurun_out/g2_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/g2_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/g2_bench.json").read().strip().splitlines()[-1])
for k in ("value","e2e","c3_bucketed","c4_train_step","extras_seconds"):
    print(k, json.dumps(d.get(k))[:700])
PY
