#!/bin/bash
# usage: bash scripts/gpu_run_ngpu.sh N  -- bench.py at N GPUs through torchrun, as the driver launches it
N=$1
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/g${N}_bench.json 2> gpurun_out/g${N}_bench.err; echo "bench rc=$?"
tail -c 400 gpurun_out/g${N}_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/g${N}_bench.json").read().strip().splitlines()[-1])
for k in ("value","e2e","c3_bucketed","c4_train_step","extras_seconds","clocks"):
    print(k, json.dumps(d.get(k))[:600])
PY
