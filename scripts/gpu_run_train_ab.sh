#!/bin/bash
# training parity, then a same-box A/B of the training step over HCTR_TRAIN_FOLD_STATS (boxes of the pool differ by 2-3 %)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train_kernels.py tests/test_gpu_train.py -q -m gpu --timeout 300 > gpurun_out/t_pytest.log 2>&1; echo "pytest rc=$?"
tail -n 12 gpurun_out/t_pytest.log | cut -c1-300
for rep in 1 2; do
for v in 1 0 2; do
  for n in 2 16; do
    HCTR_TRAIN_FOLD_STATS=$v timeout 300 python scripts/train_bench.py --lines-per-gpu $n --steps 10 --warmup 4 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('fold=$v lines', d['lines_per_gpu'], 'ms', round(d['ms_per_step'],3), 'loss', d['loss'])"
  done
done
done
