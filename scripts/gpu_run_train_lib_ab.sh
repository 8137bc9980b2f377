#!/bin/bash
# same-box A/B of the training step: the library as built vs csrc/_build/libhctr_old.so, after the training parity tests
mkdir -p gpurun_out
L=handwritten-chinese-ocr-samples_b200/libhctr_b200.so
timeout 600 python -m pytest tests/test_gpu_train_kernels.py tests/test_gpu_train.py -q -m gpu --timeout 300 > gpurun_out/t_pytest.log 2>&1; echo "pytest rc=$?"
tail -n 3 gpurun_out/t_pytest.log
cp $L /tmp/new.so
for rep in 1 2; do
for v in new old; do
  if [ $v = old ]; then cp handwritten-chinese-ocr-samples_b200/csrc/_build/libhctr_old.so $L; else cp /tmp/new.so $L; fi
  for n in 2 16; do
    timeout 300 python scripts/train_bench.py --lines-per-gpu $n --steps 10 --warmup 4 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v lines', d['lines_per_gpu'], 'ms', round(d['ms_per_step'],3))"
  done
done
done
cp /tmp/new.so $L
