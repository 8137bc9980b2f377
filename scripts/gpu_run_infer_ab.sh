#!/bin/bash
# same-box A/B of the inference step: the library as built vs csrc/_build/libhctr_old.so (+ parity of the new one first)
mkdir -p gpurun_out
L=handwritten-chinese-ocr-samples_b200/libhctr_b200.so
timeout 900 python -m pytest tests/test_gpu_backbone.py tests/test_gpu_train_kernels.py -q -m gpu --timeout 300 -x > gpurun_out/i_pytest.log 2>&1; echo "pytest rc=$?"
tail -n 4 gpurun_out/i_pytest.log | cut -c1-300
cp $L /tmp/new.so
for rep in 1 2; do
for v in new old; do
  if [ $v = old ]; then cp handwritten-chinese-ocr-samples_b200/csrc/_build/libhctr_old.so $L; else cp /tmp/new.so $L; fi
  timeout 600 python bench.py --steps 6 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null > gpurun_out/i_bench_${v}_$rep.json
  python - <<PY
import json
d=json.loads(open("gpurun_out/i_bench_${v}_$rep.json").read().strip().splitlines()[-1])
kb=d["kernel_breakdown"]
thin=sum(v["ms_per_step"] for k,v in kb.items() if isinstance(v,dict) and ("_64_" in k or "_128_128" in k or "64_128" in k or "128_256" in k or "conv1x1" in k))
print("$v", "lines/s %.1f  ms %.2f  thin+1x1 %.2f ms  cls %.2f  clocks %s" % (d["value"], d["ms_per_step"], thin, kb["classifier"]["ms_per_step"], d["clocks"]["sm_mhz"]))
print("   ", {k: round(v["ms_per_step"],2) for k,v in kb.items() if isinstance(v,dict) and ("conv1x1" in k or "_64_" in k or "128_128" in k)})
PY
done
done
cp /tmp/new.so $L
