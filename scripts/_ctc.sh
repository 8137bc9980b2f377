python -m pytest tests/test_gpu_ctc_loss.py -x -q -m gpu 2>&1 | tail -2
for B in 4 8 16 24 32 64; do
for m in 4 2; do
echo "B=$B mode=$m $(HCTR_CTC_OVERLAP=$m python scripts/ctc_bench.py $B 2>&1 | grep '"ms"')"
done
done
python scripts/codec_micro.py 2>&1 | python -c "
import sys, json
t = sys.stdin.read(); d = json.loads(t[t.index('{'):])
for k, v in d.items():
    if k.startswith('ctc'): print(k, {a: (round(b['ms'], 3), round(b['frac'], 3)) for a, b in v.items()})
"
