"""Experiment: forward of B=64 as one batch vs. two half-batches on two streams (HBM-bound passes of one half overlap the
tensor-bound convolutions of the other). Prints ms per 64 lines for each schedule."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import hctr_b200
from hctr_b200.models.handwritten_ctr_model import hctr_model

torch.manual_seed(1234)
m = hctr_model(7375).cuda().eval()
m.logits_dtype = torch.bfloat16
B, W = 64, 2048
x = (torch.rand(B, 1, 128, W, device="cuda") * 2 - 1)

def timeit(fn, n=5, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

def whole():
    with torch.no_grad():
        return m(x)

streams = [torch.cuda.Stream() for _ in range(4)]
def split(k):
    def run():
        cur = torch.cuda.current_stream()
        outs = []
        chunk = B // k
        for i in range(k):
            s = streams[i]
            s.wait_stream(cur)
            with torch.cuda.stream(s), torch.no_grad():
                outs.append(m(x[i * chunk:(i + 1) * chunk]))
        for i in range(k):
            cur.wait_stream(streams[i])
        return outs
    return run

print("whole      : %.2f ms" % timeit(whole), flush=True)
for k in (2, 4):
    print("split x%d   : %.2f ms" % (k, timeit(split(k))), flush=True)
print("whole again: %.2f ms" % timeit(whole), flush=True)
a = whole(); b = torch.cat(split(2)(), dim=1); torch.cuda.synchronize()
print("bit-identical:", bool((a == b).all()))
