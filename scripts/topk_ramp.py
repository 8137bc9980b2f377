"""Sensitivity of the bf16 top-k timing to OTHER live, touched allocations (found in round 2: 0.55 ms with the 3.87 GB fp32
tensor still allocated, 0.43 ms after it was freed).  HCTR_RAMP_N=iterations per state (2 under ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, hctr_b200
from hctr_b200 import native as nat
import bench_extras as bx
lib = nat.lib(); dev = torch.device("cuda", 0)
N = int(os.environ.get("HCTR_RAMP_N", "50"))
T, B, C, k = 512, 256, 7375, 10
ti = torch.empty((T, B, k), dtype=torch.int32, device=dev); tp = torch.empty((T, B, k), dtype=torch.float32, device=dev)
lse = torch.empty((T, B), dtype=torch.float32, device=dev)
def run(x):
    code = nat.HCTR_BF16 if x.dtype == torch.bfloat16 else nat.HCTR_F32
    fn = lambda: nat.check(lib.hctr_ctc_topk_logsoftmax(nat.ptr(x), code, T, B, C, x.stride(0), x.stride(1), k, nat.ptr(ti), nat.ptr(tp), nat.ptr(lse), nat.stream_ptr()))
    return round(bx._timeit(fn, N, max(1, N // 5)), 4)
def state(tag, x):
    print("%-44s %.4f ms   alloc %.1f GB reserved %.1f GB" % (tag, run(x), torch.cuda.memory_allocated() / 1e9, torch.cuda.memory_reserved() / 1e9), flush=True)
xb = bx.beam_logits_device(T, B, C, 0, dev).to(torch.bfloat16)
torch.cuda.empty_cache()
state("bf16 alone", xb)
for gb in (0.25, 0.5, 1, 2, 4, 8, 32):
    big = torch.empty(int(gb * 1e9), dtype=torch.uint8, device=dev); big.zero_()
    state("+ %.2f GB written" % gb, xb)
    del big; torch.cuda.empty_cache()
state("freed", xb)
big = torch.empty(int(4e9), dtype=torch.uint8, device=dev)
state("+ 4 GB untouched", xb)
s = big.view(torch.int32).sum().item()
state("+ 4 GB read only", xb)
del big
state("4 GB written earlier now only cached by torch", xb)
torch.cuda.empty_cache()
state("freed", xb)
xb2 = xb.clone()
state("second bf16 copy alive, run on the first", xb)
state("second bf16 copy alive, run on the second", xb2)
