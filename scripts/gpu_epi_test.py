"""Isolate the cost of the residual read in the conv epilogue: hctr_conv_dgrad with and without `add` (no gate)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import hctr_b200
from hctr_b200 import native as nat
lib = nat.lib(); dev = "cuda"
def timeit(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
B, W = 64, 2048
for (H, C) in [(32, 256), (16, 512)]:
    x = torch.randn(B, H, W, C, device=dev).to(torch.bfloat16); wp = (torch.randn(C, 3, 3, C, device=dev) / (3 * C ** 0.5)).to(torch.bfloat16)
    ones = torch.ones(C, device=dev); zeros = torch.zeros(C, device=dev); add = torch.randn(B, H, W, C, device=dev).to(torch.bfloat16)
    gate = torch.rand(B, C, device=dev)
    y = torch.empty(B, H, W, C, dtype=torch.bfloat16, device=dev)
    fl = 2.0 * B * H * W * C * C * 9
    ms = timeit(lambda: nat.check(lib.hctr_conv_dgrad(nat.ptr(x), nat.ptr(wp), nat.ptr(ones), nat.ptr(zeros), None, nat.ptr(y), B, H, W, C, C, 3, nat.stream_ptr())))
    print("H%d C%d no add        : %.3f ms %.0f TFLOP/s" % (H, C, ms, fl / ms / 1e9), flush=True)
    ms = timeit(lambda: nat.check(lib.hctr_conv_dgrad(nat.ptr(x), nat.ptr(wp), nat.ptr(ones), nat.ptr(zeros), nat.ptr(add), nat.ptr(y), B, H, W, C, C, 3, nat.stream_ptr())))
    print("H%d C%d add           : %.3f ms %.0f TFLOP/s" % (H, C, ms, fl / ms / 1e9), flush=True)
    ms = timeit(lambda: nat.check(lib.hctr_conv_bn_gate_res_fwd(nat.ptr(x), nat.ptr(wp), nat.ptr(ones), nat.ptr(zeros), nat.ptr(gate), nat.ptr(add), nat.ptr(y), B, H, W, C, C, 3, 1, nat.stream_ptr())))
    print("H%d C%d gate+add+relu : %.3f ms %.0f TFLOP/s" % (H, C, ms, fl / ms / 1e9), flush=True)
