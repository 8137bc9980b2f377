"""Time one wide convolution on the single-CTA kernel and on the CTA-pair kernel (cta_group::2)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
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
B, W = int(os.environ.get("PB", "64")), 2048
for (H, Cin, Cout) in [(16, 512, 512), (32, 256, 256)]:
    xn = torch.randn(B, H, W, Cin, device=dev).to(torch.bfloat16); wp = (torch.randn(Cout, 3, 3, Cin, device=dev) / (3 * Cin ** 0.5)).to(torch.bfloat16)
    sc = torch.ones(Cout, device=dev); sh = torch.zeros(Cout, device=dev)
    ys = []
    for mode in (0, 1):
        lib.hctr_debug_set_pair_mode(mode)
        y = torch.empty(B, H, W, Cout, dtype=torch.bfloat16, device=dev)
        ms = timeit(lambda: nat.check(lib.hctr_conv_bn_act_fwd(nat.ptr(xn), nat.ptr(wp), nat.ptr(sc), nat.ptr(sh), nat.ptr(y), B, H, W, Cin, Cout, 3, 1, 0, nat.stream_ptr())))
        print("B%d H%d %d->%d pair=%d: %.3f ms  %.0f TFLOP/s" % (B, H, Cin, Cout, mode, ms, 2.0 * B * H * W * Cout * Cin * 9 / ms / 1e9), flush=True)
        ys.append(y)
    print("   max |diff| between kernels:", (ys[0].float() - ys[1].float()).abs().max().item())
