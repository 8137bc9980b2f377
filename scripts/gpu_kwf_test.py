import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, torch.nn.functional as F
import hctr_b200
from hctr_b200 import native as nat
torch.backends.cuda.matmul.allow_tf32 = False; torch.backends.cudnn.allow_tf32 = False
lib = nat.lib(); dev = "cuda"
def case(B, H, W, Cin, Cout, relu, pool):
    g = torch.Generator().manual_seed(B + W + Cin)
    x = torch.randn(B, Cin, H, W, generator=g).to(dev).to(torch.bfloat16)
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) / (Cin * 9) ** 0.5).to(dev).to(torch.bfloat16)
    scale = (torch.rand(Cout, generator=g) + 0.5).to(dev); shift = (0.1 * torch.randn(Cout, generator=g)).to(dev)
    xn = x.permute(0, 2, 3, 1).contiguous(); wp = w.permute(0, 2, 3, 1).contiguous()
    ref = F.conv2d(x.float(), w.float(), padding=1) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
    if relu: ref = ref.relu()
    if pool: ref = F.max_pool2d(ref, (2, 1), (2, 1))
    out = []
    for mode in (0, 1, 2):
        lib.hctr_debug_set_kwf_mode(mode)
        y = torch.full((B, H // 2 if pool else H, W, Cout), float("nan"), dtype=torch.bfloat16, device=dev)
        nat.check(lib.hctr_conv_bn_act_fwd(nat.ptr(xn), nat.ptr(wp), nat.ptr(scale), nat.ptr(shift), nat.ptr(y), B, H, W, Cin, Cout, 3, relu, pool, nat.stream_ptr()))
        torch.cuda.synchronize()
        err = (y.permute(0, 3, 1, 2).float() - ref).abs().max().item() / ref.abs().max().item()
        out.append(err)
    print("B%d H%d W%d %d->%d r%d p%d: rel err mode0 %.2e  mode1(base_offset) %.2e  mode2(no base_offset) %.2e" % (B, H, W, Cin, Cout, relu, pool, *out), flush=True)
for c in [(1, 4, 128, 64, 64, 1, 0), (2, 8, 256, 64, 64, 1, 1), (2, 8, 200, 64, 128, 1, 0), (1, 6, 96, 128, 128, 0, 0), (2, 4, 384, 128, 128, 1, 1)]:
    case(*c)
# timing of the thin layers at bench size
def timeit(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
for (H, Cin, Cout) in [(128, 64, 64), (64, 64, 128), (64, 128, 128)]:
    B, W = 64, 2048
    xn = torch.randn(B, H, W, Cin, device=dev).to(torch.bfloat16); wp = torch.randn(Cout, 3, 3, Cin, device=dev).to(torch.bfloat16)
    sc = torch.ones(Cout, device=dev); sh = torch.zeros(Cout, device=dev); y = torch.empty(B, H, W, Cout, dtype=torch.bfloat16, device=dev)
    for mode in (0, 1):
        lib.hctr_debug_set_kwf_mode(mode)
        ms = timeit(lambda: nat.check(lib.hctr_conv_bn_act_fwd(nat.ptr(xn), nat.ptr(wp), nat.ptr(sc), nat.ptr(sh), nat.ptr(y), B, H, W, Cin, Cout, 3, 1, 0, nat.stream_ptr())))
        print("H%d %d->%d mode %d: %.3f ms  %.0f TFLOP/s" % (H, Cin, Cout, mode, ms, 2.0 * B * H * W * Cout * Cin * 9 / ms / 1e9))
