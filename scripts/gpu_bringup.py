"""First-contact check of every kernel against plain torch ops on the GPU box (not a pytest file)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
import hctr_b200
from hctr_b200 import native as nat

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
dev = torch.device("cuda:0")
lib = nat.lib()
print("device ok:", lib.hctr_device_supported(0), torch.cuda.get_device_name(0))
res = {}

def report(name, got, ref):
    err = (got.float() - ref.float()).abs().max().item()
    scale = ref.float().abs().max().item()
    print("%-40s max_abs_err %.4e  ref_absmax %.4e  rel %.3e" % (name, err, scale, err / max(scale, 1e-30)), flush=True)
    res[name] = err / max(scale, 1e-30)

def conv_case(B, H, W, Cin, Cout, k, relu, pool, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = torch.randn(B, Cin, H, W, generator=g).to(dev)
    w = (torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5).to(dev)
    scale = (torch.rand(Cout, generator=g) + 0.5).to(dev) * torch.where(torch.rand(Cout, generator=g) < 0.2, -1.0, 1.0).to(dev)
    shift = torch.randn(Cout, generator=g).to(dev) * 0.1
    xb = x.to(torch.bfloat16); wb = w.to(torch.bfloat16)
    x_nhwc = xb.permute(0, 2, 3, 1).contiguous()
    wp = wb.permute(0, 2, 3, 1).contiguous()
    Ho = H // 2 if pool else H
    y = torch.full((B, Ho, W, Cout), float("nan"), dtype=torch.bfloat16, device=dev)
    nat.check(lib.hctr_conv_bn_act_fwd(nat.ptr(x_nhwc), nat.ptr(wp), nat.ptr(scale), nat.ptr(shift), nat.ptr(y),
                                       B, H, W, Cin, Cout, k, int(relu), int(pool), nat.stream_ptr()), "conv")
    torch.cuda.synchronize()
    ref = F.conv2d(xb.float(), wb.float(), padding=k // 2) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
    if relu: ref = ref.relu()
    if pool: ref = F.max_pool2d(ref, (2, 1), (2, 1))
    report("conv B%d H%d W%d %d->%d k%d r%d p%d" % (B, H, W, Cin, Cout, k, relu, pool), y.permute(0, 3, 1, 2), ref)

try:
    conv_case(1, 4, 128, 64, 64, 3, 1, 0)
    conv_case(2, 8, 256, 64, 64, 3, 1, 1)
    conv_case(2, 8, 200, 64, 128, 3, 1, 0)
    conv_case(1, 6, 96, 128, 128, 3, 0, 0)
    conv_case(2, 4, 384, 128, 256, 3, 1, 1)
    conv_case(1, 4, 130, 256, 512, 3, 1, 0)
    conv_case(1, 4, 256, 512, 512, 3, 1, 1)
    conv_case(2, 8, 256, 64, 128, 1, 0, 0)
    conv_case(1, 4, 100, 256, 512, 1, 0, 0)
except Exception as e:
    print("CONV FAILED:", repr(e)); res["conv_exception"] = repr(e)

# classifier
try:
    for (B, W, N, dt) in [(2, 256, 7375, torch.float32), (1, 200, 7375, torch.bfloat16), (2, 96, 50, torch.float32)]:
        g = torch.Generator().manual_seed(1)
        feat = torch.randn(B, 4, W, 512, generator=g).to(dev).to(torch.bfloat16)
        w = (torch.randn(N, 2048, generator=g) / 45).to(dev).to(torch.bfloat16)
        bias = torch.randn(N, generator=g).to(dev)
        pitch = (N + 7) // 8 * 8
        out = torch.full((B, W, pitch), float("nan"), dtype=dt, device=dev)
        nat.check(lib.hctr_classifier_fwd(nat.ptr(feat), nat.ptr(w), nat.ptr(bias), nat.ptr(out),
                                          nat.HCTR_F32 if dt == torch.float32 else nat.HCTR_BF16, pitch, B, 4, W, 512, N,
                                          nat.stream_ptr()), "cls")
        torch.cuda.synchronize()
        a = feat.float().permute(0, 2, 1, 3).reshape(B * W, 2048)
        ref = (a @ w.float().t() + bias).reshape(B, W, N)
        report("classifier B%d W%d N%d %s" % (B, W, N, dt), out[:, :, :N], ref)
except Exception as e:
    print("CLS FAILED:", repr(e)); res["cls_exception"] = repr(e)

# stem
try:
    g = torch.Generator().manual_seed(2)
    B, H, W = 2, 128, 200
    x = (torch.rand(B, 1, H, W, generator=g) * 2 - 1).to(dev)
    w = torch.randn(64, 1, 3, 3, generator=g).to(dev) / 3
    scale = (torch.rand(64, generator=g) + 0.5).to(dev); shift = torch.randn(64, generator=g).to(dev) * 0.1
    y = torch.empty(B, H, W, 64, dtype=torch.bfloat16, device=dev)
    nat.check(lib.hctr_stem_conv_fwd(nat.ptr(x), nat.ptr(w.reshape(64, 9).contiguous()), nat.ptr(scale), nat.ptr(shift),
                                     nat.ptr(y), B, H, W, 1, nat.stream_ptr()), "stem")
    ref = (F.conv2d(x, w, padding=1) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)).relu()
    report("stem", y.permute(0, 3, 1, 2), ref)
    # SE
    C = 256; H2, W2 = 32, 200
    v = torch.randn(B, H2, W2, C, generator=g).to(dev).to(torch.bfloat16)
    r = torch.randn(B, H2, W2, C, generator=g).to(dev).to(torch.bfloat16)
    w1 = torch.randn(C // 16, C, generator=g).to(dev) / 16; w2 = torch.randn(C, C // 16, generator=g).to(dev) / 4
    slices = lib.hctr_se_slices(H2, W2)
    partial = torch.empty(B, slices, C, device=dev); gate = torch.empty(B, C, device=dev)
    nat.check(lib.hctr_se_squeeze(nat.ptr(v), nat.ptr(partial), B, H2, W2, C, nat.stream_ptr()))
    nat.check(lib.hctr_se_excite(nat.ptr(partial), slices, nat.ptr(w1), nat.ptr(w2), nat.ptr(gate), B, C, C // 16, H2 * W2, nat.stream_ptr()))
    out = torch.empty_like(v)
    nat.check(lib.hctr_se_scale_residual_relu(nat.ptr(v), nat.ptr(gate), nat.ptr(r), nat.ptr(out), B, H2, W2, C, nat.stream_ptr()))
    mean = v.float().mean(dim=(1, 2))
    gref = torch.sigmoid(torch.relu(mean @ w1.t()) @ w2.t())
    report("se gate", gate, gref)
    report("se apply", out, (v.float() * gref.view(B, 1, 1, C) + r.float()).relu())
except Exception as e:
    print("POINTWISE FAILED:", repr(e)); res["pw_exception"] = repr(e)

# greedy
try:
    for dt in (torch.float32, torch.bfloat16):
        T, B, C = 300, 3, 7375
        g = torch.Generator().manual_seed(3)
        lg = torch.randn(B, T, C, generator=g).to(dev).to(dt)
        am = torch.empty(B, T, dtype=torch.int32, device=dev); oi = torch.empty(B, T, dtype=torch.int32, device=dev)
        ol = torch.empty(B, dtype=torch.int32, device=dev)
        nat.check(lib.hctr_ctc_greedy_decode(nat.ptr(lg), nat.HCTR_F32 if dt == torch.float32 else nat.HCTR_BF16, T, B, C,
                                             C, T * C, nat.ptr(am), nat.ptr(oi), nat.ptr(ol), nat.stream_ptr()))
        ref = lg.float().argmax(2).int()
        print("greedy argmax equal (%s):" % dt, bool((ref == am).all().item()), "lens", ol.tolist())
        res["greedy_%s" % dt] = 0.0 if bool((ref == am).all().item()) else 1.0
except Exception as e:
    print("GREEDY FAILED:", repr(e)); res["greedy_exception"] = repr(e)

# full model smoke + timing
try:
    from hctr_b200.models.handwritten_ctr_model import hctr_model
    torch.manual_seed(1234)
    m = hctr_model(7375).to(dev).eval()
    x = (torch.rand(2, 1, 128, 256, device=dev) * 2 - 1)
    with torch.no_grad():
        y = m(x)
    torch.cuda.synchronize()
    print("model out", tuple(y.shape), y.dtype, float(y.abs().max()), bool(torch.isfinite(y).all()))
    m.logits_dtype = torch.bfloat16
    for (B, W) in [(8, 2048), (64, 2048)]:
        x = (torch.rand(B, 1, 128, W, device=dev) * 2 - 1)
        with torch.no_grad():
            for _ in range(2): y = m(x)
            torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3): y = m(x)
            e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        tf = 1358901248 * W * B / (ms * 1e-3) / 1e12
        print("forward B=%d W=%d: %.2f ms  -> %.1f lines/s  %.1f TFLOP/s" % (B, W, ms, B / (ms * 1e-3), tf), flush=True)
        res["fwd_ms_B%d" % B] = ms
except Exception as e:
    import traceback; traceback.print_exc()
    res["model_exception"] = repr(e)

os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/bringup.json", "w"), indent=1)
