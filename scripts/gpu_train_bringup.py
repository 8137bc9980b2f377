"""First-contact check of the training kernels against torch autograd on the GPU box."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, torch.nn.functional as F
import hctr_b200
from hctr_b200 import native as nat
torch.backends.cuda.matmul.allow_tf32 = False; torch.backends.cudnn.allow_tf32 = False
lib = nat.lib(); dev = "cuda"
S = nat.stream_ptr

def rep(name, got, ref):
    err = (got.float() - ref.float()).abs().max().item(); sc = ref.float().abs().max().item()
    print("%-46s err %.3e  scale %.3e  rel %.2e" % (name, err, sc, err / max(sc, 1e-30)), flush=True)

def ws(nbytes):
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=dev)

# ---------------- wgrad / dgrad
for (B, H, W, Cin, Cout, k) in [(2, 4, 128, 64, 64, 3), (2, 8, 200, 64, 128, 3), (1, 6, 96, 128, 128, 3), (2, 4, 300, 256, 512, 3),
                                 (1, 4, 256, 512, 512, 3), (2, 8, 256, 64, 128, 1), (1, 4, 130, 256, 512, 1)]:
    try:
        g = torch.Generator().manual_seed(B + W + Cin)
        x = torch.randn(B, Cin, H, W, generator=g).to(dev).to(torch.bfloat16)
        w = (torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5).to(dev).to(torch.bfloat16)
        dz = torch.randn(B, Cout, H, W, generator=g).to(dev).to(torch.bfloat16)
        xr = x.float().requires_grad_(True); wr = w.float().requires_grad_(True)
        F.conv2d(xr, wr, padding=k // 2).backward(dz.float())
        xn = x.permute(0, 2, 3, 1).contiguous(); dzn = dz.permute(0, 2, 3, 1).contiguous()
        dw = torch.full((Cout, Cin, k, k), float("nan"), device=dev)
        nb = lib.hctr_wgrad_workspace_bytes(B, H, W, Cout, Cin, k * k)
        wsp = ws(nb)
        nat.check(lib.hctr_conv_wgrad(nat.ptr(dzn), nat.ptr(xn), nat.ptr(dw), B, H, W, Cout, Cin, k, nat.ptr(wsp), nb, S()), "wgrad")
        rep("wgrad B%d H%d W%d %d->%d k%d" % (B, H, W, Cin, Cout, k), dw, wr.grad)
        wt = w.permute(1, 2, 3, 0).contiguous()          # [Cin][kh][kw][Cout]
        ones = torch.ones(Cin, device=dev); zeros = torch.zeros(Cin, device=dev)
        add = torch.randn(B, H, W, Cin, generator=g).to(dev).to(torch.bfloat16)
        dx = torch.full((B, H, W, Cin), float("nan"), dtype=torch.bfloat16, device=dev)
        nat.check(lib.hctr_conv_dgrad(nat.ptr(dzn), nat.ptr(wt), nat.ptr(ones), nat.ptr(zeros), nat.ptr(add), nat.ptr(dx),
                                      B, H, W, Cout, Cin, k, S()), "dgrad")
        rep("dgrad(+add) same", dx.permute(0, 3, 1, 2), xr.grad + add.float().permute(0, 3, 1, 2))
    except Exception as e:
        print("WGRAD/DGRAD FAILED", (B, H, W, Cin, Cout, k), repr(e))

# ---------------- pointwise train fwd/bwd
def pointwise_case(B, H, W, C, gate, res, relu, pool):
    g = torch.Generator().manual_seed(C + H)
    z = (torch.randn(B, C, H, W, generator=g) * 1.5 + 0.3).to(dev).to(torch.bfloat16)
    gamma = (torch.rand(C, generator=g) + 0.5).to(dev); beta = (0.2 * torch.randn(C, generator=g)).to(dev)
    r = torch.randn(B, C, H, W, generator=g).to(dev).to(torch.bfloat16) if res else None
    Cr = C // 16
    w1 = (torch.randn(Cr, C, generator=g) / C ** 0.5).to(dev); w2 = (torch.randn(C, Cr, generator=g) / Cr ** 0.5).to(dev)
    Ho = H // 2 if pool else H
    dout = torch.randn(B, C, Ho, W, generator=g).to(dev).to(torch.bfloat16)
    # torch reference
    zr = z.float().requires_grad_(True); gr = gamma.clone().requires_grad_(True); br = beta.clone().requires_grad_(True)
    w1r = w1.clone().requires_grad_(True); w2r = w2.clone().requires_grad_(True)
    rr = r.float().requires_grad_(True) if res else None
    y = F.batch_norm(zr, None, None, gr, br, True, 0.1, 1e-5)
    if gate:
        m = y.mean(dim=(2, 3)); gt = torch.sigmoid(torch.relu(m @ w1r.t()) @ w2r.t()); y = y * gt.view(B, C, 1, 1)
    if res: y = y + rr
    if relu: y = y.relu()
    if pool: y = F.max_pool2d(y, (2, 1), (2, 1))
    y.backward(dout.float())
    # ours
    zn = z.permute(0, 2, 3, 1).contiguous(); rn = r.permute(0, 2, 3, 1).contiguous() if res else None
    dn = dout.permute(0, 2, 3, 1).contiguous()
    slices = lib.hctr_stat_slices(B, H, W)
    ps = torch.empty(B, slices, C, device=dev); pq = torch.empty(B, slices, C, device=dev)
    nat.check(lib.hctr_chan_stats(nat.ptr(zn), nat.ptr(ps), nat.ptr(pq), B, H, W, C, S()))
    mean = torch.empty(C, device=dev); invstd = torch.empty(C, device=dev); scale = torch.empty(C, device=dev); shift = torch.empty(C, device=dev)
    line = torch.empty(B, C, device=dev); rm = torch.zeros(C, device=dev); rv = torch.ones(C, device=dev)
    nat.check(lib.hctr_bn_finalize_train(nat.ptr(ps), nat.ptr(pq), B, slices, C, H * W, nat.ptr(gamma), nat.ptr(beta), 1e-5, 0.1,
                                         nat.ptr(rm), nat.ptr(rv), nat.ptr(mean), nat.ptr(invstd), nat.ptr(scale), nat.ptr(shift), nat.ptr(line), S()))
    gt_ = hid = sem = None
    if gate:
        gt_ = torch.empty(B, C, device=dev); hid = torch.empty(B, Cr, device=dev); sem = torch.empty(B, C, device=dev)
        nat.check(lib.hctr_se_excite_train(nat.ptr(line), nat.ptr(scale), nat.ptr(shift), nat.ptr(w1), nat.ptr(w2), nat.ptr(sem), nat.ptr(hid), nat.ptr(gt_), B, C, Cr, H * W, S()))
    out = torch.empty(B, Ho, W, C, dtype=torch.bfloat16, device=dev)
    nat.check(lib.hctr_train_apply_fwd(nat.ptr(zn), nat.ptr(scale), nat.ptr(shift), nat.ptr(gt_), nat.ptr(rn), nat.ptr(out), B, H, W, C, relu, pool, 0.0, 0, S()))
    tag = "pw C%d H%d g%d r%d relu%d pool%d" % (C, H, gate, res, relu, pool)
    rep(tag + " fwd", out.permute(0, 3, 1, 2), y.detach())
    zref = z.float(); rep(tag + " running_var", rv, 0.9 + 0.1 * zref.transpose(0, 1).reshape(C, -1).var(dim=1, unbiased=True))
    a2 = torch.empty(B, slices, C, device=dev); a3 = torch.empty(B, slices, C, device=dev)
    nat.check(lib.hctr_train_bwd_reduce(nat.ptr(dn), nat.ptr(zn), nat.ptr(scale), nat.ptr(shift), nat.ptr(gt_), nat.ptr(rn), nat.ptr(a2), nat.ptr(a3), B, H, W, C, relu, pool, 0.0, 0, S()))
    dgam = torch.empty(C, device=dev); dbet = torch.empty(C, device=dev); dbias = torch.empty(C, device=dev)
    P = torch.empty(B, C, device=dev); Q = torch.empty(B, C, device=dev); R = torch.empty(C, device=dev)
    dw1 = torch.empty(Cr, C, device=dev) if gate else None; dw2 = torch.empty(C, Cr, device=dev) if gate else None
    nat.check(lib.hctr_train_bwd_finalize(nat.ptr(a2), nat.ptr(a3), slices, B, C, H * W, nat.ptr(gamma), nat.ptr(mean), nat.ptr(invstd), nat.ptr(scale), nat.ptr(shift),
                                          nat.ptr(line), nat.ptr(gt_), nat.ptr(hid), nat.ptr(sem), nat.ptr(w1) if gate else None, nat.ptr(w2) if gate else None, Cr,
                                          nat.ptr(dw1), nat.ptr(dw2), nat.ptr(dgam), nat.ptr(dbet), nat.ptr(dbias), nat.ptr(P), nat.ptr(Q), nat.ptr(R), S()))
    dzo = torch.empty(B, H, W, C, dtype=torch.bfloat16, device=dev)
    dro = torch.empty(B, H, W, C, dtype=torch.bfloat16, device=dev) if res else None
    nat.check(lib.hctr_train_bwd_apply(nat.ptr(dn), nat.ptr(zn), nat.ptr(scale), nat.ptr(shift), nat.ptr(gt_), nat.ptr(rn), nat.ptr(P), nat.ptr(Q), nat.ptr(R),
                                       nat.ptr(dzo), nat.ptr(dro), B, H, W, C, relu, pool, 0.0, 0, S()))
    rep(tag + " dz", dzo.permute(0, 3, 1, 2), zr.grad); rep(tag + " dgamma", dgam, gr.grad); rep(tag + " dbeta", dbet, br.grad)
    rep(tag + " dbias(sum dz)", dbias, zr.grad.sum(dim=(0, 2, 3)) + 0 * dbias) if True else None
    if res: rep(tag + " dres", dro.permute(0, 3, 1, 2), rr.grad)
    if gate: rep(tag + " dW1", dw1, w1r.grad); rep(tag + " dW2", dw2, w2r.grad)

for case in [(2, 8, 200, 64, 0, 0, 1, 1), (2, 8, 136, 128, 0, 0, 1, 0), (2, 8, 136, 128, 1, 1, 1, 0), (3, 4, 300, 256, 1, 1, 1, 0),
             (2, 4, 130, 512, 1, 1, 1, 0), (2, 4, 130, 512, 0, 0, 1, 1), (2, 4, 100, 256, 0, 0, 0, 0)]:
    try:
        pointwise_case(*case)
    except Exception as e:
        import traceback; traceback.print_exc(); print("POINTWISE FAILED", case)

# dropout statistics
try:
    B, H, W, C = 2, 8, 512, 128
    z = torch.ones(B, H, W, C, dtype=torch.bfloat16, device=dev); one = torch.ones(C, device=dev); zero = torch.zeros(C, device=dev)
    for p_ in (0.1, 0.3, 0.9):
        out = torch.empty_like(z)
        nat.check(lib.hctr_train_apply_fwd(nat.ptr(z), nat.ptr(one), nat.ptr(zero), None, None, nat.ptr(out), B, H, W, C, 0, 0, p_, 1234, S()))
        kept = (out != 0).float().mean().item()
        print("dropout p=%.1f keep-frac %.5f (expect %.5f) kept value %.4f (expect %.4f)" % (p_, kept, 1 - p_, out.float().max().item(), 1 / (1 - p_)))
except Exception as e:
    print("DROPOUT FAILED", repr(e))

# ---------------- classifier backward
try:
    B, W, N, Cf, Hf = 2, 200, 7375, 512, 4
    g = torch.Generator().manual_seed(5)
    feat = torch.randn(B, Hf, W, Cf, generator=g).to(dev).to(torch.bfloat16)
    wl = (torch.randn(N, Hf * Cf, generator=g) / 45).to(dev).to(torch.bfloat16)       # reference layout d = c*Hf + h
    pitch = 7376
    dl = torch.zeros(B, W, pitch, dtype=torch.bfloat16, device=dev); dl[:, :, :N] = (torch.randn(B, W, N, generator=g) * 0.1).to(dev).to(torch.bfloat16)
    fr = feat.float().requires_grad_(True); wr = wl.float().requires_grad_(True)
    a = fr.permute(0, 2, 3, 1).reshape(B * W, Cf * Hf)        # [b,w,c,h] -> d = c*Hf+h
    (a @ wr.t()).backward(dl[:, :, :N].float().reshape(B * W, N))
    # w_t: [Hf*Cf][pitch] rows k = h*Cf + c
    wk = wl.reshape(N, Cf, Hf).permute(2, 1, 0).reshape(Hf * Cf, N)
    wt = torch.zeros(Hf * Cf, pitch, dtype=torch.bfloat16, device=dev); wt[:, :N] = wk
    ones = torch.ones(Cf, device=dev); zeros = torch.zeros(Cf, device=dev)
    dfeat = torch.full((B, Hf, W, Cf), float("nan"), dtype=torch.bfloat16, device=dev)
    nat.check(lib.hctr_classifier_dgrad(nat.ptr(dl), pitch, nat.ptr(wt), nat.ptr(ones), nat.ptr(zeros), nat.ptr(dfeat), B, Hf, W, Cf, N, S()), "cls dgrad")
    rep("classifier dgrad", dfeat, fr.grad)
    nb = lib.hctr_linear_wgrad_workspace_bytes(B, Hf, W, Cf, N); wsp = ws(nb)
    dw = torch.full((N, Cf * Hf), float("nan"), device=dev)
    nat.check(lib.hctr_linear_wgrad(nat.ptr(dl), pitch, nat.ptr(feat), nat.ptr(dw), B, Hf, W, Cf, N, nat.ptr(wsp), nb, S()), "cls wgrad")
    rep("classifier wgrad", dw, wr.grad)
    nb = lib.hctr_colsum_workspace_bytes(B * W, N); wsp = ws(nb); db = torch.empty(N, device=dev)
    nat.check(lib.hctr_colsum_bf16(nat.ptr(dl), B * W, N, pitch, nat.ptr(db), nat.ptr(wsp), nb, S()))
    rep("classifier dbias", db, dl[:, :, :N].float().sum(dim=(0, 1)))
except Exception as e:
    import traceback; traceback.print_exc(); print("CLASSIFIER BWD FAILED")

# ---------------- stem wgrad + SGD
try:
    B, H, W = 2, 128, 200
    g = torch.Generator().manual_seed(6)
    x = (torch.rand(B, 1, H, W, generator=g) * 2 - 1).to(dev); dz = torch.randn(B, 64, H, W, generator=g).to(dev).to(torch.bfloat16)
    wr = torch.randn(64, 1, 3, 3, device=dev, requires_grad=True)
    F.conv2d(x, wr, padding=1).backward(dz.float())
    nb = lib.hctr_stem_wgrad_workspace_bytes(B, H, W); wsp = ws(nb); dw = torch.empty(64, 9, device=dev)
    nat.check(lib.hctr_stem_wgrad(nat.ptr(dz.permute(0, 2, 3, 1).contiguous()), nat.ptr(x), nat.ptr(dw), B, H, W, nat.ptr(wsp), nb, S()))
    rep("stem wgrad", dw.view(64, 1, 3, 3), wr.grad)
    n = 1000003
    p0 = torch.randn(n, device=dev); g0 = torch.randn(n, device=dev) * 0.02
    pr = p0.clone().requires_grad_(True); opt = torch.optim.SGD([pr], lr=1e-3, momentum=0.9, weight_decay=1e-4)
    pm = p0.clone(); buf = torch.zeros(n, device=dev); normo = torch.zeros(2, device=dev); wsp = ws(lib.hctr_sgd_workspace_bytes())
    for step in range(3):
        gs = g0 * (step + 1)
        pr.grad = gs.clone(); tn = torch.nn.utils.clip_grad_norm_([pr], 5.0); opt.step()
        nat.check(lib.hctr_sgd_clip_step(nat.ptr(pm), nat.ptr(gs), nat.ptr(buf), n, 1.0, 5.0, 1e-3, 0.9, 1e-4, int(step == 0), nat.ptr(normo), nat.ptr(wsp), S()))
        print("sgd step %d: norm ours %.5f torch %.5f  param err %.3e" % (step, normo[0].item(), tn.item(), (pm - pr.detach()).abs().max().item()))
except Exception as e:
    import traceback; traceback.print_exc(); print("STEM/SGD FAILED")
