"""Kernel-level GPU parity of the training kernels (through the C ABI) against torch autograd on the same bf16-rounded
operands. Tolerances: fp32 outputs from bf16 operands with fp32 accumulate <= 1e-4 relative; bf16 outputs <= 2^-7 of the
output scale (half an ulp + reduction-order noise)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
GATE = 2.0 ** -7


@pytest.fixture(scope="module", autouse=True)
def _no_tf32():
    a, b = torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield
    torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = a, b


def _nat():
    from hctr_b200 import native
    return native


def _ws(n):
    return torch.empty(max(int(n), 16), dtype=torch.uint8, device="cuda")


def _close(got, ref, rel):
    assert torch.isfinite(got.float()).all()
    assert (got.float() - ref.float()).abs().max().item() <= rel * ref.float().abs().max().item()


@pytest.mark.parametrize("B,H,W,Cin,Cout,k", [(2, 4, 128, 64, 64, 3), (2, 8, 200, 64, 128, 3), (1, 6, 96, 128, 128, 3),
                                               (2, 4, 300, 256, 512, 3), (1, 4, 256, 512, 512, 3), (2, 8, 256, 64, 128, 1),
                                               (1, 4, 130, 256, 512, 1), (3, 2, 1, 64, 64, 3)])
def test_conv_wgrad_and_dgrad(B, H, W, Cin, Cout, k):
    nat = _nat(); lib = nat.lib(); S = nat.stream_ptr
    g = torch.Generator().manual_seed(B + W + Cin)
    x = torch.randn(B, Cin, H, W, generator=g).cuda().to(torch.bfloat16)
    w = (torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5).cuda().to(torch.bfloat16)
    dz = torch.randn(B, Cout, H, W, generator=g).cuda().to(torch.bfloat16)
    xr, wr = x.float().requires_grad_(True), w.float().requires_grad_(True)
    F.conv2d(xr, wr, padding=k // 2).backward(dz.float())
    xn, dzn = x.permute(0, 2, 3, 1).contiguous(), dz.permute(0, 2, 3, 1).contiguous()
    dw = torch.full((Cout, Cin, k, k), float("nan"), device="cuda")
    nb = lib.hctr_wgrad_workspace_bytes(B, H, W, Cout, Cin, k * k)
    ws = _ws(nb)
    nat.check(lib.hctr_conv_wgrad(nat.ptr(dzn), nat.ptr(xn), nat.ptr(dw), B, H, W, Cout, Cin, k, nat.ptr(ws), nb, S()))
    _close(dw, wr.grad, 1e-4)
    dw2 = torch.empty_like(dw)
    nat.check(lib.hctr_conv_wgrad(nat.ptr(dzn), nat.ptr(xn), nat.ptr(dw2), B, H, W, Cout, Cin, k, nat.ptr(ws), nb, S()))
    assert torch.equal(dw, dw2)                                   # split-K reduction is fixed-order: deterministic
    wt = w.permute(1, 2, 3, 0).contiguous()
    ones, zeros = torch.ones(Cin, device="cuda"), torch.zeros(Cin, device="cuda")
    add = torch.randn(B, H, W, Cin, generator=g).cuda().to(torch.bfloat16)
    for a in (None, add):
        dx = torch.full((B, H, W, Cin), float("nan"), dtype=torch.bfloat16, device="cuda")
        nat.check(lib.hctr_conv_dgrad(nat.ptr(dzn), nat.ptr(wt), nat.ptr(ones), nat.ptr(zeros), nat.ptr(a), nat.ptr(dx),
                                      B, H, W, Cout, Cin, k, S()))
        ref = xr.grad + (a.float().permute(0, 3, 1, 2) if a is not None else 0)
        _close(dx.permute(0, 3, 1, 2), ref, GATE)


@pytest.mark.parametrize("B,H,W,C,gate,res,relu,pool", [(2, 8, 200, 64, 0, 0, 1, 1), (2, 8, 136, 128, 0, 0, 1, 0),
                                                         (2, 8, 136, 128, 1, 1, 1, 0), (3, 4, 300, 256, 1, 1, 1, 0),
                                                         (2, 4, 130, 512, 1, 1, 1, 0), (2, 4, 130, 512, 0, 0, 1, 1),
                                                         (2, 4, 100, 256, 0, 0, 0, 0), (16, 2, 64, 512, 1, 1, 1, 0),
                                                         (2, 16, 520, 256, 1, 1, 1, 0), (2, 32, 512, 64, 0, 0, 1, 1),
                                                         (5, 8, 264, 128, 0, 1, 1, 0), (33, 2, 40, 64, 1, 0, 1, 0)])
def test_bn_se_act_unit_forward_backward(B, H, W, C, gate, res, relu, pool):
    """BatchNorm(train) [+SE] [+residual] [+ReLU] [+(2,1) pool]: forward, running stats, dz, dres, dgamma, dbeta, SE grads."""
    nat = _nat(); lib = nat.lib(); S = nat.stream_ptr
    dev = "cuda"
    g = torch.Generator().manual_seed(C + H)
    z = (torch.randn(B, C, H, W, generator=g) * 1.5 + 0.3).cuda().to(torch.bfloat16)
    gamma = (torch.rand(C, generator=g) + 0.5).cuda(); beta = (0.2 * torch.randn(C, generator=g)).cuda()
    r = torch.randn(B, C, H, W, generator=g).cuda().to(torch.bfloat16) if res else None
    Cr = C // 16
    w1 = (torch.randn(Cr, C, generator=g) / C ** 0.5).cuda(); w2 = (torch.randn(C, Cr, generator=g) / Cr ** 0.5).cuda()
    Ho = H // 2 if pool else H
    dout = torch.randn(B, C, Ho, W, generator=g).cuda().to(torch.bfloat16)
    zr = z.float().requires_grad_(True); gr = gamma.clone().requires_grad_(True); br = beta.clone().requires_grad_(True)
    w1r = w1.clone().requires_grad_(True); w2r = w2.clone().requires_grad_(True)
    rr = r.float().requires_grad_(True) if res else None
    y = F.batch_norm(zr, None, None, gr, br, True, 0.1, 1e-5)
    if gate:
        mm = y.mean(dim=(2, 3)); gt = torch.sigmoid(torch.relu(mm @ w1r.t()) @ w2r.t()); y = y * gt.view(B, C, 1, 1)
    if res: y = y + rr
    if relu: y = y.relu()
    if pool: y = F.max_pool2d(y, (2, 1), (2, 1))
    y.backward(dout.float())
    zn = z.permute(0, 2, 3, 1).contiguous(); rn = r.permute(0, 2, 3, 1).contiguous() if res else None
    dn = dout.permute(0, 2, 3, 1).contiguous()
    slices = lib.hctr_stat_slices(B, H, W)
    ps = torch.empty(B, slices, C, device=dev); pq = torch.empty(B, slices, C, device=dev)
    nat.check(lib.hctr_chan_stats(nat.ptr(zn), nat.ptr(ps), nat.ptr(pq), B, H, W, C, S()))
    st = torch.empty(4, C, device=dev); line = torch.empty(B, C, device=dev)
    rm = torch.zeros(C, device=dev); rv = torch.ones(C, device=dev)
    nat.check(lib.hctr_bn_finalize_train(nat.ptr(ps), nat.ptr(pq), B, slices, C, H * W, nat.ptr(gamma), nat.ptr(beta), 1e-5, 0.1,
                                         nat.ptr(rm), nat.ptr(rv), nat.ptr(st[0]), nat.ptr(st[1]), nat.ptr(st[2]), nat.ptr(st[3]),
                                         nat.ptr(line), S()))
    gt_ = hid = sem = None
    if gate:
        gt_ = torch.empty(B, C, device=dev); hid = torch.empty(B, Cr, device=dev); sem = torch.empty(B, C, device=dev)
        nat.check(lib.hctr_se_excite_train(nat.ptr(line), nat.ptr(st[2]), nat.ptr(st[3]), nat.ptr(w1), nat.ptr(w2), nat.ptr(sem),
                                           nat.ptr(hid), nat.ptr(gt_), B, C, Cr, H * W, S()))
    out = torch.empty(B, Ho, W, C, dtype=torch.bfloat16, device=dev)
    mask = torch.empty(B, H, W, C // 8, dtype=torch.uint8, device=dev)
    nat.check(lib.hctr_train_apply_fwd(nat.ptr(zn), nat.ptr(st[2]), nat.ptr(st[3]), nat.ptr(gt_), nat.ptr(rn), nat.ptr(out),
                                       nat.ptr(mask), B, H, W, C, relu, pool, 0.0, 0, S()))
    _close(out.permute(0, 3, 1, 2), y.detach(), GATE)
    zf = z.float().transpose(0, 1).reshape(C, -1)
    _close(rm, 0.1 * zf.mean(dim=1), 1e-4)
    _close(rv, 0.9 + 0.1 * zf.var(dim=1, unbiased=True), 1e-4)
    a2 = torch.empty(B, slices, C, device=dev); a3 = torch.empty(B, slices, C, device=dev)
    nat.check(lib.hctr_train_bwd_reduce(nat.ptr(dn), nat.ptr(zn), nat.ptr(mask), nat.ptr(a2), nat.ptr(a3), B, H, W, C, pool, 0.0, S()))
    dgam = torch.empty(C, device=dev); dbet = torch.empty(C, device=dev); dbias = torch.empty(C, device=dev)
    PQ = torch.empty(2, B, C, device=dev); R = torch.empty(C, device=dev)
    dw1 = torch.empty(Cr, C, device=dev) if gate else None; dw2 = torch.empty(C, Cr, device=dev) if gate else None
    nat.check(lib.hctr_train_bwd_finalize(nat.ptr(a2), nat.ptr(a3), slices, B, C, H * W, nat.ptr(gamma), nat.ptr(st[0]), nat.ptr(st[1]),
                                          nat.ptr(st[2]), nat.ptr(st[3]), nat.ptr(line), nat.ptr(gt_), nat.ptr(hid), nat.ptr(sem),
                                          nat.ptr(w1) if gate else None, nat.ptr(w2) if gate else None, Cr, nat.ptr(dw1), nat.ptr(dw2),
                                          nat.ptr(dgam), nat.ptr(dbet), nat.ptr(dbias), nat.ptr(PQ[0]), nat.ptr(PQ[1]), nat.ptr(R), S()))
    dzo = torch.empty(B, H, W, C, dtype=torch.bfloat16, device=dev)
    dro = torch.empty(B, H, W, C, dtype=torch.bfloat16, device=dev) if res else None
    nat.check(lib.hctr_train_bwd_apply(nat.ptr(dn), nat.ptr(zn), nat.ptr(mask), nat.ptr(PQ[0]), nat.ptr(PQ[1]), nat.ptr(R),
                                       nat.ptr(dzo), nat.ptr(dro), B, H, W, C, pool, 0.0, S()))
    _close(dzo.permute(0, 3, 1, 2), zr.grad, GATE)
    _close(dgam, gr.grad, 1e-4); _close(dbet, br.grad, 1e-4)
    assert dbias.abs().max().item() <= 1e-3 * max(1.0, dbet.abs().max().item())     # sum dz == 0 behind a train-mode BN
    if res:
        _close(dro.permute(0, 3, 1, 2), rr.grad, GATE)
    if gate:
        _close(dw1, w1r.grad, 1e-4); _close(dw2, w2r.grad, 1e-4)


def test_dropout_mask_statistics_and_backward_consistency():
    nat = _nat(); lib = nat.lib(); S = nat.stream_ptr
    B, H, W, C = 2, 8, 512, 128
    z = torch.ones(B, H, W, C, dtype=torch.bfloat16, device="cuda")
    one, zero = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
    for p in (0.1, 0.3, 0.9):
        out = torch.empty_like(z)
        mask = torch.empty(B, H, W, C // 8, dtype=torch.uint8, device="cuda")
        nat.check(lib.hctr_train_apply_fwd(nat.ptr(z), nat.ptr(one), nat.ptr(zero), None, None, nat.ptr(out), nat.ptr(mask), B, H, W, C,
                                           0, 0, p, 1234, S()))
        kept = (out != 0)
        n = kept.numel()
        assert abs(kept.float().mean().item() - (1 - p)) <= 5 * (p * (1 - p) / n) ** 0.5 + 1e-4      # 5 sigma
        assert abs(out.float().max().item() - 1 / (1 - p)) <= 2.0 ** -7 / (1 - p)
        # per-channel keep rates are uniform too
        assert (kept.float().mean(dim=(0, 1, 2)) - (1 - p)).abs().max().item() <= 0.03
        # the stored keep-mask drives the backward: with P=1, Q=R=0 -> dz = d_pre = dout * mask / (1-p)
        dout = torch.ones_like(z)
        P = torch.ones(B, C, device="cuda"); Q = torch.zeros(B, C, device="cuda"); R = torch.zeros(C, device="cuda")
        dz = torch.empty_like(z)
        nat.check(lib.hctr_train_bwd_apply(nat.ptr(dout), nat.ptr(z), nat.ptr(mask), nat.ptr(P), nat.ptr(Q), nat.ptr(R), nat.ptr(dz),
                                           None, B, H, W, C, 0, p, S()))
        assert torch.equal(dz != 0, kept)
        assert abs(dz.float().max().item() - 1 / (1 - p)) <= 2.0 ** -7 / (1 - p)
        out2 = torch.empty_like(z)
        nat.check(lib.hctr_train_apply_fwd(nat.ptr(z), nat.ptr(one), nat.ptr(zero), None, None, nat.ptr(out2), None, B, H, W, C, 0, 0, p, 99, S()))
        assert not torch.equal(out2 != 0, kept)


def test_classifier_backward_kernels():
    nat = _nat(); lib = nat.lib(); S = nat.stream_ptr
    B, W, N, Cf, Hf = 2, 200, 7375, 512, 4
    g = torch.Generator().manual_seed(5)
    feat = torch.randn(B, Hf, W, Cf, generator=g).cuda().to(torch.bfloat16)
    wl = (torch.randn(N, Hf * Cf, generator=g) / 45).cuda().to(torch.bfloat16)          # reference layout d = c*Hf + h
    pitch = 7376
    dl = torch.full((B, W, pitch), float("nan"), dtype=torch.bfloat16, device="cuda")    # the pad column must never be read
    dl[:, :, :N] = (torch.randn(B, W, N, generator=g) * 0.1).cuda().to(torch.bfloat16)
    fr, wr = feat.float().requires_grad_(True), wl.float().requires_grad_(True)
    a = fr.permute(0, 2, 3, 1).reshape(B * W, Cf * Hf)
    (a @ wr.t()).backward(dl[:, :, :N].float().reshape(B * W, N))
    wt = torch.zeros(Hf * Cf, pitch, dtype=torch.bfloat16, device="cuda")
    wt[:, :N] = wl.reshape(N, Cf, Hf).permute(2, 1, 0).reshape(Hf * Cf, N)
    ones, zeros = torch.ones(Cf, device="cuda"), torch.zeros(Cf, device="cuda")
    dfeat = torch.full((B, Hf, W, Cf), float("nan"), dtype=torch.bfloat16, device="cuda")
    nat.check(lib.hctr_classifier_dgrad(nat.ptr(dl), pitch, nat.ptr(wt), nat.ptr(ones), nat.ptr(zeros), nat.ptr(dfeat), B, Hf, W, Cf, N, S()))
    _close(dfeat, fr.grad, GATE)
    nb = lib.hctr_linear_wgrad_workspace_bytes(B, Hf, W, Cf, N); ws = _ws(nb)
    dw = torch.full((N, Cf * Hf), float("nan"), device="cuda")
    nat.check(lib.hctr_linear_wgrad(nat.ptr(dl), pitch, nat.ptr(feat), nat.ptr(dw), B, Hf, W, Cf, N, nat.ptr(ws), nb, S()))
    _close(dw, wr.grad, 1e-4)
    nb = lib.hctr_colsum_workspace_bytes(B * W, N); ws = _ws(nb); db = torch.empty(N, device="cuda")
    nat.check(lib.hctr_colsum_bf16(nat.ptr(dl), B * W, N, pitch, nat.ptr(db), nat.ptr(ws), nb, S()))
    _close(db, dl[:, :, :N].float().sum(dim=(0, 1)), 1e-5)


def test_stem_wgrad_and_fused_sgd():
    nat = _nat(); lib = nat.lib(); S = nat.stream_ptr
    B, H, W = 2, 128, 200
    g = torch.Generator().manual_seed(6)
    x = (torch.rand(B, 1, H, W, generator=g) * 2 - 1).cuda(); dz = torch.randn(B, 64, H, W, generator=g).cuda().to(torch.bfloat16)
    wr = torch.randn(64, 1, 3, 3, device="cuda", requires_grad=True)
    F.conv2d(x, wr, padding=1).backward(dz.float())
    nb = lib.hctr_stem_wgrad_workspace_bytes(B, H, W); ws = _ws(nb); dw = torch.empty(64, 9, device="cuda")
    nat.check(lib.hctr_stem_wgrad(nat.ptr(dz.permute(0, 2, 3, 1).contiguous()), nat.ptr(x), nat.ptr(dw), B, H, W, nat.ptr(ws), nb, S()))
    _close(dw.view(64, 1, 3, 3), wr.grad, 1e-4)
    # clip_grad_norm_(5.0) + SGD(momentum 0.9, wd 1e-4) (main.py:210-213,430-438) vs torch
    n = 1000003
    p0 = torch.randn(n, device="cuda"); g0 = torch.randn(n, device="cuda") * 0.02
    pr = p0.clone().requires_grad_(True); opt = torch.optim.SGD([pr], lr=1e-3, momentum=0.9, weight_decay=1e-4)
    pm = p0.clone(); buf = torch.zeros(n, device="cuda"); normo = torch.zeros(4, device="cuda"); ws = _ws(lib.hctr_sgd_workspace_bytes())
    for step in range(4):
        gs = g0 * (step + 1)
        pr.grad = gs.clone(); tn = torch.nn.utils.clip_grad_norm_([pr], 5.0); opt.step()
        nat.check(lib.hctr_sgd_clip_step(nat.ptr(pm), nat.ptr(gs), nat.ptr(buf), n, 1.0, 5.0, 1e-3, 0.9, 1e-4, int(step == 0),
                                         nat.ptr(normo), nat.ptr(ws), S()))
        assert abs(normo[0].item() - tn.item()) <= 1e-5 * tn.item()
        assert (pm - pr.detach()).abs().max().item() <= 1e-6
        assert normo[2].item() == 0.0
    # a non-finite gradient skips the update (main.py:413, GradScaler.step): parameters and momentum stay bit-identical
    for bad in (float("nan"), float("inf")):
        gs = g0.clone(); gs[12345] = bad
        p_before, b_before = pm.clone(), buf.clone()
        nat.check(lib.hctr_sgd_clip_step(nat.ptr(pm), nat.ptr(gs), nat.ptr(buf), n, 1.0, 5.0, 1e-3, 0.9, 1e-4, 0,
                                         nat.ptr(normo), nat.ptr(ws), S()))
        assert normo[2].item() == 1.0 and normo[1].item() == 0.0
        assert torch.equal(pm, p_before) and torch.equal(buf, b_before)
    # ... and a skipped FIRST step leaves a zero momentum buffer
    buf2 = torch.full((n,), 7.0, device="cuda"); gs = g0.clone(); gs[0] = float("nan"); p_before = pm.clone()
    nat.check(lib.hctr_sgd_clip_step(nat.ptr(pm), nat.ptr(gs), nat.ptr(buf2), n, 1.0, 5.0, 1e-3, 0.9, 1e-4, 1,
                                     nat.ptr(normo), nat.ptr(ws), S()))
    assert torch.equal(pm, p_before) and buf2.abs().max().item() == 0.0


def test_pack_weights_matches_torch_layouts():
    """hctr_pack_weights: every operand layout of a step in one launch vs the torch permutes it replaces."""
    import ctypes
    nat = _nat(); lib = nat.lib()
    g = torch.Generator().manual_seed(9)
    shapes = [(64, 64, 3), (128, 64, 1), (256, 128, 3), (96, 160, 3)]            # (Cout, Cin, k); the last one is ragged (tiles of 32)
    ws = [torch.randn(co, ci, k, k, generator=g).cuda() for co, ci, k in shapes]
    n, cf, hf = 7375, 512, 4
    lw = torch.randn(n, cf * hf, generator=g).cuda()
    pitch = 7376
    items = [(w, w.shape[0], w.shape[1], w.shape[2] * w.shape[3], 0, 0) for w in ws] + [(lw, n, cf, hf, 1, pitch)]
    descs = (nat.PackDesc * len(items))()
    outs, tiles = [], 0
    for i, (w, co, ci, t, mode, p) in enumerate(items):
        f = torch.full((co * ci * t,), float("nan"), dtype=torch.bfloat16, device="cuda")
        b = torch.zeros((t * ci * p if mode else co * ci * t,), dtype=torch.bfloat16, device="cuda")
        outs.append((f, b))
        descs[i].src, descs[i].dst_fwd, descs[i].dst_bwd = w.data_ptr(), f.data_ptr(), b.data_ptr()
        descs[i].cout, descs[i].cin, descs[i].taps, descs[i].bwd_mode, descs[i].bwd_pitch, descs[i].tile_start = co, ci, t, mode, p, tiles
        tiles += ((co + 31) // 32) * ((ci + 31) // 32)
    raw = torch.frombuffer(bytearray(bytes(descs)), dtype=torch.uint8).cuda()
    nat.check(lib.hctr_pack_weights(nat.ptr(raw), len(items), tiles, nat.stream_ptr()))
    torch.cuda.synchronize()
    for (w, co, ci, t, mode, p), (f, b) in zip(items, outs):
        if mode == 0:
            k = int(round(t ** 0.5))
            assert torch.equal(f.view(co, k, k, ci), w.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16))
            assert torch.equal(b.view(ci, k, k, co), w.permute(1, 2, 3, 0).contiguous().to(torch.bfloat16))
        else:
            wk = w.reshape(co, ci, t).permute(0, 2, 1).contiguous().to(torch.bfloat16).reshape(co, t * ci)      # k = h*Cf + c
            assert torch.equal(f.view(co, t * ci), wk)
            assert torch.equal(b.view(t * ci, p)[:, :co], wk.t()) and b.view(t * ci, p)[:, co:].abs().max().item() == 0.0


@pytest.mark.parametrize("B,H,W,Cin,Cout,k", [(2, 4, 300, 256, 512, 3), (3, 8, 136, 256, 256, 3), (2, 6, 200, 128, 128, 3),
                                               (2, 8, 130, 64, 64, 3), (2, 4, 200, 128, 256, 1), (2, 5, 96, 64, 128, 3)])
def test_conv_with_batch_statistics_in_the_epilogue(B, H, W, Cin, Cout, k):
    """hctr_conv_stats_fwd: z equals hctr_conv_bn_act_fwd's bit for bit, and its per-slot sums of z and z*z, finished by
    hctr_bn_finalize_train, give the statistics of the STORED (bf16) z - CTA-pair and single-CTA kernels, ragged widths, odd H."""
    nat = _nat(); lib = nat.lib(); S = nat.stream_ptr
    g = torch.Generator().manual_seed(B * 7 + W + Cout)
    x = torch.randn(B, Cin, H, W, generator=g).cuda().to(torch.bfloat16)
    w = (torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5).cuda().to(torch.bfloat16)
    bias = (0.5 * torch.randn(Cout, generator=g)).cuda()
    ones = torch.ones(Cout, device="cuda")
    xn = x.permute(0, 2, 3, 1).contiguous(); wp = w.permute(0, 2, 3, 1).contiguous()
    z0 = torch.empty(B, H, W, Cout, dtype=torch.bfloat16, device="cuda")
    nat.check(lib.hctr_conv_bn_act_fwd(nat.ptr(xn), nat.ptr(wp), nat.ptr(ones), nat.ptr(bias), nat.ptr(z0), B, H, W, Cin, Cout, k, 0, 0, S()))
    slices = lib.hctr_conv_sum_slices(H, W, Cin, Cout, k)
    ps = torch.full((B, slices, Cout), float("nan"), device="cuda"); pq = torch.full((B, slices, Cout), float("nan"), device="cuda")
    z1 = torch.empty_like(z0)
    nat.check(lib.hctr_conv_stats_fwd(nat.ptr(xn), nat.ptr(wp), nat.ptr(ones), nat.ptr(bias), nat.ptr(z1), nat.ptr(ps), nat.ptr(pq),
                                      B, H, W, Cin, Cout, k, S()))
    assert torch.equal(z0, z1)
    zf = z1.double()
    assert torch.isfinite(ps).all() and torch.isfinite(pq).all()
    s_ref, q_ref = zf.sum(dim=(1, 2)), (zf * zf).sum(dim=(1, 2))
    assert (ps.double().sum(dim=1) - s_ref).abs().max().item() <= 1e-5 * zf.abs().sum(dim=(1, 2)).max().item()
    assert (pq.double().sum(dim=1) - q_ref).abs().max().item() <= 1e-5 * q_ref.max().item()
    gamma = (torch.rand(Cout, generator=g) + 0.5).cuda(); beta = (0.2 * torch.randn(Cout, generator=g)).cuda()
    st = torch.empty(4, Cout, device="cuda"); line = torch.empty(B, Cout, device="cuda")
    rm = torch.zeros(Cout, device="cuda"); rv = torch.ones(Cout, device="cuda")
    nat.check(lib.hctr_bn_finalize_train(nat.ptr(ps), nat.ptr(pq), B, slices, Cout, H * W, nat.ptr(gamma), nat.ptr(beta), 1e-5, 0.1,
                                         nat.ptr(rm), nat.ptr(rv), nat.ptr(st[0]), nat.ptr(st[1]), nat.ptr(st[2]), nat.ptr(st[3]),
                                         nat.ptr(line), S()))
    zc = zf.reshape(-1, Cout)
    mean, var = zc.mean(dim=0), zc.var(dim=0, unbiased=False)
    _close(st[0], mean.float(), 1e-5)
    _close(st[1], (1.0 / torch.sqrt(var + 1e-5)).float(), 1e-4)
    _close(line, s_ref.float(), 1e-5)
    _close(rv, (0.9 + 0.1 * zc.var(dim=0, unbiased=True)).float(), 1e-4)
