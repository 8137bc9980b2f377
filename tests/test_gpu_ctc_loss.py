"""GPU parity of the fused log-softmax + CTC loss fwd/bwd kernels vs the fp64 oracle (pinned to torch's CTCLoss via
tests/golden/ctc_loss.npz). Tolerance: <=1e-4 relative on the loss (north star), 1e-5 abs on fp32 gradients."""
import numpy as np
import pytest
import torch

import oracle
import synth

pytestmark = pytest.mark.gpu


def _run(x, tg, tl, il, dtype=torch.float32):
    from hctr_b200.ctc_loss import CTCLoss
    crit = CTCLoss(zero_infinity=True).cuda()
    xt = torch.from_numpy(x).cuda().to(dtype).requires_grad_(True)
    loss = crit(xt, torch.from_numpy(tg), torch.IntTensor(il), torch.from_numpy(tl))
    loss.backward()
    return float(loss.item()), xt.grad.float().cpu().numpy()


@pytest.mark.parametrize("name", ["small", "mid", "wide", "len1", "infeasible"])
def test_ctc_loss_golden(golden, name):
    g = golden("ctc_loss")
    T, B, C, seed, Lmin, Lmax = [int(v) for v in g[name + "_shape"]]
    x = synth.ctc_like_logits(T, B, C, seed, peak=4.0)
    tg, tl = synth.ctc_targets(B, C, Lmin, Lmax, seed + 100, repeat_frac=0.3)
    loss, grad = _run(x, tg, tl, [T] * B)
    ref = float(g[name + "_loss"])
    assert abs(loss - ref) <= 1e-4 * max(1.0, abs(ref))
    oloss, _, ograd = oracle.ctc_loss(x, tg, [T] * B, tl)
    assert abs(loss - oloss) <= 1e-4 * max(1.0, abs(oloss))
    assert np.abs(grad - ograd).max() <= 1e-5
    if name == "infeasible":
        assert loss == 0.0 and np.abs(grad).max() == 0.0


def test_ctc_loss_same_through_log_softmax_call_form():
    """main.py:406 passes preds.log_softmax(2); the gradient reaching the logits must be the same."""
    from hctr_b200.ctc_loss import CTCLoss
    T, B, C = 50, 3, 300
    x = synth.ctc_like_logits(T, B, C, 5, peak=4.0)
    tg, tl = synth.ctc_targets(B, C, 3, 12, 6)
    crit = CTCLoss(zero_infinity=True)
    a = torch.from_numpy(x).cuda().requires_grad_(True)
    la = crit(a.log_softmax(2), torch.from_numpy(tg), torch.IntTensor([T] * B), torch.from_numpy(tl))
    la.backward()
    b = torch.from_numpy(x).cuda().requires_grad_(True)
    lb = CTCLoss.from_logits(b, torch.from_numpy(tg), torch.IntTensor([T] * B), torch.from_numpy(tl))
    lb.backward()
    assert abs(la.item() - lb.item()) <= 1e-5 * abs(lb.item())
    assert (a.grad - b.grad).abs().max().item() <= 2e-6


def test_ctc_loss_variable_input_lengths_and_model_layout():
    T, B, C = 96, 4, 7375
    x = synth.ctc_like_logits(T, B, C, 8, peak=4.0)
    tg, tl = synth.ctc_targets(B, C, 4, 20, 9, repeat_frac=0.2)
    il = [96, 70, 96, 41]
    oloss, _, ograd = oracle.ctc_loss(x, tg, il, tl)
    from hctr_b200.ctc_loss import CTCLoss
    pitch = 7376
    buf = torch.zeros((B, T, pitch), device="cuda")
    buf[:, :, :C] = torch.from_numpy(x).cuda().permute(1, 0, 2)
    view = buf[:, :, :C].permute(1, 0, 2).detach().requires_grad_(True)      # the model's [W,B,C] view
    loss = CTCLoss.from_logits(view, torch.from_numpy(tg), torch.IntTensor(il), torch.from_numpy(tl))
    loss.backward()
    assert abs(loss.item() - oloss) <= 1e-4 * abs(oloss)
    assert np.abs(view.grad.cpu().numpy() - ograd).max() <= 1e-5
    assert view.grad[70:, 1].abs().max().item() == 0.0                        # frames past input_length


def test_ctc_loss_bf16_logits():
    T, B, C = 128, 2, 7375
    x = synth.ctc_like_logits(T, B, C, 10, peak=4.0)
    xb = torch.from_numpy(x).to(torch.bfloat16)
    tg, tl = synth.ctc_targets(B, C, 10, 30, 11)
    oloss, _, ograd = oracle.ctc_loss(xb.float().numpy(), tg, [T] * B, tl)
    loss, grad = _run(xb.float().numpy(), tg, tl, [T] * B, dtype=torch.bfloat16)
    assert abs(loss - oloss) <= 1e-4 * abs(oloss)
    assert np.abs(grad - ograd).max() <= 2.0 ** -8 * np.abs(ograd).max() + 1e-6    # bf16 rounding (half an ulp) of the stored gradient


def test_ctc_gradient_rows_sum_to_zero_full_width():
    """Config-4-sized property (T=2048, C=7375): each frame's gradient sums to ~0 (softmax minus a distribution)."""
    T, B, C = 2048, 2, 7375
    x = synth.ctc_like_logits(T, B, C, 12, peak=3.0, period=40)
    tg, tl = synth.ctc_targets(B, C, 20, 60, 13)
    loss, grad = _run(x, tg, tl, [T] * B)
    assert np.isfinite(loss) and loss > 0
    assert np.abs(grad.sum(2)).max() <= 1e-6


# ---- the two recursions behind pass B: scaled linear space (one warp per sequence) and the log-space fallback ----------

def _run_flags(x, tg, tl, il, need_grad=True):
    """-> (loss, grad or None, int32 [B] flags: 1 = sequence recomputed by the log-space recursion)"""
    from hctr_b200.ctc_loss import CTCLoss, _CtcFromLogits
    _CtcFromLogits.record_fallback = True
    try:
        xt = torch.from_numpy(x).cuda().requires_grad_(need_grad)
        loss = CTCLoss.from_logits(xt, torch.from_numpy(tg), torch.IntTensor(il), torch.from_numpy(tl))
        if need_grad:
            loss.backward()
        flags = _CtcFromLogits.last_fallback.numpy().copy()
    finally:
        _CtcFromLogits.record_fallback = False
    return float(loss.item()), (xt.grad.cpu().numpy() if need_grad else None), flags


def _aligned_logits(T, B, C, tg, tl, peak, seed, noise=1.0):
    """Logits that follow the labels: every label gets a run of frames, blanks in between."""
    rs = np.random.RandomState(seed)
    x = (noise * rs.randn(T, B, C)).astype(np.float32)
    off = 0
    for b in range(B):
        L = int(tl[b])
        span = T // max(L, 1)
        for t in range(T):
            k = min(t // span, L - 1) if L else 0
            c = int(tg[off + k]) if (L and (t % span) < max(1, span // 2)) else 0
            x[t, b, c] += peak
        off += L
    return x


@pytest.mark.parametrize("L,K", [(100, 8), (200, 16), (400, 0)])
def test_ctc_loss_long_targets_use_wider_lanes(L, K):
    """S = 2L+1 up to 513 states: K = 8/16 states per lane in the scan kernel; beyond that the log-space recursion."""
    T, B, C = 900, 3, 40
    rs = np.random.RandomState(L)
    x = (1.5 * rs.randn(T, B, C)).astype(np.float32)
    tg, tl = synth.ctc_targets(B, C, L - 10, L, L + 1, repeat_frac=0.2)
    il = [T, T - 57, T]
    oloss, _, ograd = oracle.ctc_loss(x, tg, il, tl)
    loss, grad, flags = _run_flags(x, tg, tl, il)
    assert flags.tolist() == ([0, 0, 0] if K else [1, 1, 1])
    assert abs(loss - oloss) <= 1e-4 * abs(oloss)
    assert np.abs(grad - ograd).max() <= 1e-5


def test_ctc_loss_peaky_aligned_sequences_stay_on_the_linear_path():
    """A confident model whose frames follow the labels (logit gap 14, label runs of ~9 frames): states left behind decay
    by ~2^-28 per frame, stay inside the 2^-760 range a lane can hold, or are proven negligible by the verify pass."""
    T, B, C = 1024, 4, 500
    tg, tl = synth.ctc_targets(B, C, 40, 60, 21, repeat_frac=0.15)
    x = _aligned_logits(T, B, C, tg, tl, peak=14.0, seed=22)
    oloss, _, ograd = oracle.ctc_loss(x, tg, [T] * B, tl)
    loss, grad, flags = _run_flags(x, tg, tl, [T] * B)
    assert flags.tolist() == [0] * B
    assert abs(loss - oloss) <= 1e-4 * max(1.0, abs(oloss))
    assert np.abs(grad - ograd).max() <= 1e-5


def test_ctc_loss_extremely_confident_frames_any_path():
    """Logit gap 45: neighbouring states drift > 2^760 apart inside one lane while both still matter (alpha huge where beta
    is tiny); whichever recursion each sequence ends up on, the result must match the oracle."""
    T, B, C = 1024, 4, 500
    tg, tl = synth.ctc_targets(B, C, 40, 60, 21, repeat_frac=0.15)
    x = _aligned_logits(T, B, C, tg, tl, peak=45.0, seed=22)
    oloss, _, ograd = oracle.ctc_loss(x, tg, [T] * B, tl)
    loss, grad, flags = _run_flags(x, tg, tl, [T] * B)
    assert flags.sum() >= 1                                   # the verify pass must have caught at least one of them
    assert abs(loss - oloss) <= 1e-4 * max(1.0, abs(oloss))
    assert np.abs(grad - ograd).max() <= 1e-5


def test_ctc_loss_falls_back_to_log_space_when_linear_space_cannot_hold_it():
    """(1) label probabilities below e^-80, (2) confident frames that contradict the labels so that early- and
    late-emitting paths are > 2^700 apart mid-sequence although they weigh the same in the end."""
    T, B, C = 400, 3, 60
    rs = np.random.RandomState(3)
    x = (0.5 * rs.randn(T, B, C)).astype(np.float32)
    tg, tl = synth.ctc_targets(B, C, 30, 40, 31, repeat_frac=0.0)
    x[:, 0, 0] += 120.0                                        # sequence 0: every label at lp ~ -120 -> case (1)
    x[:, 1, 0] += 60.0                                         # sequence 1: each emission costs e^-60 whenever it happens -> case (2)
    oloss, _, ograd = oracle.ctc_loss(x, tg, [T] * B, tl)
    loss, grad, flags = _run_flags(x, tg, tl, [T] * B)
    assert flags.tolist() == [1, 1, 0]
    assert abs(loss - oloss) <= 1e-4 * abs(oloss)
    assert np.abs(grad - ograd).max() <= 1e-5


def test_ctc_loss_only_no_gradient():
    T, B, C = 300, 3, 80
    x = synth.ctc_like_logits(T, B, C, 41, peak=4.0)
    tg, tl = synth.ctc_targets(B, C, 10, 30, 42)
    tl[2] = 0                                                   # an empty target: nll = -sum log p(blank)
    tg = tg[:int(tl[:2].sum())]
    oloss, _, _ = oracle.ctc_loss(x, tg, [T] * B, tl, need_grad=False)
    loss, _, flags = _run_flags(x, tg, tl, [T] * B, need_grad=False)
    assert abs(loss - oloss) <= 1e-4 * abs(oloss)
    loss2, grad, _ = _run_flags(x, tg, tl, [T] * B)
    assert abs(loss2 - loss) <= 1e-6 * abs(loss)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("T,B,C,lo,hi", [(2048, 4, 7375, 20, 60), (257, 3, 101, 1, 9), (64, 2, 37, 0, 3), (1, 2, 11, 0, 1)])
def test_ctc_loss_overlapped_and_sequential_paths_agree(monkeypatch, dtype, T, B, C, lo, hi):
    """Round 2: the split schedule reads every logits row once for the log-sum-exp and the label gather, then runs
    the alpha/beta scans with the dense part of the gradient on a helper stream underneath them, then a sparse fix-up;
    HCTR_CTC_OVERLAP=2 is the one-pass rows kernel (lse + gather + dense gradient, the row in the registers of four warps)
    back to back with the scans, =1 overlaps those two through progress counters, =0 runs round 1's three sequential passes.
    Same loss, same gradient for all of them, for contiguous and model-layout (pitched, permuted) logits, with ragged input
    lengths, odd T, and T = 1."""
    from hctr_b200.ctc_loss import CTCLoss
    x = synth.ctc_like_logits(T, B, C, 7 + T)
    tg, tl = synth.ctc_targets(B, C, lo, hi, 9 + T)
    il = [T] * B
    if T > 8:
        il[-1] = T - 5
    results = []
    # default, split (normalised tables), one-pass rows back to back, overlapped, round-1 passes, split in its relative form
    # (label gather relative to the row's largest label logit, log-sum-exp pass beside the scans; the default for large calls)
    for mode in ("", "4", "2", "1", "0", "5"):
        monkeypatch.setenv("HCTR_CTC_OVERLAP", mode)
        for layout in ("contiguous", "model"):
            if layout == "contiguous":
                xt = torch.from_numpy(x).cuda().to(dtype).requires_grad_(True)
                view = xt
            else:
                pitch = (C + 7) // 8 * 8
                buf = torch.zeros(B, T, pitch, device="cuda", dtype=dtype)
                buf[:, :, :C] = torch.from_numpy(x).cuda().to(dtype).permute(1, 0, 2)
                xt = buf.requires_grad_(True)
                view = xt[:, :, :C].permute(1, 0, 2)
            loss = CTCLoss.from_logits(view, torch.from_numpy(tg), torch.IntTensor(il), torch.from_numpy(tl))
            loss.backward()
            g = xt.grad.float()
            if layout == "model":
                assert g[:, :, C:].abs().max().item() == 0.0 if pitch > C else True
                g = g[:, :, :C].permute(1, 0, 2)
            results.append((loss.item(), g.contiguous().cpu()))
    xo = torch.from_numpy(x).to(dtype).float().numpy()
    oloss, _, ograd = oracle.ctc_loss(xo, tg, il, tl)
    gate = 1e-5 if dtype == torch.float32 else 2.0 ** -8 * float(np.abs(ograd).max()) + 1e-7
    for loss, g in results:
        assert abs(loss - oloss) <= 1e-4 * abs(oloss) + 1e-6, (loss, oloss)
        assert np.abs(g.numpy() - ograd).max() <= gate
    assert abs(results[0][0] - results[2][0]) <= 1e-6 * abs(oloss) + 1e-7          # default vs forced split schedule
    assert abs(results[0][0] - results[4][0]) <= 1e-5 * abs(oloss) + 1e-6          # default vs one-pass rows kernel
    assert abs(results[4][0] - results[6][0]) <= 1e-6 * abs(oloss) + 1e-7          # rows kernel: back to back vs overlapped
    assert abs(results[0][0] - results[8][0]) <= 1e-5 * abs(oloss) + 1e-6          # default vs round-1 passes
    assert abs(results[2][0] - results[10][0]) <= 1e-6 * abs(oloss) + 1e-7         # split: normalised vs relative label tables
    assert np.abs(results[2][1].numpy() - results[10][1].numpy()).max() <= (2e-6 if dtype == torch.float32 else gate)


def test_ctc_loss_bad_lengths_and_labels_do_not_touch_memory_out_of_bounds():
    """Device-side lengths / labels are not trusted: out-of-range values are clamped for addressing and the loss is NaN."""
    from hctr_b200 import native as nat
    T, B, C, max_l = 32, 2, 17, 4
    lib = nat.lib()
    x = torch.randn(T, B, C, device="cuda")
    tg = torch.tensor([1, 2, 3, 99, 5, 6, 7, 8], dtype=torch.int32, device="cuda")          # label 99 >= C
    tl = torch.tensor([4, 4], dtype=torch.int32, device="cuda")
    il = torch.full((B,), T, dtype=torch.int32, device="cuda")
    nb = lib.hctr_ctc_loss_workspace_bytes(T, B, max_l)
    ws = torch.empty(nb + 256, dtype=torch.uint8, device="cuda"); off = (-ws.data_ptr()) % 256
    nll = torch.empty(B, device="cuda"); loss = torch.empty(1, device="cuda"); grad = torch.empty_like(x)

    def run(tlen):
        nat.check(lib.hctr_ctc_loss_fwd_bwd(nat.ptr(x), nat.HCTR_F32, T, B, C, x.stride(0), x.stride(1), nat.ptr(tg), nat.ptr(tlen),
                                            nat.ptr(il), max_l, None, nat.ptr(nll), nat.ptr(loss), nat.ptr(grad), 1.0,
                                            nat.c_void_p(ws.data_ptr() + off), nb, nat.stream_ptr()))
        torch.cuda.synchronize()
        return loss.item()
    assert run(tl) != run(tl)                                                  # NaN: a label is out of range
    tg[3] = 4
    assert run(tl) == run(tl)                                                  # finite and reproducible once repaired
    bad = run(torch.tensor([4, 4000], dtype=torch.int32, device="cuda"))       # length beyond max_target_len
    assert bad != bad
