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
