"""CTC loss for the training step, fused with log-softmax, on sm_100a kernels.

Reference call site (main.py:205,406-409):
    criterion = CTCLoss(zero_infinity=True)
    loss = criterion(preds.log_softmax(2), targets, input_lengths, target_lengths)

`CTCLoss` here keeps that call form (log-probs in, 0-d loss out, autograd) - since the gradient of
log_softmax applied to the CTC gradient is the identity on it (its class-sum is zero), feeding
log-probs or raw logits gives the same result. `CTCLoss.from_logits(...)` is the fused variant the
B200 training step uses: logits are read twice and the gradient written once.
Semantics: blank=0, reduction='mean' (per-sequence nll / max(L,1), then batch mean), zero_infinity.
"""
import numpy as np
import torch
import torch.nn as nn

from . import native as nat


def _as_i32(x, device):
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(x)
    elif not isinstance(x, torch.Tensor):
        x = torch.tensor(list(x))
    return x.to(device=device, dtype=torch.int32).contiguous()


class _CtcFromLogits(torch.autograd.Function):
    # diagnostics: when True, `last_fallback` receives the per-sequence int32 flags of the last call (1 = the sequence
    # was recomputed by the log-space recursion)
    record_fallback = False
    last_fallback = None

    @staticmethod
    def forward(ctx, logits, targets, input_lengths, target_lengths, max_target_len):
        if not logits.is_cuda:
            raise RuntimeError("hctr_b200 CTCLoss: logits must be a CUDA tensor (no CPU fallback)")
        if logits.dim() != 3:
            raise RuntimeError("hctr_b200 CTCLoss: expected [T,B,C] input, got %s" % (tuple(logits.shape),))
        if logits.dtype not in (torch.float32, torch.bfloat16):
            logits = logits.float()
        if logits.stride(2) != 1:
            logits = logits.contiguous()
        T, B, C = logits.shape
        dev = logits.device
        lib = nat.lib()
        need_grad = ctx.needs_input_grad[0]
        with torch.cuda.device(dev):
            nll = torch.empty((B,), dtype=torch.float32, device=dev)
            loss = torch.empty((1,), dtype=torch.float32, device=dev)
            grad = torch.empty_strided(logits.shape, logits.stride(), dtype=logits.dtype, device=dev) if need_grad else None
            ws_bytes = lib.hctr_ctc_loss_workspace_bytes(T, B, max_target_len)
            ws = torch.empty((ws_bytes + 256,), dtype=torch.uint8, device=dev)
            off = (-ws.data_ptr()) % 256
            ws = ws[off:off + ws_bytes]
            code = nat.HCTR_F32 if logits.dtype == torch.float32 else nat.HCTR_BF16
            nat.check(lib.hctr_ctc_loss_fwd_bwd(
                nat.ptr(logits), code, T, B, C, logits.stride(0), logits.stride(1), nat.ptr(targets),
                nat.ptr(target_lengths), nat.ptr(input_lengths), max_target_len, None, nat.ptr(nll), nat.ptr(loss),
                nat.ptr(grad), 1.0, nat.ptr(ws), ws_bytes, nat.stream_ptr()), "ctc_loss_fwd_bwd")
            if _CtcFromLogits.record_fallback:
                foff = lib.hctr_ctc_loss_flag_offset(T, B, max_target_len)
                _CtcFromLogits.last_fallback = ws[foff:foff + 4 * B].clone().view(torch.int32).cpu()
                _CtcFromLogits.last_workspace = ws
        ctx.grad = grad
        ctx.nll = nll
        return loss.reshape(())

    @staticmethod
    def backward(ctx, grad_out):
        g = ctx.grad
        if g is None:
            return None, None, None, None, None
        ctx.grad = None
        # in place: keeps the [B,W,pitch] kernel layout so the model's backward can consume the buffer without a copy
        return g.mul_(grad_out.to(g.dtype)), None, None, None, None


class CTCLoss(nn.Module):
    """Drop-in for `torch.nn.CTCLoss(blank=0, reduction='mean', zero_infinity=True)` as used by the reference."""

    def __init__(self, blank=0, reduction='mean', zero_infinity=False):
        super().__init__()
        if blank != 0:
            raise ValueError("hctr_b200 CTCLoss supports blank=0 only (the reference's setting)")
        if reduction != 'mean':
            raise ValueError("hctr_b200 CTCLoss supports reduction='mean' only (the reference's setting)")
        if not zero_infinity:
            raise ValueError("hctr_b200 CTCLoss implements zero_infinity=True (main.py:205)")
        self.blank, self.reduction, self.zero_infinity = blank, reduction, zero_infinity

    def forward(self, log_probs, targets, input_lengths, target_lengths):
        return self.from_logits(log_probs, targets, input_lengths, target_lengths)

    @staticmethod
    def from_logits(logits, targets, input_lengths, target_lengths):
        dev = logits.device
        tl_host = target_lengths.cpu() if isinstance(target_lengths, torch.Tensor) else torch.as_tensor(np.asarray(target_lengths))
        max_l = int(tl_host.max().item()) if tl_host.numel() else 0
        tg = _as_i32(targets, dev)
        if tg.dim() != 1:
            raise RuntimeError("hctr_b200 CTCLoss: targets must be the 1-D concatenated form the reference uses")
        return _CtcFromLogits.apply(logits, tg, _as_i32(input_lengths, dev), _as_i32(target_lengths, dev), max_l)
