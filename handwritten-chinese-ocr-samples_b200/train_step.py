"""The reference training step (main.py:359-438) as one B200-native call:

    forward (train mode) -> log_softmax + CTCLoss(zero_infinity) fwd/bwd -> backward -> gradient all-reduce (DDP average,
    main.py:222-237) -> clip_grad_norm_(5.0) -> SGD(momentum 0.9, weight_decay 1e-4)          [bf16 activations, fp32 master]

Parameters are re-pointed at views of ONE flat fp32 buffer (same nn.Parameters, same state_dict), gradients are written
by the kernels straight into a second flat buffer, so the data-parallel exchange is a handful of NCCL all-reduces over
contiguous ranges launched as each stage of the backward finishes (overlapping the remaining backward kernels on
NVLink/NVSwitch), and the optimizer tail is one fused pass (global norm, clip coefficient, momentum update).
"""
import torch
import torch.distributed as dist

from . import native as nat
from .train_engine import TrainEngine


def plan_buckets(named_numels, align=4):
    """Host logic: lay parameters out in registration order with `align`-element alignment and cut the flat range into
    the buckets that complete together during the backward: [stem+conv0_2], [stage1], [stage2], [stage3], [stage4],
    [linear]. Returns (offsets {name: (start, numel)}, total, buckets [(name, start, end)] in completion order)."""
    offsets, cur = {}, 0
    groups = {}
    order = []
    for name, n in named_numels:
        start = cur
        offsets[name] = (start, n)
        cur = (start + n + align - 1) // align * align
        if name.startswith("linear."):
            key = "linear"
        else:
            digits = [ch for ch in name.split(".")[1] if ch.isdigit()]
            key = "stage%s" % digits[0] if digits else "stage0"
        if key not in groups:
            groups[key] = [start, cur]
            order.append(key)
        groups[key][1] = cur
    buckets = [(k, groups[k][0], groups[k][1]) for k in reversed(order)]       # the backward finishes the last group first
    return offsets, cur, buckets


def allreduce_bucket(flat, start, end, group=None, async_op=False):
    """Sum-all-reduce of one contiguous gradient range (the division by world size is folded into the optimizer)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return None
    return dist.all_reduce(flat[start:end], op=dist.ReduceOp.SUM, group=group, async_op=async_op)


class TrainStep(object):
    def __init__(self, model, lr=1e-3, momentum=0.9, weight_decay=1e-4, max_norm=5.0, process_group=None):
        self.model = model
        self.lr, self.momentum, self.weight_decay, self.max_norm = lr, momentum, weight_decay, max_norm
        self.group = process_group
        self.engine = TrainEngine(model)
        named = [(k, p) for k, p in model.named_parameters()]
        dev = named[0][1].device
        if dev.type != "cuda":
            raise RuntimeError("hctr_b200 TrainStep: parameters must live on a CUDA device (no CPU fallback)")
        self.offsets, total, self.buckets = plan_buckets([(k, p.numel()) for k, p in named])
        self.flat_params = torch.zeros((total,), dtype=torch.float32, device=dev)
        self.flat_grads = torch.zeros((total,), dtype=torch.float32, device=dev)
        self.momentum_buf = torch.zeros((total,), dtype=torch.float32, device=dev)
        self.grad_views = {}
        for k, p in named:
            s, n = self.offsets[k]
            self.flat_params[s:s + n].copy_(p.detach().reshape(-1))
            p.data = self.flat_params[s:s + n].view(p.shape)                  # same Parameter object, new storage
            self.grad_views[k] = self.flat_grads[s:s + n].view(p.shape)
        self.norm = torch.zeros((2,), dtype=torch.float32, device=dev)        # {total_norm, clip coefficient}
        self._ws = torch.empty((nat.lib().hctr_sgd_workspace_bytes(),), dtype=torch.uint8, device=dev)
        self.steps = 0
        self._pin = None
        self.world = dist.get_world_size(process_group) if (dist.is_available() and dist.is_initialized()) else 1

    def step(self, x, targets, target_lengths, seed=None):
        """x: fp32 [B,1,128,W] CUDA; targets: int32 concatenated labels; target_lengths: int32 [B] (host or device).
        Returns the per-rank mean CTC loss as a 0-d device tensor (the reference does not reduce it across ranks)."""
        m, lib = self.model, nat.lib()
        dev = x.device
        with torch.cuda.device(dev), torch.no_grad():
            st = nat.stream_ptr()
            if seed is None:
                seed = int(torch.randint(0, 2 ** 62, (1,)).item())
            self.engine.dropout_enabled = bool(getattr(m, "dropout_enabled", True))
            self.engine.need_backward = True
            logits, ctx = self.engine.forward(x.detach().float().contiguous(), seed)
            B, W, pitch = logits.shape
            C = m.noutput
            tl_host = target_lengths.cpu() if isinstance(target_lengths, torch.Tensor) else torch.as_tensor(target_lengths)
            max_l = int(tl_host.max().item())
            # labels travel through pinned staging buffers with async copies: a pageable H2D would block the host until
            # every kernel of the previous step has finished and serialise CPU enqueue with GPU execution
            tg_host = torch.as_tensor(targets).to(torch.int32)
            n_t = tg_host.numel()
            if self._pin is None or self._pin.numel() < n_t + B:
                self._pin = torch.empty((max(4096, 2 * (n_t + B)),), dtype=torch.int32).pin_memory()
            self._pin[:n_t].copy_(tg_host.reshape(-1))
            self._pin[n_t:n_t + B].copy_(tl_host.to(torch.int32).reshape(-1))
            staged = self._pin[:n_t + B].to(dev, non_blocking=True)
            tg, tl = staged[:n_t], staged[n_t:n_t + B]
            il = torch.full((B,), W, dtype=torch.int32, device=dev)            # preds_sizes = [T]*B (main.py:388)
            nll = torch.empty((B,), dtype=torch.float32, device=dev)
            loss = torch.empty((1,), dtype=torch.float32, device=dev)
            dlogits = torch.empty_like(logits)
            nb = lib.hctr_ctc_loss_workspace_bytes(W, B, max_l)
            ws = torch.empty((nb + 256,), dtype=torch.uint8, device=dev)
            off = (-ws.data_ptr()) % 256
            nat.check(lib.hctr_ctc_loss_fwd_bwd(nat.ptr(logits), nat.HCTR_BF16, W, B, C, pitch, W * pitch, nat.ptr(tg), nat.ptr(tl),
                                                nat.ptr(il), max_l, nat.ptr(ctx["row_lse"]), nat.ptr(nll), nat.ptr(loss), nat.ptr(dlogits), 1.0,
                                                nat.c_void_p(ws.data_ptr() + off), nb, st), "ctc_loss_fwd_bwd")
            works = []
            done = [0]

            def stage_done(_name):
                # the parameters of bucket `done` have their final gradients enqueued: start its all-reduce now
                name, s, e = self.buckets[done[0]]
                done[0] += 1
                w = allreduce_bucket(self.flat_grads, s, e, self.group, async_op=True)
                if w is not None:
                    works.append(w)

            self.engine.backward(ctx, dlogits, self.grad_views, on_stage_done=stage_done)
            for w in works:
                w.wait()
            nat.check(lib.hctr_sgd_clip_step(nat.ptr(self.flat_params), nat.ptr(self.flat_grads), nat.ptr(self.momentum_buf),
                                             self.flat_params.numel(), 1.0 / self.world, self.max_norm, self.lr, self.momentum,
                                             self.weight_decay, int(self.steps == 0), nat.ptr(self.norm), nat.ptr(self._ws), st),
                      "sgd_clip_step")
            self.steps += 1
        return loss.reshape(())
