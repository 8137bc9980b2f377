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


def broadcast_state(tensors, src=0, group=None):
    """Host logic of DDP's construction-time sync (main.py:237: DistributedDataParallel broadcasts rank 0's parameters
    and buffers): every tensor in `tensors` is overwritten in place with rank `src`'s. No-op without a process group."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return 0
    n = 0
    for t in tensors:
        dist.broadcast(t, src=src, group=group)
        n += 1
    return n


def sgd_state_dict(names, offsets, shapes, momentum_flat, group, have_momentum):
    """Host logic: the flat momentum buffer as a `torch.optim.SGD.state_dict()` (the 'optimizer' entry of the reference's
    checkpoints, main.py:348 / :262): parameters are numbered in registration order, each with a `momentum_buffer` of the
    parameter's shape (absent before the first step, as in torch)."""
    state = {}
    if have_momentum:
        for i, k in enumerate(names):
            s, n = offsets[k]
            state[i] = {"momentum_buffer": momentum_flat[s:s + n].detach().clone().view(shapes[k])}
    pg = {k: v for k, v in group.items() if k != "params"}
    pg["params"] = list(range(len(names)))
    return {"state": state, "param_groups": [pg]}


def load_sgd_state_dict(sd, names, offsets, shapes, momentum_flat, group):
    """Inverse of sgd_state_dict: accepts the state_dict of a torch.optim.SGD over the same parameters (same order).
    Returns True when momentum buffers were present (the next step is then not a 'first step')."""
    groups = sd["param_groups"]
    if len(groups) != 1 or len(groups[0]["params"]) != len(names):
        raise ValueError("loaded state dict has a different number of parameter groups / parameters "
                         "(expected 1 group of %d)" % len(names))
    if groups[0].get("nesterov") or groups[0].get("dampening", 0) not in (0, 0.0):
        raise ValueError("hctr_b200 TrainStep: nesterov / dampening are not implemented (the reference uses neither)")
    for k in ("lr", "momentum", "weight_decay"):
        if k in groups[0]:
            group[k] = groups[0][k]
    ids = groups[0]["params"]
    state = sd.get("state", {})
    have = False
    momentum_flat.zero_()
    for i, k in enumerate(names):
        st = state.get(ids[i], state.get(str(ids[i])))
        if st is None or st.get("momentum_buffer") is None:
            continue
        buf = st["momentum_buffer"]
        if tuple(buf.shape) != tuple(shapes[k]):
            raise ValueError("momentum_buffer of parameter %d (%s) has shape %s, expected %s"
                             % (i, k, tuple(buf.shape), tuple(shapes[k])))
        s, n = offsets[k]
        momentum_flat[s:s + n].copy_(buf.reshape(-1).to(momentum_flat.dtype))
        have = True
    return have


def adjust_learning_rate(optimizer, epoch, args):
    """The reference's schedule (main.py:579-584): initial LR decayed by 10 every 30 epochs. Works on a TrainStep or
    any torch optimizer (both expose `param_groups`)."""
    lr = args.lr * (0.1 ** (epoch // 30))
    for param_group in optimizer.param_groups:
        param_group['lr'] = lr


_PIN_RING = 3


def _bump_generation(model):
    """Tell the model that parameters or buffers were updated through raw pointers (no torch `_version` bump): the
    eval-mode plan (folded BN, packed weights) must be rebuilt on the next eval forward."""
    model.__dict__["_param_generation"] = model.__dict__.get("_param_generation", 0) + 1


class TrainStep(object):
    def __init__(self, model, lr=1e-3, momentum=0.9, weight_decay=1e-4, max_norm=5.0, process_group=None):
        self.model = model
        self.max_norm = max_norm
        self.group = process_group
        self.engine = TrainEngine(model)
        named = [(k, p) for k, p in model.named_parameters()]
        dev = named[0][1].device
        if dev.type != "cuda":
            raise RuntimeError("hctr_b200 TrainStep: parameters must live on a CUDA device (no CPU fallback)")
        self.offsets, total, self.buckets = plan_buckets([(k, p.numel()) for k, p in named])
        self.names = [k for k, _ in named]
        self.shapes = {k: tuple(p.shape) for k, p in named}
        # one parameter group, laid out like torch.optim.SGD's so that `for g in optimizer.param_groups: g['lr'] = ...`
        # (main.py:582-583) steers the fused update
        self.param_groups = [{"lr": lr, "momentum": momentum, "dampening": 0, "weight_decay": weight_decay,
                              "nesterov": False, "params": [p for _, p in named]}]
        self.flat_params = torch.zeros((total,), dtype=torch.float32, device=dev)
        self.flat_grads = torch.zeros((total,), dtype=torch.float32, device=dev)
        self.momentum_buf = torch.zeros((total,), dtype=torch.float32, device=dev)
        self.grad_views = {}
        for k, p in named:
            s, n = self.offsets[k]
            self.flat_params[s:s + n].copy_(p.detach().reshape(-1))
            p.data = self.flat_params[s:s + n].view(p.shape)                  # same Parameter object, new storage
            self.grad_views[k] = self.flat_grads[s:s + n].view(p.shape)
        self.norm = torch.zeros((4,), dtype=torch.float32, device=dev)        # {total_norm, clip coefficient, skipped, -}
        self._ws = torch.empty((nat.lib().hctr_sgd_workspace_bytes(),), dtype=torch.uint8, device=dev)
        self.steps = 0
        self._have_momentum = False
        # label staging: a ring of pinned buffers, each guarded by the event of the H2D copy that last read it (the host
        # runs ahead of the GPU; rewriting a single buffer would hand step N the labels of step N+1)
        self._pins = [None] * _PIN_RING
        self._pin_events = [None] * _PIN_RING
        self._pin_next = 0
        self.world = dist.get_world_size(process_group) if (dist.is_available() and dist.is_initialized()) else 1
        if self.world > 1:
            self.sync_replicas()

    def sync_replicas(self, src=0):
        """What DistributedDataParallel does at construction (main.py:237): every replica starts from rank `src`'s
        parameters and buffers (the reference's --seed defaults to None, so unseeded ranks would otherwise average
        gradients into diverging weights). Also called after load_state_dict so that momentum is replicated too."""
        if self.world <= 1:
            return
        broadcast_state([self.flat_params, self.momentum_buf] + list(self.model.buffers()), src, self.group)
        _bump_generation(self.model)

    @property
    def skipped(self):
        """Device flag (0-d fp32 tensor): 1 when the last step met a non-finite gradient norm and left parameters and
        momentum untouched - the reference's `if not torch.isfinite(loss): continue` / GradScaler skip (main.py:413,433)."""
        return self.norm[2]

    def _stage_labels(self, tg_host, tl_host, dev):
        n_t, B = tg_host.numel(), tl_host.numel()
        i = self._pin_next
        self._pin_next = (i + 1) % _PIN_RING
        if self._pin_events[i] is not None:
            self._pin_events[i].synchronize()          # the copy that last read this buffer has run
        if self._pins[i] is None or self._pins[i].numel() < n_t + B:
            self._pins[i] = torch.empty((max(4096, 2 * (n_t + B)),), dtype=torch.int32).pin_memory()
        pin = self._pins[i]
        pin[:n_t].copy_(tg_host.reshape(-1))
        pin[n_t:n_t + B].copy_(tl_host.to(torch.int32).reshape(-1))
        staged = pin[:n_t + B].to(dev, non_blocking=True)
        ev = self._pin_events[i] or torch.cuda.Event()
        ev.record()
        self._pin_events[i] = ev
        return staged[:n_t], staged[n_t:n_t + B]

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
            if tg_host.is_cuda:
                tg_host = tg_host.cpu()
            if tl_host.numel() != B:
                raise ValueError("target_lengths has %d entries for a batch of %d" % (tl_host.numel(), B))
            if int(tl_host.sum().item()) != tg_host.numel() or int(tl_host.min().item()) < 0:
                raise ValueError("targets holds %d labels but target_lengths sums to %d"
                                 % (tg_host.numel(), int(tl_host.sum().item())))
            tg, tl = self._stage_labels(tg_host, tl_host, dev)
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
                                             self.weight_decay, int(not self._have_momentum), nat.ptr(self.norm), nat.ptr(self._ws), st),
                      "sgd_clip_step")
            self.steps += 1
            self._have_momentum = True
            _bump_generation(m)              # parameters / running stats changed behind torch's back (raw pointers)
        return loss.reshape(())

    # hyper-parameters live in param_groups[0] (torch.optim convention)
    lr = property(lambda self: float(self.param_groups[0]["lr"]),
                  lambda self, v: self.param_groups[0].__setitem__("lr", v))
    momentum = property(lambda self: float(self.param_groups[0]["momentum"]),
                        lambda self, v: self.param_groups[0].__setitem__("momentum", v))
    weight_decay = property(lambda self: float(self.param_groups[0]["weight_decay"]),
                            lambda self, v: self.param_groups[0].__setitem__("weight_decay", v))

    def state_dict(self):
        """Same layout as `torch.optim.SGD(model.parameters(), ...).state_dict()`, i.e. what the reference stores under
        checkpoint['optimizer'] (main.py:348) and restores at :262."""
        return sgd_state_dict(self.names, self.offsets, self.shapes, self.momentum_buf, self.param_groups[0],
                              self._have_momentum)

    def load_state_dict(self, sd):
        self._have_momentum = load_sgd_state_dict(sd, self.names, self.offsets, self.shapes, self.momentum_buf,
                                                  self.param_groups[0])
        broadcast_state([self.momentum_buf], 0, self.group)

    def zero_grad(self, set_to_none=False):
        """Kept for call-site compatibility (main.py:425): the kernels overwrite the flat gradient buffer every step."""
        return None
