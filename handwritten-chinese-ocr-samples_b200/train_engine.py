"""Train-mode forward and backward of the HCTR model on the sm_100a kernels (reference: main.py:367,383-438 with
models/handwritten_ctr_model.py in train(): batch-statistics BatchNorm, Dropout 0.1/0.3/0.3/0.3/0.9, SE, residuals).

Per conv+BN unit:
  forward   z = conv(x)+bias (tcgen05 implicit GEMM, bf16)  ->  chan_stats(z)  ->  bn_finalize (scale/shift, running
            stats)  ->  [se_excite]  ->  apply (affine, gate, residual, ReLU, (2,1) pool, dropout) in one pass
  backward  bwd_reduce  ->  bwd_finalize (BN/SE backward on [B,C] data)  ->  bwd_apply (dz, dres)  ->  wgrad (tcgen05,
            K = pixels)  ->  dgrad (the forward implicit GEMM with mirrored taps, residual gradient added in its epilogue)
Parameters stay fp32 `nn.Parameter`s in the reference layout; gradients are produced in that layout (optionally straight
into one flat buffer for the NCCL all-reduce + fused clip/SGD tail). No torch op touches an activation tensor.
"""
import os

import torch

from . import native as nat


def _mix_seed(base, k):
    x = (base + 0x9E3779B97F4A7C15 * (k + 1)) & 0xFFFFFFFFFFFFFFFF
    x ^= x >> 31
    x = (x * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    x ^= x >> 29
    return int(x & 0xFFFFFFFF)


_FOLD_STATS = int(os.environ.get("HCTR_TRAIN_FOLD_STATS", "1"))


class _Saved(object):
    """What one conv+BN unit keeps for its backward."""
    __slots__ = ("name", "x", "z", "mask", "scale", "shift", "mean", "invstd", "line_sum", "gate", "hidden", "se_mean", "res",
                 "relu", "pool", "drop_p", "seed", "B", "H", "W", "cin", "cout", "ksize", "stem", "se_name")


class TrainEngine(object):
    def __init__(self, model):
        self.model = model
        self.lib = nat.lib()
        self.dropout_enabled = True
        self.need_backward = True
        self._ones = {}
        self._zeros = {}
        self._tracked = {}            # id(bn) -> host copy of num_batches_tracked (only for momentum=None modules)
        self._pack = None             # packed bf16 weight operands of every tensor-core layer (rebuilt when parameters move)
        self._bn_seen = []            # BatchNorm modules of the running forward (their counters are bumped in one op)

    def _batches_tracked(self, bn):
        n = self._tracked.get(id(bn))
        if n is None:
            n = int(bn.num_batches_tracked.item()) if bn.num_batches_tracked is not None else 0
            self._tracked[id(bn)] = n
        return n

    # ------------------------------------------------------------------ small helpers
    def _const(self, cache, n, val, dev):
        key = (n, dev)
        t = cache.get(key)
        if t is None:
            t = torch.full((n,), val, dtype=torch.float32, device=dev)
            cache[key] = t
        return t

    def ones(self, n, dev):
        return self._const(self._ones, n, 1.0, dev)

    def zeros(self, n, dev):
        return self._const(self._zeros, n, 0.0, dev)

    @staticmethod
    def _ws(nbytes, dev):
        return torch.empty((max(int(nbytes), 16),), dtype=torch.uint8, device=dev)

    # ------------------------------------------------------------------ weight operands: one pack launch per step
    def _packed_layers(self):
        m = self.model
        cnn = m.cnn
        convs = [cnn.conv0_2]
        for stage in range(1, 5):
            for unit in getattr(cnn, "block%d" % stage):
                convs += [unit.conv1, unit.conv2]
                if unit.downsample is not None:
                    convs.append(unit.downsample[0])
            convs.append(getattr(cnn, "conv%d" % stage))
        return convs

    def pack_weights(self):
        """bf16 operand layouts of all 32 tensor-core convolutions and the classifier (forward [Cout][tap][Cin]; data
        gradient [Cin][tap][Cout], classifier [tap][Cin][pitch]) from the current fp32 parameters: ONE kernel over a
        descriptor table (hctr_pack_weights). The optimizer rewrites every parameter each step, so this runs per forward;
        the table and the two arenas are rebuilt only when a parameter's storage moves."""
        import ctypes
        m = self.model
        lin = m.linear
        convs = self._packed_layers()
        key = tuple(c.weight.data_ptr() for c in convs) + (lin.weight.data_ptr(),)
        pk = self._pack
        if pk is None or pk["key"] != key:
            dev = lin.weight.device
            n, d = lin.weight.shape
            hf = 4
            cf = d // hf
            pitch = (n + 7) // 8 * 8
            items = [(c.weight, c.weight.shape[0], c.weight.shape[1], c.weight.shape[2] * c.weight.shape[3], 0, 0) for c in convs]
            items.append((lin.weight, n, cf, hf, 1, pitch))
            fwd_total = sum(co * ci * t for _, co, ci, t, _, _ in items)
            bwd_total = sum((t * ci * p) if mode == 1 else (co * ci * t) for _, co, ci, t, mode, p in items)
            fwd = torch.empty((fwd_total,), dtype=torch.bfloat16, device=dev)
            bwd = torch.zeros((bwd_total,), dtype=torch.bfloat16, device=dev)     # classifier pad columns stay zero
            descs = (nat.PackDesc * len(items))()
            views_f, views_b = {}, {}
            fo = bo = tiles = 0
            for i, (w, co, ci, t, mode, p) in enumerate(items):
                nf = co * ci * t
                nb = t * ci * p if mode == 1 else nf
                descs[i].src = w.data_ptr()
                descs[i].dst_fwd = fwd.data_ptr() + 2 * fo
                descs[i].dst_bwd = bwd.data_ptr() + 2 * bo
                descs[i].cout, descs[i].cin, descs[i].taps, descs[i].bwd_mode = co, ci, t, mode
                descs[i].bwd_pitch = p
                descs[i].tile_start = tiles
                tiles += ((co + 31) // 32) * ((ci + 31) // 32)
                if mode == 1:
                    views_f[id(w)] = fwd[fo:fo + nf].view(co, t * ci)
                    views_b[id(w)] = bwd[bo:bo + nb].view(t * ci, p)
                else:
                    k = int(round(t ** 0.5))
                    views_f[id(w)] = fwd[fo:fo + nf].view(co, k, k, ci)
                    views_b[id(w)] = bwd[bo:bo + nb].view(ci, k, k, co)
                fo += nf
                bo += nb
            raw = torch.frombuffer(bytearray(bytes(descs)), dtype=torch.uint8).to(dev)
            pk = {"key": key, "fwd": fwd, "bwd": bwd, "descs": raw, "n": len(items), "tiles": tiles, "vf": views_f, "vb": views_b,
                  "pitch": pitch}
            self._pack = pk
        nat.check(self.lib.hctr_pack_weights(nat.ptr(pk["descs"]), pk["n"], pk["tiles"], nat.stream_ptr()), "pack_weights")
        return pk

    # ------------------------------------------------------------------ one conv + BN (+SE, +residual) unit
    def unit_forward(self, name, x, conv, bn, B, H, W, relu, pool, drop_p, seed, se=None, res=None, stem=False):
        lib, st, dev = self.lib, nat.stream_ptr(), conv.weight.device
        cout, cin, k = conv.weight.shape[0], conv.weight.shape[1], conv.weight.shape[2]
        bias = conv.bias.detach().float().contiguous() if conv.bias is not None else self.zeros(cout, dev)
        z = torch.empty((B, H, W, cout), dtype=torch.bfloat16, device=dev)
        # BatchNorm batch statistics: taken in the conv epilogue (per-(row, span, warp) sums of z and z*z as stored) where the
        # epilogue hides under the main loop - the CTA-pair layers, Cout >= 256 - and by a pass over z otherwise (the thin and 1x1
        # layers' epilogues are exposed; measured). HCTR_TRAIN_FOLD_STATS=0 / 2: never / every tensor-core conv.
        fold = (not stem) and (_FOLD_STATS == 2 or (_FOLD_STATS == 1 and k == 3 and cout >= 256 and H % 2 == 0))
        if stem:
            w = conv.weight.detach().float().reshape(cout, 9).contiguous()
            nat.check(lib.hctr_stem_conv_fwd(nat.ptr(x), nat.ptr(w), nat.ptr(self.ones(cout, dev)), nat.ptr(bias), nat.ptr(z),
                                             B, H, W, 0, st), "stem")
        else:
            w = self._pack["vf"][id(conv.weight)]
        if fold:
            slices = lib.hctr_conv_sum_slices(H, W, cin, cout, k)
            psum = torch.empty((B, slices, cout), dtype=torch.float32, device=dev)
            psq = torch.empty((B, slices, cout), dtype=torch.float32, device=dev)
            nat.check(lib.hctr_conv_stats_fwd(nat.ptr(x), nat.ptr(w), nat.ptr(self.ones(cout, dev)), nat.ptr(bias), nat.ptr(z),
                                              nat.ptr(psum), nat.ptr(psq), B, H, W, cin, cout, k, st), "conv_stats")
        else:
            if not stem:
                nat.check(lib.hctr_conv_bn_act_fwd(nat.ptr(x), nat.ptr(w), nat.ptr(self.ones(cout, dev)), nat.ptr(bias), nat.ptr(z),
                                                   B, H, W, cin, cout, k, 0, 0, st), "conv")
            slices = lib.hctr_stat_slices(B, H, W)
            psum = torch.empty((B, slices, cout), dtype=torch.float32, device=dev)
            psq = torch.empty((B, slices, cout), dtype=torch.float32, device=dev)
            nat.check(lib.hctr_chan_stats(nat.ptr(z), nat.ptr(psum), nat.ptr(psq), B, H, W, cout, st), "chan_stats")
        stats = torch.empty((4, cout), dtype=torch.float32, device=dev)         # mean, invstd, scale, shift
        line_sum = torch.empty((B, cout), dtype=torch.float32, device=dev)
        track = bn.track_running_stats and bn.running_mean is not None
        if bn.momentum is None:
            # torch: momentum=None is the cumulative moving average, factor 1/num_batches_tracked (counted from this batch);
            # the counter is a host-side integer here so that no device read sits on the launch path
            momentum = 1.0 / float(self._batches_tracked(bn) + 1)
        else:
            momentum = float(bn.momentum)
        nat.check(lib.hctr_bn_finalize_train(nat.ptr(psum), nat.ptr(psq), B, slices, cout, H * W, nat.ptr(bn.weight.detach()),
                                             nat.ptr(bn.bias.detach()), float(bn.eps), momentum,
                                             nat.ptr(bn.running_mean) if track else None,
                                             nat.ptr(bn.running_var) if track else None, nat.ptr(stats[0]), nat.ptr(stats[1]),
                                             nat.ptr(stats[2]), nat.ptr(stats[3]), nat.ptr(line_sum), st), "bn_finalize")
        if track:
            self._bn_seen.append(bn)
            if id(bn) in self._tracked:
                self._tracked[id(bn)] += 1
        m = self.model
        m.__dict__["_param_generation"] = m.__dict__.get("_param_generation", 0) + 1    # running stats changed by a kernel
        s = _Saved()
        s.name, s.x, s.z, s.res = name, x, z, res
        s.mean, s.invstd, s.scale, s.shift, s.line_sum = stats[0], stats[1], stats[2], stats[3], line_sum
        s.gate = s.hidden = s.se_mean = None
        s.se_name = None
        if se is not None:
            w1 = se.fc[0].weight.detach()
            w2 = se.fc[2].weight.detach()
            cr = w1.shape[0]
            s.gate = torch.empty((B, cout), dtype=torch.float32, device=dev)
            s.hidden = torch.empty((B, cr), dtype=torch.float32, device=dev)
            s.se_mean = torch.empty((B, cout), dtype=torch.float32, device=dev)
            nat.check(lib.hctr_se_excite_train(nat.ptr(line_sum), nat.ptr(s.scale), nat.ptr(s.shift), nat.ptr(w1), nat.ptr(w2),
                                               nat.ptr(s.se_mean), nat.ptr(s.hidden), nat.ptr(s.gate), B, cout, cr, H * W, st),
                      "se_excite_train")
        p = float(drop_p) if self.dropout_enabled else 0.0
        s.relu, s.pool, s.drop_p, s.seed = int(relu), int(pool), p, int(seed)
        s.B, s.H, s.W, s.cin, s.cout, s.ksize, s.stem = B, H, W, cin, cout, k, stem
        out = torch.empty((B, H // 2 if pool else H, W, cout), dtype=torch.bfloat16, device=dev)
        s.mask = torch.empty((B, H, W, cout // 8), dtype=torch.uint8, device=dev) if self.need_backward else None
        nat.check(lib.hctr_train_apply_fwd(nat.ptr(z), nat.ptr(s.scale), nat.ptr(s.shift), nat.ptr(s.gate), nat.ptr(res),
                                           nat.ptr(out), nat.ptr(s.mask), B, H, W, cout, s.relu, s.pool, p, s.seed, st),
                  "train_apply_fwd")
        s.res = res is not None          # the backward only needs to know whether a residual gradient is wanted
        return out, s

    def unit_backward(self, s, dout, conv, bn, grads, se=None, need_dx=True, add=None):
        """dout: gradient wrt the unit's output. Fills grads[...] for conv/bn(/se) and returns (dx, dres)."""
        lib, st, dev = self.lib, nat.stream_ptr(), conv.weight.device
        B, H, W, C = s.B, s.H, s.W, s.cout
        slices = lib.hctr_stat_slices(B, H, W)
        a2 = torch.empty((B, slices, C), dtype=torch.float32, device=dev)
        a3 = torch.empty((B, slices, C), dtype=torch.float32, device=dev)
        nat.check(lib.hctr_train_bwd_reduce(nat.ptr(dout), nat.ptr(s.z), nat.ptr(s.mask), nat.ptr(a2), nat.ptr(a3), B, H, W, C,
                                            s.pool, s.drop_p, st), "train_bwd_reduce")
        pq = torch.empty((2, B, C), dtype=torch.float32, device=dev)
        r = torch.empty((C,), dtype=torch.float32, device=dev)
        prefix = s.name
        g_gamma, g_beta = grads[prefix[1] + ".weight"], grads[prefix[1] + ".bias"]
        g_bias = grads.get(prefix[0] + ".bias")
        w1 = w2 = dw1 = dw2 = None
        cr = 0
        if se is not None:
            w1, w2 = se.fc[0].weight.detach(), se.fc[2].weight.detach()
            cr = w1.shape[0]
            dw1, dw2 = grads[prefix[2] + ".fc.0.weight"], grads[prefix[2] + ".fc.2.weight"]
        nat.check(lib.hctr_train_bwd_finalize(
            nat.ptr(a2), nat.ptr(a3), slices, B, C, H * W, nat.ptr(bn.weight.detach()), nat.ptr(s.mean), nat.ptr(s.invstd),
            nat.ptr(s.scale), nat.ptr(s.shift), nat.ptr(s.line_sum), nat.ptr(s.gate), nat.ptr(s.hidden), nat.ptr(s.se_mean),
            nat.ptr(w1), nat.ptr(w2), cr, nat.ptr(dw1), nat.ptr(dw2), nat.ptr(g_gamma), nat.ptr(g_beta), nat.ptr(g_bias),
            nat.ptr(pq[0]), nat.ptr(pq[1]), nat.ptr(r), st), "train_bwd_finalize")
        dz = torch.empty((B, H, W, C), dtype=torch.bfloat16, device=dev)
        dres = torch.empty((B, H, W, C), dtype=torch.bfloat16, device=dev) if s.res else None
        nat.check(lib.hctr_train_bwd_apply(nat.ptr(dout), nat.ptr(s.z), nat.ptr(s.mask), nat.ptr(pq[0]), nat.ptr(pq[1]), nat.ptr(r),
                                           nat.ptr(dz), nat.ptr(dres), B, H, W, C, s.pool, s.drop_p, st), "train_bwd_apply")
        g_w = grads[prefix[0] + ".weight"]
        if s.stem:
            nb = lib.hctr_stem_wgrad_workspace_bytes(B, H, W)
            ws = self._ws(nb, dev)
            nat.check(lib.hctr_stem_wgrad(nat.ptr(dz), nat.ptr(s.x), nat.ptr(g_w), B, H, W, nat.ptr(ws), nb, st), "stem_wgrad")
            return None, dres
        nb = lib.hctr_wgrad_workspace_bytes(B, H, W, C, s.cin, s.ksize * s.ksize)
        ws = self._ws(nb, dev)
        nat.check(lib.hctr_conv_wgrad(nat.ptr(dz), nat.ptr(s.x), nat.ptr(g_w), B, H, W, C, s.cin, s.ksize, nat.ptr(ws), nb, st),
                  "conv_wgrad")
        dx = None
        if need_dx:
            wt = self._pack["vb"][id(conv.weight)]                                               # [Cin][kh][kw][Cout]
            dx = torch.empty((B, H, W, s.cin), dtype=torch.bfloat16, device=dev)
            nat.check(lib.hctr_conv_dgrad(nat.ptr(dz), nat.ptr(wt), nat.ptr(self.ones(s.cin, dev)), nat.ptr(self.zeros(s.cin, dev)),
                                          nat.ptr(add), nat.ptr(dx), B, H, W, C, s.cin, s.ksize, st), "conv_dgrad")
        return dx, dres

    # ------------------------------------------------------------------ whole model
    def forward(self, x, base_seed):
        """x: fp32 [B,1,128,W] CUDA. Returns (logits [B,W,pitch] bf16/fp32 buffer, ctx)."""
        m = self.model
        cnn = m.cnn
        B, _, H, W = x.shape
        ctx = {"units": {}, "B": B, "W": W, "x": x}
        k = [0]
        pk = self.pack_weights()
        self._bn_seen = []

        def seed():
            k[0] += 1
            return _mix_seed(base_seed, k[0])

        a, s = self.unit_forward(("cnn.conv0_1", "cnn.bn0_1"), x, cnn.conv0_1, cnn.bn0_1, B, H, W, True, False, 0.0, 0, stem=True)
        ctx["units"]["0_1"] = s
        a, s = self.unit_forward(("cnn.conv0_2", "cnn.bn0_2"), a, cnn.conv0_2, cnn.bn0_2, B, H, W, True, True, 0.0, 0)
        ctx["units"]["0_2"] = s
        H //= 2
        stage_drop = (0.3, 0.3, 0.3, 0.9)                     # dropout1..4 (models/handwritten_ctr_model.py:96-99)
        for stage in range(1, 5):
            blocks = getattr(cnn, "block%d" % stage)
            for i, unit in enumerate(blocks):
                base = "cnn.block%d.%d" % (stage, i)
                t, s1 = self.unit_forward((base + ".conv1", base + ".bn1"), a, unit.conv1, unit.bn1, B, H, W, True, False, 0.0, 0)
                res, ssc = a, None
                if unit.downsample is not None:
                    res, ssc = self.unit_forward((base + ".downsample.0", base + ".downsample.1"), a, unit.downsample[0],
                                                 unit.downsample[1], B, H, W, False, False, 0.0, 0)
                # conv2 -> bn2 -> SE -> + residual -> ReLU -> Dropout(0.1)   (:52-59)
                out, s2 = self.unit_forward((base + ".conv2", base + ".bn2", base + ".se"), t, unit.conv2, unit.bn2, B, H, W,
                                            True, False, 0.1, seed(), se=unit.se, res=res)
                ctx["units"][base] = (s1, s2, ssc)
                a = out
            conv, bn = getattr(cnn, "conv%d" % stage), getattr(cnn, "bn%d" % stage)
            a, s = self.unit_forward(("cnn.conv%d" % stage, "cnn.bn%d" % stage), a, conv, bn, B, H, W, True, True,
                                     stage_drop[stage - 1], seed())
            ctx["units"]["tail%d" % stage] = s
            H //= 2
        # classifier (bf16 logits with a 16-byte aligned pitch; fp32 accumulate)
        n = m.noutput
        pitch = (n + 7) // 8 * 8
        lin = m.linear
        cf = lin.weight.shape[1] // H
        wk = pk["vf"][id(lin.weight)]                       # [n][h*cf + c]: the reference's flatten gives d = c*4 + h
        logits = torch.empty((B, W, pitch), dtype=torch.bfloat16, device=x.device)
        # classifier GEMM fused with log_softmax: the epilogue also yields the row log-sum-exp the CTC loss needs
        row_lse = torch.empty((B, W), dtype=torch.float32, device=x.device)
        nb = self.lib.hctr_classifier_lse_workspace_bytes(B, W, n)
        ws = self._ws(nb, x.device)
        nat.check(self.lib.hctr_classifier_lse_fwd(nat.ptr(a), nat.ptr(wk), nat.ptr(lin.bias.detach()), nat.ptr(logits),
                                                   nat.HCTR_BF16, pitch, B, H, W, cf, n, nat.ptr(row_lse), nat.ptr(ws), nb,
                                                   nat.stream_ptr()), "classifier_lse")
        ctx["feat"], ctx["Hf"], ctx["cf"], ctx["pitch"], ctx["wk"] = a, H, cf, pitch, wk
        ctx["row_lse"] = row_lse
        counters = [bn.num_batches_tracked for bn in self._bn_seen if bn.num_batches_tracked is not None]
        if counters:
            torch._foreach_add_(counters, 1)                # 33 counters, one launch
        self._bn_seen = []
        return logits, ctx

    def backward(self, ctx, dlogits, grads, on_stage_done=None):
        """dlogits: bf16 [B,W,pitch] gradient buffer; grads: dict name -> fp32 tensor (reference layouts), filled here.
        on_stage_done(name) is called after the kernels producing the last gradient of each parameter group
        (linear, stage4..stage1, stage0) have been enqueued - the hook the bucketed all-reduce hangs on."""
        notify = on_stage_done if on_stage_done is not None else (lambda name: None)
        m = self.model
        cnn = m.cnn
        lib, st = self.lib, nat.stream_ptr()
        B, W, Hf, cf, pitch = ctx["B"], ctx["W"], ctx["Hf"], ctx["cf"], ctx["pitch"]
        n = m.noutput
        dev = dlogits.device
        feat = ctx["feat"]
        # ---- classifier: dW, db, dfeat
        nb = lib.hctr_linear_wgrad_workspace_bytes(B, Hf, W, cf, n)
        ws = self._ws(nb, dev)
        nat.check(lib.hctr_linear_wgrad(nat.ptr(dlogits), pitch, nat.ptr(feat), nat.ptr(grads["linear.weight"]), B, Hf, W, cf, n,
                                        nat.ptr(ws), nb, st), "linear_wgrad")
        nb = lib.hctr_colsum_workspace_bytes(B * W, n)
        ws2 = self._ws(nb, dev)
        nat.check(lib.hctr_colsum_bf16(nat.ptr(dlogits), B * W, n, pitch, nat.ptr(grads["linear.bias"]), nat.ptr(ws2), nb, st), "colsum")
        wt = self._pack["vb"][id(m.linear.weight)]            # [h*cf + c][pitch], packed with the forward operands
        d = torch.empty((B, Hf, W, cf), dtype=torch.bfloat16, device=dev)
        nat.check(lib.hctr_classifier_dgrad(nat.ptr(dlogits), pitch, nat.ptr(wt), nat.ptr(self.ones(cf, dev)),
                                            nat.ptr(self.zeros(cf, dev)), nat.ptr(d), B, Hf, W, cf, n, st), "classifier_dgrad")
        notify("linear")
        # ---- backbone, last stage first
        for stage in range(4, 0, -1):
            s = ctx["units"]["tail%d" % stage]
            d, _ = self.unit_backward(s, d, getattr(cnn, "conv%d" % stage), getattr(cnn, "bn%d" % stage), grads)
            blocks = getattr(cnn, "block%d" % stage)
            for i in range(len(blocks) - 1, -1, -1):
                unit = blocks[i]
                base = "cnn.block%d.%d" % (stage, i)
                s1, s2, ssc = ctx["units"][base]
                dt, dres = self.unit_backward(s2, d, unit.conv2, unit.bn2, grads, se=unit.se)
                if ssc is not None:
                    dsc, _ = self.unit_backward(ssc, dres, unit.downsample[0], unit.downsample[1], grads)
                    d, _ = self.unit_backward(s1, dt, unit.conv1, unit.bn1, grads, add=dsc)
                else:
                    d, _ = self.unit_backward(s1, dt, unit.conv1, unit.bn1, grads, add=dres)
            notify("stage%d" % stage)
        d, _ = self.unit_backward(ctx["units"]["0_2"], d, cnn.conv0_2, cnn.bn0_2, grads)
        self.unit_backward(ctx["units"]["0_1"], d, cnn.conv0_1, cnn.bn0_1, grads, need_dx=False)
        notify("stage0")


class _TrainFunction(torch.autograd.Function):
    """Autograd bridge: logits = f(input, *parameters); backward runs the engine and hands back parameter gradients in
    the reference layout, so `loss.backward(); optimizer.step()` of the reference training loop works unchanged."""

    @staticmethod
    def forward(ctx, engine, x, base_seed, *params):
        logits, saved = engine.forward(x, base_seed)
        ctx.engine, ctx.saved = engine, saved
        n = engine.model.noutput
        out = logits[:, :, :n].permute(1, 0, 2)               # [W,B,C] view, as the reference returns (:176)
        ctx.logits_shape = logits.shape
        return out

    @staticmethod
    def backward(ctx, grad_out):
        engine, saved = ctx.engine, ctx.saved
        B, W, pitch = ctx.logits_shape
        n = engine.model.noutput
        if (grad_out.dtype == torch.bfloat16 and tuple(grad_out.stride()) == (pitch, W * pitch, 1)
                and grad_out.storage_offset() % 8 == 0):
            buf = grad_out.permute(1, 0, 2)           # [B,W,n] view with row pitch `pitch`: already the kernel layout (CTCLoss)
        else:
            buf = torch.zeros((B, W, pitch), dtype=torch.bfloat16, device=grad_out.device)
            buf[:, :, :n] = grad_out.permute(1, 0, 2)
        names = [k for k, _ in engine.model.named_parameters()]
        params = [p for _, p in engine.model.named_parameters()]
        grads = {k: torch.empty_like(p, dtype=torch.float32) for k, p in zip(names, params)}
        with torch.cuda.device(grad_out.device):
            engine.backward(saved, buf, grads)
        ctx.saved = None
        return (None, None, None) + tuple(grads[k] for k in names)
