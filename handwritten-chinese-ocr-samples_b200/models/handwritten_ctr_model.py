"""B200-native drop-in for the reference recognition model.

Mirrors `models/handwritten_ctr_model.py` of the reference (file:line cited per item):
  * `hctr_model(num_classes=7375)` with attributes img_height/PAD/optimizer/pred/noutput (:156-169)
  * `forward(input[B,1,128,W]) -> [W,B,num_classes]` (:171-178)
  * the exact 254-entry `state_dict` (names, shapes, registration order) and the same
    construction order, so `torch.manual_seed(s); hctr_model()` yields bit-identical parameters.

The nn.Modules below only *hold* parameters; no torch op runs in `forward`. The arithmetic is the
hand-written sm_100a path behind include/hctr_b200.h: a CUDA-core stem, tcgen05 implicit-GEMM
convolutions with bias/BN/ReLU/(2,1)-max-pool folded into the epilogue, SE squeeze/excite/apply
passes and the tcgen05 classifier GEMM. There is no CPU fallback.
"""
import importlib.util
import os
import sys

import torch
import torch.nn as nn

_PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _core():
    """Return the `hctr_b200` package even when this file was imported as top-level `models.*`
    (i.e. with the package directory itself on PYTHONPATH, the drop-in arrangement)."""
    pkg = sys.modules.get("hctr_b200")
    if pkg is None:
        spec = importlib.util.spec_from_file_location(
            "hctr_b200", os.path.join(_PKG_DIR, "__init__.py"), submodule_search_locations=[_PKG_DIR])
        pkg = importlib.util.module_from_spec(spec)
        sys.modules["hctr_b200"] = pkg
        spec.loader.exec_module(pkg)
    return pkg


_STAGE_PLANES = (128, 256, 512, 512)      # reference: ResNet.__init__ inout_channels (:66-70)
_STAGE_BLOCKS = (2, 4, 5, 1)              # reference: hctr_model.__init__ (:166)
_SE_REDUCTION = 16                        # reference: SELayer default (:16)
_FEATURE_ROWS = 4                         # 128 / 2**5 after five (2,1) pools (:123-150)


class _SqueezeExcite(nn.Module):
    """Parameter holder for SELayer (:11-24): fc.0 = Linear(C, C/16), fc.2 = Linear(C/16, C), no biases."""

    def __init__(self, channels):
        super().__init__()
        hidden = channels // _SE_REDUCTION
        self.fc = nn.Sequential(
            nn.Linear(channels, hidden, bias=False), nn.ReLU(inplace=True),
            nn.Linear(hidden, channels, bias=False), nn.Sigmoid())


class _ResidualUnit(nn.Module):
    """Parameter holder for BasicBlock (:33-45); registration order conv1,bn1,conv2,bn2,se,downsample."""

    def __init__(self, cin, cout, shortcut):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, cout, 3, 1, 1)
        self.bn1 = nn.BatchNorm2d(cout)
        self.conv2 = nn.Conv2d(cout, cout, 3, 1, 1)
        self.bn2 = nn.BatchNorm2d(cout)
        self.se = _SqueezeExcite(cout)
        self.downsample = shortcut


class _Backbone(nn.Module):
    """Parameter holder for ResNet(1, 512, BasicBlock, [2,4,5,1]) (:63-99)."""

    def __init__(self):
        super().__init__()
        width = 64
        self.conv0_1 = nn.Conv2d(1, width, 3, 1, 1)
        self.bn0_1 = nn.BatchNorm2d(width)
        self.conv0_2 = nn.Conv2d(width, width, 3, 1, 1)
        self.bn0_2 = nn.BatchNorm2d(width)
        for stage, (planes, count) in enumerate(zip(_STAGE_PLANES, _STAGE_BLOCKS), start=1):
            # the reference builds the projection shortcut *before* the block's own convs
            # (_make_block, :101-113) - the RNG draw order matters for seed-identical parameters
            shortcut = None
            if width != planes:
                shortcut = nn.Sequential(nn.Conv2d(width, planes, kernel_size=1, stride=1, bias=False),
                                         nn.BatchNorm2d(planes))
            units = [_ResidualUnit(width, planes, shortcut)]
            width = planes
            units += [_ResidualUnit(width, planes, None) for _ in range(1, count)]
            setattr(self, "block%d" % stage, nn.Sequential(*units))
            setattr(self, "conv%d" % stage, nn.Conv2d(planes, planes, 3, 1, 1))
            setattr(self, "bn%d" % stage, nn.BatchNorm2d(planes))


def _fold_bn(bn, conv_bias):
    """Eval-mode BN as an fp32 epilogue: y = acc*scale + shift."""
    scale = bn.weight.detach().float() * torch.rsqrt(bn.running_var.detach().float() + bn.eps)
    bias = conv_bias.detach().float() if conv_bias is not None else torch.zeros_like(scale)
    shift = (bias - bn.running_mean.detach().float()) * scale + bn.bias.detach().float()
    return scale.contiguous(), shift.contiguous()


def _pack_conv(conv):
    """OIHW fp32 -> [Cout][kh][kw][Cin] bf16 (K-major rows for the implicit GEMM's B operand)."""
    return conv.weight.detach().permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


class _ConvSpec(object):
    __slots__ = ("w", "scale", "shift", "cin", "cout", "ksize")

    def __init__(self, conv, bn):
        self.w = _pack_conv(conv)
        self.scale, self.shift = _fold_bn(bn, conv.bias)
        self.cout, self.cin = conv.weight.shape[0], conv.weight.shape[1]
        self.ksize = conv.weight.shape[2]


class _InferencePlan(object):
    """Device-resident packed parameters for the eval-mode forward (rebuilt when parameters change)."""

    def __init__(self, model):
        cnn = model.cnn
        dev = model.linear.weight.device
        self.device = dev
        self.stem_w = cnn.conv0_1.weight.detach().float().reshape(64, 9).contiguous()
        self.stem_scale, self.stem_shift = _fold_bn(cnn.bn0_1, cnn.conv0_1.bias)
        self.conv0_2 = _ConvSpec(cnn.conv0_2, cnn.bn0_2)
        self.stages = []
        for stage in range(1, 5):
            units = []
            for unit in getattr(cnn, "block%d" % stage):
                units.append({
                    "conv1": _ConvSpec(unit.conv1, unit.bn1),
                    "conv2": _ConvSpec(unit.conv2, unit.bn2),
                    "se_w1": unit.se.fc[0].weight.detach().float().contiguous(),
                    "se_w2": unit.se.fc[2].weight.detach().float().contiguous(),
                    "shortcut": None if unit.downsample is None else _ConvSpec(unit.downsample[0], unit.downsample[1]),
                })
            tail = _ConvSpec(getattr(cnn, "conv%d" % stage), getattr(cnn, "bn%d" % stage))
            self.stages.append((units, tail))
        lin = model.linear
        n, d = lin.weight.shape
        cf = d // _FEATURE_ROWS
        # reference flatten(1,2) gives d = c*4 + h (:173); our features are NHWC so k = h*512 + c
        self.cls_w = (lin.weight.detach().reshape(n, cf, _FEATURE_ROWS).permute(0, 2, 1)
                      .contiguous().to(torch.bfloat16).reshape(n, d))
        self.cls_b = lin.bias.detach().float().contiguous()
        self.cf = cf


class hctr_model(nn.Module):
    """Drop-in for the reference `hctr_model` (models/handwritten_ctr_model.py:156-178)."""

    def __init__(self, num_classes=7375):
        super().__init__()
        self.img_height = 128
        self.PAD = 'NormalizePAD'
        self.optimizer = 'SGD'
        self.pred = 'CTC'
        self.noutput = num_classes          # 1 blank + characters + 1 unknown
        self.cnn = _Backbone()
        self.linear = nn.Linear(512 * _FEATURE_ROWS, self.noutput)
        # element type of the logits tensor handed back by forward(); the reference returns fp32
        self.logits_dtype = torch.float32
        self._plan = None
        self._plan_key = None
        # optional per-kernel trace: a list that receives (tag, flops, bytes, start_event, end_event) per launch
        self.kernel_trace = None
        self.launch_count = 0

    # -------------------------------------------------------------------------------- plan cache
    def _current_key(self):
        # torch-side changes show up as (data_ptr, _version); kernels that update parameters or running statistics through
        # raw pointers (hctr_sgd_clip_step, hctr_bn_finalize_train) bump `_param_generation` instead
        # (walks the modules' own parameter / buffer dicts - the module tree is fixed after construction - instead of the
        #  recursive generators: this runs on every forward and weighs on batch-1 latency)
        mods = self.__dict__.get("_key_modules")
        if mods is None:
            mods = self.__dict__["_key_modules"] = [m for m in self.modules() if m._parameters or m._buffers]
        key = [self.__dict__.get("_param_generation", 0)]
        for m in mods:
            for t in m._parameters.values():
                if t is not None:
                    key.append(t.data_ptr()); key.append(t._version)
            for t in m._buffers.values():
                if t is not None:
                    key.append(t.data_ptr()); key.append(t._version)
        return tuple(key)

    def _get_plan(self):
        key = self._current_key()
        if self._plan is None or key != self._plan_key:
            self._plan = _InferencePlan(self)
            self._plan_key = key
        return self._plan

    # -------------------------------------------------------------------------------- forward
    def _check_input(self, input):
        if not input.is_cuda:
            raise RuntimeError("hctr_b200: input must be a CUDA tensor on an sm_100a device (no CPU fallback)")
        if input.dim() != 4 or input.shape[1] != 1:
            raise RuntimeError("hctr_b200: expected input [B,1,%d,W], got %s" % (self.img_height, tuple(input.shape)))
        if input.shape[2] != self.img_height:
            # the reference fails in nn.Linear: (H/32)*512 features != 2048
            raise RuntimeError("mat1 and mat2 shapes cannot be multiplied: input height %d gives %d features, "
                               "linear expects %d" % (input.shape[2], (input.shape[2] // 32) * 512,
                                                      self.linear.in_features))
        if self.linear.weight.device != input.device:
            raise RuntimeError("hctr_b200: input is on %s but parameters are on %s" % (input.device, self.linear.weight.device))

    def forward(self, input):
        self._check_input(input)
        with torch.cuda.device(input.device):
            x = input.detach().float().contiguous()
            if self.training:
                return self._forward_train(x)
            return self._forward_eval(x)

    def greedy_decode(self, input, return_argmax=False):
        """eval() forward fused with greedy CTC decoding: what `codec.decode(model(input))` computes in the reference
        (test.py greedy path; models/handwritten_ctr_model.py:171-178 + utils/ctc_codec.py:70-99), with the arg-max taken in the
        classifier's epilogue, so the [W,B,C] logits are never written to memory. Returns (int32 [B,W] label indices, int32 [B]
        lengths) on the device - feed them to `ctc_codec.indices_to_text`. Bit-identical to
        `codec.greedy_indices(model(input))` with the same `logits_dtype` (the epilogue compares the values as they would have
        been stored)."""
        if self.training:
            raise RuntimeError("hctr_b200: greedy_decode() is an eval()-mode path")
        self._check_input(input)
        with torch.cuda.device(input.device):
            return self._forward_eval(input.detach().float().contiguous(), greedy=True, return_argmax=return_argmax)

    # -------------------------------------------------------------------------------- train mode
    def _engine(self):
        eng = self.__dict__.get("_train_engine")
        if eng is None:
            eng = _core().train_engine.TrainEngine(self)
            self.__dict__["_train_engine"] = eng
        return eng

    def _forward_train(self, x):
        """train(): batch-statistics BN (running stats updated), dropout, autograd to the parameters
        (reference: main.py:367,384 with models/handwritten_ctr_model.py in training mode)."""
        eng = self._engine()
        eng.dropout_enabled = bool(getattr(self, "dropout_enabled", True))
        # one seed per step from torch's CPU generator: reproducible under torch.manual_seed
        base_seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        params = [p for _, p in self.named_parameters()]
        fn = _core().train_engine._TrainFunction
        if torch.is_grad_enabled() and any(p.requires_grad for p in params):
            eng.need_backward = True
            out = fn.apply(eng, x, base_seed, *params)
        else:
            eng.need_backward = False
            logits, _ = eng.forward(x, base_seed)
            out = logits[:, :, :self.noutput].permute(1, 0, 2)
        return out if self.logits_dtype == torch.bfloat16 else out.float()

    def _launch(self, nat, tag, flops, nbytes, fn, *args):
        """One C-ABI call = one kernel launch on the current stream; optionally bracketed by CUDA events."""
        self.launch_count += 1
        trace = self.kernel_trace
        if trace is None:
            nat.check(fn(*args), tag)
            return
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        nat.check(fn(*args), tag)
        e1.record()
        trace.append((tag, flops, nbytes, e0, e1))

    def _conv(self, nat, x, spec, B, H, W, relu, pool):
        ho = H // 2 if pool else H
        y = torch.empty((B, ho, W, spec.cout), dtype=torch.bfloat16, device=x.device)
        taps = spec.ksize * spec.ksize
        flops = 2.0 * B * H * W * spec.cout * spec.cin * taps
        nbytes = 2.0 * (x.numel() + y.numel() + spec.w.numel())
        tag = "conv%dx%d_%d_%d_h%d%s" % (spec.ksize, spec.ksize, spec.cin, spec.cout, H, "_pool" if pool else "")
        self._launch(nat, tag, flops, nbytes, nat.lib().hctr_conv_bn_act_fwd,
                     nat.ptr(x), nat.ptr(spec.w), nat.ptr(spec.scale), nat.ptr(spec.shift), nat.ptr(y),
                     B, H, W, spec.cin, spec.cout, spec.ksize, int(relu), int(pool), nat.stream_ptr())
        return y

    # -------------------------------------------------------------------------------- small batches: CUDA-graph replay
    # A single 128x2048 line is ~75 kernel launches of 10-60 us each: issued one ctypes call at a time the host is the
    # slower side (3.3 ms per call for 2.9 ms of device time, most of that gaps). A shape that keeps coming back (batch-1
    # serving, BASELINE configs[0]) is therefore captured ONCE into a CUDA graph - same kernels, same C-ABI calls, made on
    # the capture stream - and replayed: one cudaGraphLaunch per forward. The graph owns a static input buffer and its
    # intermediates (torch's graph-private pool); results are handed back as copies, so callers never see a buffer that a
    # later call overwrites. Captured on the third call with a shape; at most `_GRAPH_SLOTS` shapes are kept (LRU).
    # `model.cuda_graphs = False` or HCTR_CUDA_GRAPHS=0 keeps every call eager; so do kernel traces and outer captures.
    _GRAPH_MAX_COLUMNS = 8192          # B*W up to which a forward is launch-bound enough to be worth a graph
    _GRAPH_SLOTS = 8

    def _forward_eval(self, x, greedy=False, return_argmax=False):
        B, _, _, W = x.shape
        if (B * W > self._GRAPH_MAX_COLUMNS or not getattr(self, "cuda_graphs", True) or self.kernel_trace is not None
                or os.environ.get("HCTR_CUDA_GRAPHS", "1") == "0" or torch.cuda.is_current_stream_capturing()):
            return self._forward_eval_eager(x, greedy, return_argmax)
        plan = self._get_plan()
        cache = self.__dict__.get("_graph_cache")
        if cache is None or cache["plan"] is not plan:
            cache = {"plan": plan, "entries": {}, "tick": 0}        # new weights / folded BN: every captured graph is stale
            self.__dict__["_graph_cache"] = cache
        key = (B, W, bool(greedy), bool(return_argmax), self.logits_dtype, x.device.index, bool(getattr(self, "se_from_input", True)))
        ent = cache["entries"].get(key)
        if ent is None:
            ent = cache["entries"][key] = {"seen": 0, "graph": None}
        cache["tick"] += 1
        ent["used"] = cache["tick"]
        if ent["graph"] is None:
            ent["seen"] += 1
            if ent["seen"] < 3 or ent.get("failed"):
                return self._forward_eval_eager(x, greedy, return_argmax)
            try:
                self._capture_eval_graph(cache, ent, x, greedy, return_argmax)
            except Exception:                                            # capture is an optimisation, never a requirement
                ent["failed"] = True
                ent["graph"] = None
                torch.cuda.synchronize(x.device)
                return self._forward_eval_eager(x, greedy, return_argmax)
        ent["x"].copy_(x)
        ent["graph"].replay()
        self.launch_count += ent["launches"]
        outs = ent["outs"]
        if isinstance(outs, tuple):
            return tuple(self._copy_out(o) for o in outs)
        return self._copy_out(outs)

    @staticmethod
    def _copy_out(o):
        """A private copy of a graph-owned result with the shape AND strides of the eager result (the logits are a permuted
        view of a padded [B,W,pitch] buffer, like the reference's `.permute(1, 0, 2)`)."""
        if o._base is None:
            return o.clone()
        return o._base.clone().as_strided(o.shape, o.stride(), o.storage_offset())

    def _capture_eval_graph(self, cache, ent, x, greedy, return_argmax):
        live = [k for k, e in cache["entries"].items() if e.get("graph") is not None]
        if len(live) >= self._GRAPH_SLOTS:                               # least recently used shape makes room
            victim = min(live, key=lambda k: cache["entries"][k]["used"])
            del cache["entries"][victim]
        ent["x"] = x.clone()
        g = torch.cuda.CUDAGraph()
        before = self.launch_count
        with torch.cuda.graph(g):
            ent["outs"] = self._forward_eval_eager(ent["x"], greedy, return_argmax)
        ent["launches"] = self.launch_count - before
        self.launch_count = before
        ent["graph"] = g

    def _forward_eval_eager(self, x, greedy=False, return_argmax=False):
        nat = _core().native
        lib = nat.lib()
        plan = self._get_plan()
        B, _, H, W = x.shape
        st = nat.stream_ptr()
        dev = x.device

        a = torch.empty((B, H, W, 64), dtype=torch.bfloat16, device=dev)
        self._launch(nat, "stem", 2.0 * B * H * W * 64 * 9, 4.0 * x.numel() + 2.0 * a.numel(), lib.hctr_stem_conv_fwd,
                     nat.ptr(x), nat.ptr(plan.stem_w), nat.ptr(plan.stem_scale), nat.ptr(plan.stem_shift), nat.ptr(a),
                     B, H, W, 1, st)
        a = self._conv(nat, a, plan.conv0_2, B, H, W, relu=True, pool=True)
        H //= 2
        for units, tail in plan.stages:
            for u in units:
                a = self._residual_unit(nat, lib, a, u, B, H, W, st, dev)
            a = self._conv(nat, a, tail, B, H, W, relu=True, pool=True)
            H //= 2
        if greedy:
            return self._classify_greedy(nat, a, plan, B, H, W, return_argmax)
        return self._classify(nat, a, plan, B, H, W)

    def _residual_unit(self, nat, lib, a, u, B, H, W, st, dev):
        """BasicBlock.forward (reference :47-58). Default: the SE gate is derived from conv1's output (its mean over the
        line is linear in it), so conv2's epilogue applies gate, residual and ReLU and bn2's output never exists in memory.
        `self.se_from_input = False` runs the literal sequence conv2 -> squeeze -> excite -> scale/residual/ReLU pass."""
        c1, c2 = u["conv1"], u["conv2"]
        C = c2.cout
        if not getattr(self, "se_from_input", True):
            t = self._conv(nat, a, c1, B, H, W, relu=True, pool=False)
            slices = lib.hctr_conv_se_slices(H, W, C)
            partial = torch.empty((B, slices, C), dtype=torch.float32, device=dev)
            v = torch.empty((B, H, W, C), dtype=torch.bfloat16, device=dev)
            self._launch(nat, "conv3x3_%d_%d_h%d_se" % (c2.cin, c2.cout, H), 2.0 * B * H * W * C * c2.cin * 9,
                         2.0 * (t.numel() + v.numel() + c2.w.numel()), lib.hctr_conv_bn_se_fwd,
                         nat.ptr(t), nat.ptr(c2.w), nat.ptr(c2.scale), nat.ptr(c2.shift), nat.ptr(v), nat.ptr(partial),
                         B, H, W, c2.cin, c2.cout, c2.ksize, st)
            del t
            gate = torch.empty((B, C), dtype=torch.float32, device=dev)
            self._launch(nat, "se_excite", 0.0, 4.0 * partial.numel(), lib.hctr_se_excite,
                         nat.ptr(partial), slices, nat.ptr(u["se_w1"]), nat.ptr(u["se_w2"]), nat.ptr(gate),
                         B, C, u["se_w1"].shape[0], H * W, st)
            res = a if u["shortcut"] is None else self._conv(nat, a, u["shortcut"], B, H, W, relu=False, pool=False)
            out = torch.empty_like(v)
            self._launch(nat, "se_scale_residual_relu", 0.0, 6.0 * v.numel(), lib.hctr_se_scale_residual_relu,
                         nat.ptr(v), nat.ptr(gate), nat.ptr(res), nat.ptr(out), B, H, W, C, st)
            return out
        # conv1 + bn1 + relu, with the per-channel sums of the stored tensor t
        slices = lib.hctr_conv_sum_slices(H, W, c1.cin, c1.cout, c1.ksize)
        partial = torch.empty((B, slices, C), dtype=torch.float32, device=dev)
        t = torch.empty((B, H, W, C), dtype=torch.bfloat16, device=dev)
        self._launch(nat, "conv3x3_%d_%d_h%d_sum" % (c1.cin, c1.cout, H), 2.0 * B * H * W * c1.cout * c1.cin * 9,
                     2.0 * (a.numel() + t.numel() + c1.w.numel()), lib.hctr_conv_bn_act_sum_fwd,
                     nat.ptr(a), nat.ptr(c1.w), nat.ptr(c1.scale), nat.ptr(c1.shift), nat.ptr(t), nat.ptr(partial),
                     B, H, W, c1.cin, c1.cout, c1.ksize, 1, st)
        gate = torch.empty((B, C), dtype=torch.float32, device=dev)
        ws_bytes = lib.hctr_se_gate_workspace_bytes(B, C)
        ws = torch.empty((ws_bytes // 4,), dtype=torch.float32, device=dev)
        self._launch(nat, "se_gate_from_input", 2.0 * B * C * 9 * C, 4.0 * partial.numel() + 2.0 * B * c2.w.numel(),
                     lib.hctr_se_gate_from_input, nat.ptr(t), nat.ptr(partial), slices, nat.ptr(c2.w), nat.ptr(c2.scale),
                     nat.ptr(c2.shift), nat.ptr(u["se_w1"]), nat.ptr(u["se_w2"]), nat.ptr(gate), B, H, W, C,
                     u["se_w1"].shape[0], nat.ptr(ws), ws_bytes, st)
        self.launch_count += 2                      # sums, mean and FC kernels behind one entry point
        res = a if u["shortcut"] is None else self._conv(nat, a, u["shortcut"], B, H, W, relu=False, pool=False)
        out = torch.empty((B, H, W, C), dtype=torch.bfloat16, device=dev)
        self._launch(nat, "conv3x3_%d_%d_h%d_gate_res" % (c2.cin, c2.cout, H), 2.0 * B * H * W * C * c2.cin * 9,
                     2.0 * (t.numel() + res.numel() + out.numel() + c2.w.numel()), lib.hctr_conv_bn_gate_res_fwd,
                     nat.ptr(t), nat.ptr(c2.w), nat.ptr(c2.scale), nat.ptr(c2.shift), nat.ptr(gate), nat.ptr(res), nat.ptr(out),
                     B, H, W, c2.cin, c2.cout, c2.ksize, 1, st)
        return out

    def _logits_layout(self, nat):
        n = self.noutput
        if self.logits_dtype == torch.bfloat16:
            return nat.HCTR_BF16, (n + 7) // 8 * 8               # 16-byte aligned rows
        if self.logits_dtype == torch.float32:
            return nat.HCTR_F32, (n + 3) // 4 * 4
        raise ValueError("logits_dtype must be torch.float32 or torch.bfloat16")

    def _classify_greedy(self, nat, feat, plan, B, Hf, W, return_argmax):
        """Classifier GEMM with the arg-max in its epilogue + collapse; the logits are never written (K7c)."""
        n = self.noutput
        code, pitch = self._logits_layout(nat)
        dev = feat.device
        lib = nat.lib()
        raw = torch.empty((B, W), dtype=torch.int32, device=dev)
        idx = torch.zeros((B, W), dtype=torch.int32, device=dev)
        ln = torch.zeros((B,), dtype=torch.int32, device=dev)
        wsb = lib.hctr_classifier_greedy_workspace_bytes(B, W, n)
        ws = torch.empty((wsb + 16,), dtype=torch.uint8, device=dev)
        off = (-ws.data_ptr()) % 16
        self._launch(nat, "classifier_greedy", 2.0 * B * W * n * Hf * plan.cf, 2.0 * feat.numel() + 2.0 * plan.cls_w.numel() + wsb,
                     lib.hctr_classifier_greedy_fwd, nat.ptr(feat), nat.ptr(plan.cls_w), nat.ptr(plan.cls_b), None, code, pitch,
                     B, Hf, W, plan.cf, n, nat.ptr(raw), nat.ptr(idx), nat.ptr(ln), nat.c_void_p(ws.data_ptr() + off), wsb,
                     nat.stream_ptr())
        self.launch_count += 2                      # + the arg-max fix-up and the collapse kernel
        return (idx, ln, raw) if return_argmax else (idx, ln)

    def _classify(self, nat, feat, plan, B, Hf, W):
        n = self.noutput
        code, pitch = self._logits_layout(nat)
        logits = torch.empty((B, W, pitch), dtype=self.logits_dtype, device=feat.device)
        self._launch(nat, "classifier", 2.0 * B * W * n * Hf * plan.cf,
                     2.0 * feat.numel() + logits.numel() * logits.element_size() + 2.0 * plan.cls_w.numel(),
                     nat.lib().hctr_classifier_fwd, nat.ptr(feat), nat.ptr(plan.cls_w), nat.ptr(plan.cls_b),
                     nat.ptr(logits), code, pitch, B, Hf, W, plan.cf, n, nat.stream_ptr())
        # reference: x.permute(1, 0, 2) of the contiguous [B,W,C] linear output (:176)
        return logits[:, :, :n].permute(1, 0, 2)
