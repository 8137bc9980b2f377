"""Bounded extra measurements carried on bench.py's single JSON line (round 2): every BASELINE config and every
north-star roofline, so that the driver-run BENCH / SCALE files hold them and not only builder-kept files.

  c3_bucketed           configs[2]: ragged widths 256-4096, bucketed + batch-sharded (pipeline.recognize_lines), lines/s, padding %
  c5_beam               configs[4]: prefix beam search width 10 over T=512, B=256, C=7375 logits, sequences/s
  roofline_topk         log-softmax + top-10 pass on that tensor (fp32 and bf16) vs measured HBM peak
  roofline_ctc_loss     fused log-softmax + CTC loss fwd/bwd at T=2048, C=7375, B in {2, 16, 64} bf16 (+ B=16 fp32)
  c4_train_step         configs[3]: training step (CTC fwd/bwd, NCCL bucketed all-reduce when N>1, fused clip+SGD), global batch 16
                        at N>1 (16/N lines per GPU), 2 and 16 lines at N=1; exposed communication = with minus without the exchange
  value_bn_calibrated_uniform_input   the headline step with BN-calibrated weights and uniform(-1,1) input (power sensitivity)
  gpu_library_baseline  the reference model restated with torch ops (oracle/hctr_forward.py = cuDNN/cuBLAS) under bf16 autocast,
                        channels_last, + logits D2H + CPU greedy decode, as the reference's own GPU path does (test.py:194)
  b1_latency_ms         one 128x2048 line, and the five config-1 widths, host-synchronised per call

Each leg is wrapped: a failure becomes {"error": ...} and never costs the headline. Inputs exceed the 126 MB L2 or are
rotated over >= 512 MB of distinct buffers (stated per leg)."""
import math
import os
import sys
import time

import numpy as np
import torch

NUM_CLASSES = 7375
L2_ROTATE_BYTES = 512 << 20


def _timeit(fn, n, warm):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def _guard(out, key, fn):
    try:
        out[key] = fn()
    except Exception as exc:            # noqa: BLE001 - a leg must never take the headline down
        out[key] = {"error": "%s: %s" % (type(exc).__name__, str(exc)[:300])}
    torch.cuda.synchronize()


# ------------------------------------------------------------------------------------------------ config 5 + top-k
def beam_logits_device(T, B, C, seed, dev, period=8):
    """SURVEY §8d config-5 tensor generated on the device: 2*randn with a planted path (+12 on a random class at t % period == 0,
    +12 on blank otherwise) - a non-empty greedy path and no ties inside the top 11."""
    g = torch.Generator(device=dev).manual_seed(seed)
    x = torch.randn((T, B, C), generator=g, device=dev) * 2.0
    cls = torch.randint(1, C - 1, (T, B), generator=g, device=dev)
    t = torch.arange(T, device=dev).view(T, 1)
    cls = torch.where((t % period) == 0, cls, torch.zeros_like(cls))
    x.scatter_add_(2, cls.unsqueeze(2), torch.full((T, B, 1), 12.0, device=dev))
    return x


def _ncu_traffic(key):
    """DRAM read+write bytes per launch from the committed ncu --set full captures (profiles/ncu_traffic.json), or None."""
    import json
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "ncu_traffic.json")
    try:
        with open(path) as fh:
            return json.load(fh)[key]["dram_bytes_per_launch"]
    except Exception:                   # noqa: BLE001
        return None


def _ctc_traffic():
    """DRAM bytes of one split-schedule call at B=16 bf16: the four captured kernels summed, or None if one is missing."""
    keys = ("r2_ctc_lse_chunk_bf16_T2048_B16_C7375", "r2_ctc_dense_grad_bf16_T2048_B16_C7375", "r2_ctc_fix_bf16_T2048_B16",
            "r2_ctc_scan_T2048_B16")
    vals = [_ncu_traffic(k) for k in keys]
    return None if any(v is None for v in vals) else float(sum(vals))


def codec_legs(nat, codec, dev, peaks):
    lib = nat.lib()
    T, B, C, k = 512, 256, NUM_CLASSES, 10
    x = beam_logits_device(T, B, C, 0, dev)
    out = {"roofline_topk": {}, "c5_beam": None}
    ti = torch.empty((T, B, k), dtype=torch.int32, device=dev); tp = torch.empty((T, B, k), dtype=torch.float32, device=dev)
    lse = torch.empty((T, B), dtype=torch.float32, device=dev)
    for dt, name in ((torch.float32, "f32"), (torch.bfloat16, "bf16")):
        xt = x if dt == torch.float32 else x.to(dt)
        code = nat.HCTR_F32 if dt == torch.float32 else nat.HCTR_BF16
        ms = _timeit(lambda: nat.check(lib.hctr_ctc_topk_logsoftmax(nat.ptr(xt), code, T, B, C, xt.stride(0), xt.stride(1), k,
                                                                    nat.ptr(ti), nat.ptr(tp), nat.ptr(lse), nat.stream_ptr())), 20, 10)
        nbytes = float(T) * B * C * xt.element_size()
        kern = "ctc_topk_chunk_kernel<%s> (one warp per row, the row read once in register chunks: online log-sum-exp + marking of " \
               "the vectors that reach the running top-k bound)" % ("float" if name == "f32" else "bf16")
        out["roofline_topk"][name] = {"kernel": kern, "bound": "hbm", "ms": ms,
                                      "achieved": nbytes / ms / 1e6, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                      "frac": nbytes / ms / 1e6 / peaks["hbm_gbs"], "algorithmic_bytes": nbytes,
                                      "traffic": _ncu_traffic("r2_ctc_topk_chunk_f32_T512_B256_C7375" if name == "f32"
                                                              else "r2_ctc_topk_chunk_bf16_T512_B256_C7375"),
                                      "peak_note": "peak = the copy-measured HBM figure; a read-only stream can exceed it",
                                      "l2": "%.2f GB tensor >> L2" % (nbytes / 1e9)}
    codec.set_beam_search(use_tfm_pred=False, lm_panelty=2.0, len_bonus=5.8, beam_size=10, search_depth=10)
    codec.lm_table = None
    ms = _timeit(lambda: codec.beam_search_indices(x), 5, 2)
    ms0 = None
    codec.set_beam_search(use_tfm_pred=False, lm_panelty=2.0, len_bonus=0.0, beam_size=10, search_depth=10)
    ms0 = _timeit(lambda: codec.beam_search_indices(x), 3, 1)
    codec.use_beam_search = False
    out["c5_beam"] = {"workload": "configs[4]: CTC prefix beam search width 10, T=512, B=256, C=7375, fp32 logits resident in HBM, zero LM",
                      "ms": ms, "sequences_per_s": B / ms * 1e3, "len_bonus": 5.8, "ms_len_bonus_0": ms0,
                      "includes": "log-softmax + top-10 pass and the search kernel"}
    del x
    return out


# ------------------------------------------------------------------------------------------------ CTC loss roofline
def ctc_loss_legs(nat, dev, peaks):
    import synth
    lib = nat.lib()
    C, T, pitch = NUM_CLASSES, 2048, 7376
    res = {}
    for B, dt, name in ((16, torch.bfloat16, "B16_bf16"), (2, torch.bfloat16, "B2_bf16"), (64, torch.bfloat16, "B64_bf16"),
                        (16, torch.float32, "B16_f32")):
        es = 2 if dt == torch.bfloat16 else 4
        per = 2.0 * B * T * pitch * es                                   # logits + gradient buffers of one set
        nrot = max(1, int(math.ceil(L2_ROTATE_BYTES / per)))
        bufs = [(torch.randn(B, T, pitch, device=dev) * 2).to(dt) for _ in range(nrot)]
        grads = [torch.empty_like(b) for b in bufs]
        tg, tl = synth.ctc_targets(B, C, 20, 60, 0, repeat_frac=0.1)
        tgt = torch.from_numpy(tg).to(dev); tlt = torch.from_numpy(tl).to(dev)
        il = torch.full((B,), T, dtype=torch.int32, device=dev)
        maxl = int(tl.max())
        nll = torch.empty(B, device=dev); loss = torch.empty(1, device=dev)
        wsb = lib.hctr_ctc_loss_workspace_bytes(T, B, maxl)
        ws = torch.empty(wsb + 256, dtype=torch.uint8, device=dev); off = (-ws.data_ptr()) % 256
        code = nat.HCTR_F32 if dt == torch.float32 else nat.HCTR_BF16
        it = [0]

        def run():
            i = it[0] % nrot
            it[0] += 1
            nat.check(lib.hctr_ctc_loss_fwd_bwd(nat.ptr(bufs[i]), code, T, B, C, pitch, T * pitch, nat.ptr(tgt), nat.ptr(tlt),
                                                nat.ptr(il), maxl, None, nat.ptr(nll), nat.ptr(loss), nat.ptr(grads[i]), 1.0,
                                                nat.c_void_p(ws.data_ptr() + off), wsb, nat.stream_ptr()))
        ms = _timeit(run, 20, 10)                                        # (short kernels: warm the clocks up first)
        ms_rows = None                                                   # the other schedule, for comparison
        os.environ["HCTR_CTC_OVERLAP"] = "2" if B <= 24 else "4"
        try:
            ms_rows = _timeit(run, 20, 10)
        finally:
            os.environ.pop("HCTR_CTC_OVERLAP", None)
        # the training step's configuration (TrainStep / train_engine): the rows' log-sum-exp comes from the classifier epilogue
        # (hctr_classifier_lse_fwd), so the loss call only gathers the label logits, scans, and writes the gradient
        lses = [torch.logsumexp(b_[:, :, :C].float(), dim=2).contiguous() for b_ in bufs]

        def run_lse():
            i = it[0] % nrot
            it[0] += 1
            nat.check(lib.hctr_ctc_loss_fwd_bwd(nat.ptr(bufs[i]), code, T, B, C, pitch, T * pitch, nat.ptr(tgt), nat.ptr(tlt),
                                                nat.ptr(il), maxl, nat.ptr(lses[i]), nat.ptr(nll), nat.ptr(loss), nat.ptr(grads[i]), 1.0,
                                                nat.c_void_p(ws.data_ptr() + off), wsb, nat.stream_ptr()))
        ms_lse = _timeit(run_lse, 20, 10)
        loss_lse = float(loss.item())
        run()
        split = B <= 24                                                  # csrc/ctc_loss.cu split_by_default
        alg = 3.0 * es * T * B * C                                       # (2*s_in + s_out) * T*B*C
        moved = 2.0 * es * T * B * C                                     # the one-pass rows kernel: logits read once, gradient written once
        foff = lib.hctr_ctc_loss_flag_offset(T, B, maxl)
        flags = ws[off + foff: off + foff + 4 * B].clone().view(torch.int32).cpu().numpy()
        res[name] = {"kernel": ("split schedule, relative form: ctc_prep + ctc_gather_rel_kernel (label logits relative to the row's largest "
                                "label logit: no log-sum-exp needed) -> [ctc_scan_kernel x2 on SMs of their own || ctc_lse_chunk_kernel "
                                "(one read: log-sum-exp) + ctc_dense_grad_kernel on a helper stream] -> verify + nll correction + "
                                "ctc_fix_kernel") if split else
                               "ctc_prep + ctc_rows_kernel (one pass, row in registers) + ctc_scan_kernel x2 + verify + ctc_fix_kernel",
                     "bound": "hbm", "ms": ms,
                     "ms_other_schedule": ms_rows, "other_schedule": "one-pass rows kernel, back to back" if split else "split",
                     "achieved": alg / ms / 1e6, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": alg / ms / 1e6 / peaks["hbm_gbs"],
                     "algorithmic_bytes": alg, "bytes_moved_by_design": alg if split else moved,
                     "traffic": _ctc_traffic() if name == "B16_bf16" else None,
                     "ms_with_classifier_lse": ms_lse, "frac_with_classifier_lse": moved / ms_lse / 1e6 / peaks["hbm_gbs"],
                     "with_classifier_lse": "the training step's call: row log-sum-exp taken from the classifier epilogue, frac on "
                                            "(s_in + s_out)*T*B*C bytes (logits read once by the gradient pass, gradient written once)",
                     "loss_with_classifier_lse": loss_lse,
                     "loss": float(loss.item()), "log_space_fallbacks": int(flags.sum()),
                     "l2": "rotating %d buffer set(s) of %.0f MB" % (nrot, per / 1e6)}
        del bufs, grads, ws, lses
    return res


# ------------------------------------------------------------------------------------------------ config 3
def c3_bucketed_leg(model, codec, dev, rank, world, lines=1024):
    import synth
    from hctr_b200.pipeline import bucket_lines, shard_batches, recognize_lines
    rs = np.random.RandomState(0)
    widths = (64 * rs.randint(4, 65, size=lines)).tolist()
    strip = ((synth.text_lines(1, 4096, 7)[0, 0] * 0.5 + 0.5) * 255).round().astype(np.uint8)
    images = [strip[:, :w] for w in widths]
    batches = bucket_lines(widths, 256, 131072)
    # warm-up on this rank's two largest batches
    mine = shard_batches(batches, world)[rank]
    warm = [i for bi in mine[:2] for i in batches[bi][1]]
    if warm:
        recognize_lines(model, codec, [images[i] for i in warm], 0, 1, device=dev)
    torch.cuda.synchronize()
    t0 = time.time()
    texts = recognize_lines(model, codec, images, rank, world, device=dev)
    torch.cuda.synchronize()
    secs = time.time() - t0
    useful = float(sum(widths)); padded = float(sum(wb * len(idx) for wb, idx in batches))
    return {"seconds": secs, "lines_this_rank": len(texts), "lines": lines, "batches": len(batches),
            "useful_columns": useful, "padded_columns": padded}


def c3_finish(leg, secs_max, total_lines, world):
    useful, padded = leg["useful_columns"], leg["padded_columns"]
    return {"workload": "configs[2]: %d synthetic lines, widths 64*randint(4,64), buckets of 256 columns, <=131072 columns per batch, "
                        "batches dealt to ranks by LPT; host uint8 lines -> device NormalizePAD -> model -> greedy decode -> strings" % leg["lines"],
            "n_gpus": world, "seconds": secs_max, "lines_per_s": total_lines / secs_max, "padding_overhead": padded / useful - 1.0,
            "equivalent_2048_lines_per_s_useful": useful / 2048.0 / secs_max,
            "model_tflops_on_padded_columns_per_gpu": 1358901248.0 * padded / secs_max / 1e12 / world, "collective": "none"}


# ------------------------------------------------------------------------------------------------ config 4
def c4_train_leg(dev, rank, world, lines_per_gpu, steps=5, warmup=3, measure_no_exchange=True):
    import synth
    import torch.distributed as dist
    from hctr_b200.models.handwritten_ctr_model import hctr_model
    from hctr_b200.train_step import TrainStep
    W, C = 2048, NUM_CLASSES
    x = torch.from_numpy(synth.text_lines(lines_per_gpu, W, 2000 + rank)).to(dev)
    tg, tl = synth.ctc_targets(lines_per_gpu, C, 20, 60, 3000 + rank, repeat_frac=0.1)

    def measure(group):
        torch.manual_seed(1234)
        model = hctr_model(C).to(dev).train()
        ts = TrainStep(model, lr=1e-3, momentum=0.9, weight_decay=1e-4, max_norm=5.0, process_group=group)
        for _ in range(warmup):
            loss = ts.step(x, tg, tl)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            loss = ts.step(x, tg, tl)
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        info = {"ms_per_step": ms, "loss": float(loss.item()), "grad_norm": float(ts.norm[0].item()), "world_in_step": ts.world,
                "allreduce_bytes_per_step": int(ts.flat_grads.numel() * 4) if ts.world > 1 else 0}
        del ts, model
        torch.cuda.empty_cache()
        return info

    full = measure(None)
    out = {"lines_per_gpu": lines_per_gpu, "global_batch": lines_per_gpu * world, "width": W, "ms_per_step": full["ms_per_step"],
           "lines_per_s": lines_per_gpu * world / (full["ms_per_step"] * 1e-3),
           "model_tflops_per_gpu": 3 * 1358901248.0 * W * lines_per_gpu / (full["ms_per_step"] * 1e-3) / 1e12,
           "loss": full["loss"], "grad_norm": full["grad_norm"], "allreduce_bytes_per_step": full["allreduce_bytes_per_step"],
           "collective": "NCCL all-reduce (sum) of the flat fp32 gradient in 6 buckets, overlapped with the backward" if world > 1 else "none (1 GPU)"}
    if world > 1 and measure_no_exchange:
        solo = [dist.new_group(ranks=[r]) for r in range(world)]            # every rank creates every group
        alone = measure(solo[rank])
        out["ms_per_step_without_exchange"] = alone["ms_per_step"]
        out["exposed_comm_ms"] = full["ms_per_step"] - alone["ms_per_step"]
    return out


# ------------------------------------------------------------------------------------------------ headline variants
def bn_calibrated_leg(model_cls, codec, dev, b_per_gpu, width, steps=5):
    """The headline step with BN-calibrated weights (one train-mode pass with momentum 1.0, dropout off: running statistics :=
    batch statistics) and uniform(-1,1) input: activations no longer shrink layer by layer, so operand toggling
    - and power - is that of a trained network. Same kernels, same shapes."""
    torch.manual_seed(1234)
    model = model_cls(NUM_CLASSES).to(dev)
    for mod in model.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.momentum = 1.0
    model.train(); model.dropout_enabled = False
    g = torch.Generator(device=dev).manual_seed(5)
    with torch.no_grad():
        model(torch.rand((3, 1, 128, 1024), generator=g, device=dev) * 2 - 1)
    model.eval(); model.logits_dtype = torch.bfloat16
    x = torch.rand((b_per_gpu, 1, 128, width), generator=g, device=dev) * 2 - 1

    def step():
        return codec.greedy_indices(model(x))
    with torch.no_grad():
        ms = _timeit(step, steps, 3)
        logits = model(x)
        absmax = float(logits.float().abs().max().item())
        distinct = int(logits.argmax(2).unique().numel())
    del model, x, logits
    torch.cuda.empty_cache()
    return {"value": b_per_gpu / (ms * 1e-3), "unit": "lines/s", "ms_per_step": ms, "logits_absmax": absmax,
            "distinct_argmax_classes": distinct, "weights": "random-init (seed 1234) + BN running statistics from one train-mode pass",
            "input": "uniform(-1,1)"}


def gpu_library_leg(dev, width, lines=16, steps=3):
    """the reference's own GPU path as the library yardstick - the model restated with torch ops (cuDNN convolutions,
    cuBLAS linear) under torch.autocast(bfloat16), channels_last, then `preds.cpu().numpy()` + the CPU greedy decode exactly as
    test.py:194 / main.py:495 do. This leg is a BASELINE measured beside the product (oracle/ is the checker's restatement)."""
    import oracle
    from oracle import hctr_forward
    sd = {k: v.to(dev) for k, v in hctr_forward.random_state_dict(NUM_CLASSES, 1234).items()}
    for k in list(sd):
        if sd[k].dim() == 4:
            sd[k] = sd[k].contiguous(memory_format=torch.channels_last)
    g = torch.Generator(device=dev).manual_seed(6)
    x = (torch.rand((lines, 1, 128, width), generator=g, device=dev) * 2 - 1).contiguous(memory_format=torch.channels_last)

    def fwd():
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            return hctr_forward.forward(x, sd)
    fwd_ms = _timeit(fwd, steps, 2)
    torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(steps):
        y = fwd()
        host = y.float().cpu().numpy()
        oracle.greedy_decode(np.ascontiguousarray(host))
    e2e_ms = (time.time() - t0) / steps * 1e3
    del sd, x, y
    torch.cuda.empty_cache()
    return {"lines_per_s_forward_only": lines / (fwd_ms * 1e-3), "lines_per_s_with_host_decode": lines / (e2e_ms * 1e-3),
            "forward_ms": fwd_ms, "e2e_ms": e2e_ms, "batch": lines, "width": width,
            "model_tflops_forward": 1358901248.0 * width * lines / (fwd_ms * 1e-3) / 1e12,
            "what": "torch ops (cuDNN/cuBLAS) under autocast(bfloat16), channels_last; decode = logits D2H + CPU arg-max/collapse (reference test.py:194)"}


def b1_latency_leg(model, codec, dev):
    """Latency of a single line through the public API (device-resident fp32 input -> strings), synchronised per call:
    BASELINE configs[0] shape (batch 1, the five bundled widths) and one 128x2048 line. Measured twice: every call issued
    kernel by kernel (`*_eager`), and with the shapes replayed from CUDA graphs (the default for small batches: captured on
    the third call with a shape, hctr_model._forward_eval)."""
    import synth
    out = {}
    prev = getattr(model, "cuda_graphs", True)
    with torch.no_grad():
        for name, widths in (("line_128x2048", [2048]), ("config1_widths_3514_908_2375_1913_488", [3514, 908, 2375, 1913, 488])):
            xs = [torch.from_numpy(synth.text_lines(1, w, 40 + i)).to(dev) for i, w in enumerate(widths)]
            out[name] = {"calls": len(widths)}
            for graphs, sfx in ((False, "_eager"), (True, "")):
                model.cuda_graphs = graphs
                for _ in range(4):                                  # warm-up: plan, shapes, allocator, graph capture
                    for x in xs:
                        codec.decode(model(x))
                torch.cuda.synchronize()
                samples = []
                for _ in range(10):
                    t0 = time.perf_counter()
                    for x in xs:
                        codec.decode(model(x))
                    samples.append((time.perf_counter() - t0) * 1e3)
                samples.sort()
                out[name]["median_ms" + sfx] = samples[len(samples) // 2]
                out[name]["min_ms" + sfx] = samples[0]
                # device time alone (events around the same calls, no host sync in between)
                out[name]["device_ms" + sfx] = _timeit(lambda: [codec.greedy_indices(model(x)) for x in xs], 10, 2)
    model.cuda_graphs = prev
    return out
