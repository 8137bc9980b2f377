#!/usr/bin/env python
"""bench.py - text-lines/s of the HCTR recognition hot path (BASELINE.json metric) on N B200s.

One "step" = one pass of the hot path over one batch of synthetic input: configs[1] of BASELINE.json,
B=64 synthetic 128x2048 lines per GPU, bf16 logits, greedy CTC decode (forward + decode kernels).
  value     lines/s with the input batch already resident in HBM (CUDA events, max over ranks)
  e2e       the same step through the public API with HOST buffers: pinned fp32 images -> H2D -> hctr_model ->
            ctc_codec.decode -> Python strings (D2H of the compact label arrays inside the timed region)
  roofline  dominant kernel (tcgen05 implicit-GEMM conv on CTA pairs, 256x256 tiles) vs the measured bf16 peak, timed live with
            CUDA events around each of its launches in an instrumented pass of the same step
  cpu_baseline  the oracle port (torch fp32 restatement + C greedy decode) on the host cores, bounded sample
`--impl reference` times that CPU implementation as its own arm (rank 0 only).
Multi-GPU: lines are independent -> batch-sharded replicas, no data-path collective, scaling "weak".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

B_PER_GPU = 64
WIDTH = 2048
HEIGHT = 128
NUM_CLASSES = 7375
FLOPS_PER_COLUMN = 1358901248          # SURVEY.md App. A / BASELINE.md §3: 33 convs + classifier, 2*MAC
METRIC = "text_lines_per_sec_128x2048"
UNIT = "lines/s"
WORKLOAD = "configs[1]: batched HCTR inference B=64 synthetic 128x2048 lines bf16, greedy decode, per B200"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            d = json.load(fh)
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons, power = [], None, set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                clk, mxv = float(parts[0]), float(parts[1])
            except ValueError:
                continue
            mx = mxv
            if t0 - 0.05 <= ts <= t1 + 0.15:
                sm.append(clk)
                try:
                    power.append(float(parts[2]))
                except ValueError:
                    pass
                for n, v in zip(names, parts[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power) if power else None}


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_reference_step(sd, x, tables):
    """The reference's CPU path restated (oracle port): fp32 forward (models/handwritten_ctr_model.py:171-178) +
    ctc_codec.__greedy_search__ (utils/ctc_codec.py:70-99). Returns the decoded strings."""
    import torch
    import oracle
    from oracle import hctr_forward
    with torch.no_grad():
        logits = hctr_forward.forward(x, sd)
    _, idx, ln = oracle.greedy_decode(logits.contiguous().numpy())
    return tables.to_text(idx, ln)


def make_cpu_state(seed=1234):
    from oracle import hctr_forward
    from oracle.codec import CodecTables
    import synth
    sd = hctr_forward.random_state_dict(NUM_CLASSES, seed)        # nothing from the product package on this arm
    return sd, CodecTables(synth.charset(NUM_CLASSES - 2))


def run_cpu_sample(width, lines, repeats=1):
    """Time `repeats` CPU steps over `lines` synthetic lines of 128 x width; returns (lines/s at 2048-column
    equivalents, seconds per step)."""
    import torch
    import synth
    sd, tables = make_cpu_state()
    x = torch.from_numpy(synth.text_lines(lines, width, 99))
    t0 = time.time()
    for _ in range(repeats):
        cpu_reference_step(sd, x, tables)
    dt = (time.time() - t0) / repeats
    return lines * (width / float(WIDTH)) / dt, dt


def impl_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import torch
    import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd, tables = make_cpu_state()
    width, lines = WIDTH, 1
    x = torch.from_numpy(synth.text_lines(lines, width, 99))
    t0 = time.time()
    cpu_reference_step(sd, x, tables)                      # first warm-up step doubles as the sizing probe
    probe = time.time() - t0
    total_steps = args.steps + args.warmup
    if probe * total_steps > 240.0:                        # keep the whole arm within a few minutes
        width = 512
        x = torch.from_numpy(synth.text_lines(lines, width, 99))
    for _ in range(max(args.warmup - 1, 0)):
        cpu_reference_step(sd, x, tables)
    t0 = time.time()
    for _ in range(args.steps):
        cpu_reference_step(sd, x, tables)
    dt = (time.time() - t0) / args.steps
    value = lines * (width / float(WIDTH)) / dt
    sample = "%d line(s) of 128x%d per step (fp32, %d threads), scaled to 2048-column lines" % (lines, width, cores)
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out))
    return 0


# ------------------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the bounded extra legs (configs 3/4/5, CTC / top-k rooflines, ...)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else max(args.warmup, 1)
    if args.impl == "reference":
        return impl_reference(args)

    import torch
    import torch.distributed as dist
    import hctr_b200
    from hctr_b200 import native
    from hctr_b200.models.handwritten_ctr_model import hctr_model
    from hctr_b200.utils.ctc_codec import ctc_codec
    import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU: the hot path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    native.check(native.lib().hctr_device_supported(local), "device")
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", ""):
            os.environ["NCCL_DEBUG"] = "WARN"          # keep stdout to the single JSON line
        dist.init_process_group("nccl", device_id=dev)

    torch.manual_seed(1234)
    model = hctr_model(NUM_CLASSES).to(dev).eval()         # random-init weights of the reference architecture
    model.logits_dtype = torch.bfloat16
    codec = ctc_codec(synth.charset(NUM_CLASSES - 2))
    # per-rank shard: B_PER_GPU independent lines (seeded per rank); a few distinct lines tiled to the batch
    base = synth.text_lines(8, WIDTH, 1000 + rank)
    host = torch.from_numpy(base).repeat(B_PER_GPU // 8, 1, 1, 1).contiguous().pin_memory()
    x_dev = host.to(dev)

    def step_device():
        logits = model(x_dev)                              # the reference's two calls: logits written, then read by the decode
        return codec.greedy_indices(logits)

    def step_device_fused():
        return model.greedy_decode(x_dev)                  # classifier with the arg-max in its epilogue: no logits in memory

    def step_e2e():
        x = host.to(dev, non_blocking=True)
        logits = model(x)
        idx, ln = codec.greedy_indices(logits)
        return codec.indices_to_text(idx, ln)              # D2H of idx/len + host index->char mapping

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    with torch.no_grad():
        # ---------------- device-resident throughput
        for _ in range(args.warmup):
            step_device()
        sampler = ClockSampler(local)
        sampler.start()
        barrier()
        model.launch_count = 0
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        t_wall0 = time.time()
        e0.record()
        for _ in range(args.steps):
            step_device()
        e1.record()
        barrier()
        t_wall1 = time.time()
        clocks = sampler.stop(t_wall0, t_wall1)
        ms_step = max_over_ranks(e0.elapsed_time(e1) / args.steps)
        launches = model.launch_count + 2 * args.steps      # + argmax and collapse kernels of the decode
        value = world * B_PER_GPU / (ms_step * 1e-3)

        # ---------------- the same step with the arg-max fused into the classifier epilogue (hctr_model.greedy_decode)
        for _ in range(2):
            step_device_fused()
        barrier()
        e0.record()
        for _ in range(args.steps):
            step_device_fused()
        e1.record()
        barrier()
        ms_fused = max_over_ranks(e0.elapsed_time(e1) / args.steps)

        # ---------------- end to end through the public API (host buffers in, strings out)
        for _ in range(2):
            texts = step_e2e()
        barrier()
        t0 = time.time()
        e0.record()
        for _ in range(args.steps):
            texts = step_e2e()
        e1.record()
        barrier()
        e2e_ms = max_over_ranks(max(e0.elapsed_time(e1), (time.time() - t0) * 1e3) / args.steps)
        e2e_value = world * B_PER_GPU / (e2e_ms * 1e-3)
        assert len(texts) == B_PER_GPU

        # ---------------- roofline of the dominant kernel, timed live around each of its launches
        model.kernel_trace = []
        for _ in range(min(args.steps, 3)):
            logits = model(x_dev)
        trace_steps = min(args.steps, 3)
        dec = []
        for _ in range(trace_steps):
            a0 = torch.cuda.Event(enable_timing=True); a1 = torch.cuda.Event(enable_timing=True)
            a0.record(); codec.greedy_indices(logits); a1.record()
            dec.append((a0, a1))
        torch.cuda.synchronize()
        trace, model.kernel_trace = model.kernel_trace, None
    peaks = measured_peaks()
    by_tag = {}
    for tag, flops, nbytes, a, b in trace:
        d = by_tag.setdefault(tag, [0.0, 0.0, 0.0, 0])
        d[0] += a.elapsed_time(b); d[1] += flops; d[2] += nbytes; d[3] += 1
    total_ms = sum(v[0] for v in by_tag.values())
    # dominant kernel = the 512->512 3x3 conv launches at H=16 (58.6 % of all FLOPs)
    dom = [(t, v) for t, v in by_tag.items() if t.startswith("conv3x3_512_512_h16") and not t.endswith("_pool")]
    dom_ms = sum(v[0] for _, v in dom); dom_fl = sum(v[1] for _, v in dom); dom_n = sum(v[3] for _, v in dom)
    achieved = dom_fl / (dom_ms * 1e-3) / 1e12
    # DRAM bytes per launch from the ncu --set full captures (profiles/ncu_traffic.json), launch-weighted over the two
    # variants of this layer: conv1 (+channel sums; reads x, writes t) and conv2 (+gate+residual+ReLU; reads t and the residual)
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath) and dom_n:
        with open(tpath) as fh:
            tj = json.load(fh)
        n_res = sum(v[3] for t, v in dom if t.endswith("_gate_res"))
        traffic = (tj["igemm_conv3x3_512_512_B64_H16_W2048_gate_res"]["dram_bytes_per_launch"] * n_res
                   + tj["igemm_conv3x3_512_512_B64_H16_W2048"]["dram_bytes_per_launch"] * (dom_n - n_res)) / dom_n
    dom_bytes = sum(v[2] for _, v in dom)
    roofline = {
        "kernel": "igemm_pair_kernel (tcgen05.mma.cta_group::2, M=256 x N=256 per CTA pair, 3 stages x 128 K, 2 TMEM accumulator stages) on the 512->512 3x3 convs (B=64,H=16,W=2048)",
        "bound": "tensor", "achieved": achieved, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
        "frac": achieved / peaks["bf16_tflops_sustained"], "peak_kind": "bf16_tflops_sustained (%s)" % peaks["source"],
        "frac_of_burst_peak": achieved / peaks["bf16_tflops"],
        "flops_per_launch": dom_fl / max(dom_n, 1), "traffic": traffic, "traffic_unit": "bytes of DRAM read+write per launch (ncu --set full, profiles/ncu_traffic.json)",
        "algorithmic_bytes_per_launch": dom_bytes / max(dom_n, 1),
        "launches_per_step": dom_n // trace_steps, "avg_launch_ms": dom_ms / max(dom_n, 1),
        "share_of_forward": dom_ms / total_ms,
    }
    dec_ms = sum(a.elapsed_time(b) for a, b in dec) / len(dec)
    dec_bytes = 2.0 * WIDTH * B_PER_GPU * NUM_CLASSES + 4.0 * B_PER_GPU * WIDTH
    roofline_decode = {"kernel": "ctc_argmax_kernel<bf16> + ctc_collapse_kernel", "bound": "hbm",
                       "achieved": dec_bytes / (dec_ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                       "frac": dec_bytes / (dec_ms * 1e-3) / 1e9 / peaks["hbm_gbs"], "ms": dec_ms, "traffic": None}
    model_tflops = FLOPS_PER_COLUMN * WIDTH * B_PER_GPU / (ms_step * 1e-3) / 1e12
    breakdown = {t: {"ms_per_step": v[0] / trace_steps, "launches": v[3] // trace_steps,
                     "tflops": (v[1] / (v[0] * 1e-3) / 1e12) if v[1] else None,
                     "gbs": v[2] / (v[0] * 1e-3) / 1e9}
                 for t, v in sorted(by_tag.items(), key=lambda kv: -kv[1][0])}

    # ---------------- bounded extra legs: the other BASELINE configs and north-star rooflines (bench_extras.py)
    extras = {}
    if not args.no_extras:
        import bench_extras as bx
        t_extras = time.time()
        if rank == 0:
            bx._guard(extras, "codec", lambda: bx.codec_legs(native, codec, dev, peaks))
            if isinstance(extras.get("codec"), dict) and "error" not in extras["codec"]:
                extras.update(extras.pop("codec"))
            bx._guard(extras, "roofline_ctc_loss", lambda: bx.ctc_loss_legs(native, dev, peaks))
            bx._guard(extras, "value_bn_calibrated_uniform_input", lambda: bx.bn_calibrated_leg(hctr_model, codec, dev, B_PER_GPU, WIDTH))
            bx._guard(extras, "gpu_library_baseline", lambda: bx.gpu_library_leg(dev, WIDTH))
            bx._guard(extras, "b1_latency_ms", lambda: bx.b1_latency_leg(model, codec, dev))
        # config 3: every rank decodes its share of the ragged lines (no collective on the data path)
        c3 = {}
        bx._guard(c3, "leg", lambda: bx.c3_bucketed_leg(model, codec, dev, rank, world))
        barrier()
        if "error" in c3["leg"]:
            extras["c3_bucketed"] = c3["leg"]
        else:
            secs = max_over_ranks(c3["leg"]["seconds"])
            n_lines = c3["leg"]["lines_this_rank"]
            if world > 1:
                t = torch.tensor([n_lines], dtype=torch.float64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.SUM)
                n_lines = int(t.item())
            extras["c3_bucketed"] = bx.c3_finish(c3["leg"], secs, n_lines, world)
        # config 4: the training step; the one collective of the path runs here when N > 1
        logits = None
        torch.cuda.empty_cache()
        if world > 1:
            per_gpu = max(1, 16 // world)
            bx._guard(extras, "c4_train_step", lambda: bx.c4_train_leg(dev, rank, world, per_gpu))
        else:
            c4 = {}
            bx._guard(c4, "lines_2", lambda: bx.c4_train_leg(dev, rank, world, 2))
            bx._guard(c4, "lines_16", lambda: bx.c4_train_leg(dev, rank, world, 16, steps=3, warmup=2))
            extras["c4_train_step"] = c4
        extras["extras_seconds"] = time.time() - t_extras

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        # bounded sample: one untimed warm-up line (thread pool, oneDNN primitives), then ~10-15 s of CPU work
        run_cpu_sample(WIDTH, 1, repeats=1)
        cval, cdt = run_cpu_sample(WIDTH, 1, repeats=8)
        cpu_baseline = {"value": cval, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": "8 steps of 1 line 128x2048 after 1 warm-up (fp32 torch forward + C greedy decode, %.1f s/step)" % cdt}

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch_per_gpu": B_PER_GPU, "global_batch": B_PER_GPU * world, "height": HEIGHT,
                       "width": WIDTH, "num_classes": NUM_CLASSES, "weights": "random-init (seed 1234), eval mode",
                       "parallelism": "batch-sharded replicas x%d, no collective" % world,
                       "l2": "inputs+activations per step (~200 GB streamed) exceed the 126 MB L2; no explicit flush"},
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": int(host.numel() * 4), "d2h_bytes_per_step": int(B_PER_GPU * WIDTH * 4 + B_PER_GPU * 4)},
            "gpu_launches": launches,
            "model_tflops_per_gpu": model_tflops,
            "frac_of_bf16_peak_whole_step": model_tflops / peaks["bf16_tflops_sustained"],
            "value_fused_argmax_epilogue": {
                "value": world * B_PER_GPU / (ms_fused * 1e-3), "unit": UNIT, "ms_per_step": ms_fused,
                "path": "hctr_model.greedy_decode: arg-max in the classifier epilogue + collapse; the 1.93 GB of bf16 logits are neither "
                        "written nor read (SURVEY 8d: report both). Same transcripts bit for bit; about the same speed - the "
                        "classifier's epilogue is not overlapped with its main loop, so the compares cost what the stores saved"},
            "roofline": roofline, "roofline_decode": roofline_decode, "cpu_baseline": cpu_baseline,
            "clocks": clocks, "kernel_breakdown": breakdown,
        }
        out.update(extras)
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
