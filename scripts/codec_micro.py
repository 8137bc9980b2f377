"""A/B timings of the two memory-bound codec passes (round 2): log-softmax + top-k (one warp per row vs one CTA per row) and
the CTC loss (overlapped one-pass rows kernel vs the three sequential passes). CUDA events, device-resident arguments."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import hctr_b200, synth
from hctr_b200 import native as nat
import bench_extras as bx

lib = nat.lib(); dev = torch.device("cuda", 0); HBM = 6547.8
out = {}
T, B, C, k = 512, 256, 7375, 10
x = bx.beam_logits_device(T, B, C, 0, dev)
ti = torch.empty((T, B, k), dtype=torch.int32, device=dev); tp = torch.empty((T, B, k), dtype=torch.float32, device=dev)
lse = torch.empty((T, B), dtype=torch.float32, device=dev)
for dt, name in ((torch.float32, "f32"), (torch.bfloat16, "bf16")):
    xt = x if dt == torch.float32 else x.to(dt)
    code = nat.HCTR_F32 if dt == torch.float32 else nat.HCTR_BF16
    res = {}
    keys = ("HCTR_TOPK_WARP", "HCTR_TOPK_THREADS", "HCTR_TOPK_NBUF", "HCTR_TOPK_CHUNK")
    base = None
    for label, env in (("chunk", {"HCTR_TOPK_CHUNK": "1"}), ("warp", {"HCTR_TOPK_CHUNK": "0", "HCTR_TOPK_WARP": "1"}),
                       ("cta256", {"HCTR_TOPK_CHUNK": "0", "HCTR_TOPK_WARP": "0", "HCTR_TOPK_THREADS": "256"}),
                       ("cta128", {"HCTR_TOPK_CHUNK": "0", "HCTR_TOPK_WARP": "0", "HCTR_TOPK_THREADS": "128"}),
                       ("default", {})):
        for k2 in keys:
            os.environ.pop(k2, None)
        os.environ.update(env)
        ti.zero_(); tp.zero_(); lse.zero_()
        ms = bx._timeit(lambda: nat.check(lib.hctr_ctc_topk_logsoftmax(nat.ptr(xt), code, T, B, C, xt.stride(0), xt.stride(1), k,
                                                                       nat.ptr(ti), nat.ptr(tp), nat.ptr(lse), nat.stream_ptr())), 20, 5)
        res[label] = {"ms": ms, "frac_hbm": T * B * C * xt.element_size() / ms / 1e6 / HBM, "chk": int(ti.sum().item())}
        if base is None:
            base = (ti.clone(), tp.clone(), lse.clone())
        else:       # every variant must give the same candidates; log-probs to rounding
            res[label]["idx_equal_to_chunk"] = bool(torch.equal(ti, base[0]))
            res[label]["max_abs_dlogp"] = float((tp - base[1]).abs().max()); res[label]["max_abs_dlse"] = float((lse - base[2]).abs().max())
    for k2 in keys:
        os.environ.pop(k2, None)
    out["topk_" + name] = res
del x
peaks = {"hbm_gbs": HBM}
os.environ.pop("HCTR_TOPK_CTAS", None)
for mode in ("", "4", "2", "1", "0"):
    os.environ["HCTR_CTC_OVERLAP"] = mode
    os.environ.pop("HCTR_CTC_ROWS_CTAS", None)
    if mode.startswith("c"):
        os.environ["HCTR_CTC_ROWS_CTAS"] = mode[1:]
    r = bx.ctc_loss_legs(nat, dev, peaks)
    out[{"": "ctc_default", "4": "ctc_split_schedule", "2": "ctc_rows_scan_fix", "1": "ctc_overlapped_schedule", "0": "ctc_round1_passes"}[mode]] = {k2: {"ms": v["ms"], "frac": v["frac"], "loss": v["loss"]} for k2, v in r.items()}
print(json.dumps(out, indent=1))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "codec_micro.json"), "w"), indent=1)
