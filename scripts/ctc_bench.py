"""Time hctr_ctc_loss_fwd_bwd through the C ABI alone (device-resident arguments, CUDA events), at BASELINE config-4
sizes and larger batches. Prints one JSON object; run under ncu for the per-kernel split."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import hctr_b200, synth
from hctr_b200 import native as nat

lib = nat.lib(); dev = "cuda"
HBM = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6547.8
C, T = 7375, 2048
out = {}
cases = [(2, torch.bfloat16), (16, torch.bfloat16), (16, torch.float32), (64, torch.bfloat16)]
if len(sys.argv) > 1:
    cases = [(int(sys.argv[1]), torch.bfloat16)]
for B, dt in cases:
    buf = (torch.randn(B, T, 7376, device=dev) * 2).to(dt)
    grad = torch.empty_like(buf)
    tg, tl = synth.ctc_targets(B, C, 20, 60, 0, repeat_frac=0.1)
    tgt = torch.from_numpy(tg).to(dev); tlt = torch.from_numpy(tl).to(dev); il = torch.full((B,), T, dtype=torch.int32, device=dev)
    maxl = int(tl.max())
    nll = torch.empty(B, device=dev); loss = torch.empty(1, device=dev)
    wsb = lib.hctr_ctc_loss_workspace_bytes(T, B, maxl)
    ws = torch.empty(wsb + 256, dtype=torch.uint8, device=dev); off = (-ws.data_ptr()) % 256; ws = ws[off:off + wsb]
    code = nat.HCTR_F32 if dt == torch.float32 else nat.HCTR_BF16
    def run():
        nat.check(lib.hctr_ctc_loss_fwd_bwd(nat.ptr(buf), code, T, B, C, 7376, T * 7376, nat.ptr(tgt), nat.ptr(tlt), nat.ptr(il),
                                            maxl, None, nat.ptr(nll), nat.ptr(loss), nat.ptr(grad), 1.0, nat.ptr(ws), wsb, nat.stream_ptr()))
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    n = 10
    e0.record()
    for _ in range(n): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    nbytes = 3.0 * buf.element_size() * T * B * C
    foff = lib.hctr_ctc_loss_flag_offset(T, B, maxl)
    flags = ws[foff:foff + 4 * B].clone().view(torch.int32).cpu().numpy()
    out["ctc_loss_fwd_bwd_%s_T%d_B%d" % (str(dt).split(".")[1], T, B)] = {
        "ms": ms, "algorithmic_GBs": nbytes / ms / 1e6, "frac_hbm": nbytes / ms / 1e6 / HBM, "loss": float(loss.item()),
        "log_space_fallbacks": int(flags.sum())}
print(json.dumps(out, indent=1))
