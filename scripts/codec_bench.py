"""Timings of the codec-side kernels at BASELINE sizes: greedy decode (config 2), log-softmax+top-k and prefix beam search
(config 5: T=512, B=256, C=7375, width 10), fused log-softmax + CTC loss fwd/bwd (config 4 tensor: T=2048, B=16)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import hctr_b200, synth, oracle
from hctr_b200 import native as nat
from hctr_b200.utils.ctc_codec import ctc_codec
from hctr_b200.ctc_loss import CTCLoss
lib = nat.lib(); dev = "cuda"; HBM = 6547.8
out = {}

def timeit(fn, n=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

C = 7375
codec = ctc_codec(synth.charset(C - 2))
# ---- config 5: beam search
T, B = 512, 256
x = torch.from_numpy(synth.beam_logits(T, B, C, 0, 8)).to(dev)             # fp32 [T,B,C] = 3.87 GB
for dt in (torch.float32, torch.bfloat16):
    xt = x.to(dt)
    k = 10
    ti = torch.empty((T, B, k), dtype=torch.int32, device=dev); tp = torch.empty((T, B, k), dtype=torch.float32, device=dev)
    lse = torch.empty((T, B), dtype=torch.float32, device=dev)
    code = nat.HCTR_F32 if dt == torch.float32 else nat.HCTR_BF16
    ms = timeit(lambda: nat.check(lib.hctr_ctc_topk_logsoftmax(nat.ptr(xt), code, T, B, C, xt.stride(0), xt.stride(1), k, nat.ptr(ti), nat.ptr(tp), nat.ptr(lse), nat.stream_ptr())))
    nbytes = T * B * C * xt.element_size()
    out["topk_logsoftmax_%s" % str(dt).split(".")[1]] = {"ms": ms, "GBs": nbytes / ms / 1e6, "frac_hbm": nbytes / ms / 1e6 / HBM}
for bonus, tab in ((0.0, None), (5.8, None), (5.8, synth.lm_table(C, 9))):
    codec.set_beam_search(use_tfm_pred=False, lm_panelty=2.0, len_bonus=bonus); codec.lm_table = tab
    ms = timeit(lambda: codec.beam_search_indices(x), n=3, warm=1)
    out["beam_T512_B256_bonus%.1f_%s" % (bonus, "table" if tab is not None else "zero")] = {"ms_total": ms, "sequences_per_s": B / ms * 1e3}
# the same search with a synthetic 5-gram back-off model (~150k n-grams over 3000 characters) scored inside the kernel
from hctr_b200.ngram_lm import NgramLM
t0 = time.time()
lm5 = NgramLM.from_arpa_text(synth.arpa_text(synth.charset(C - 2)[:3000], 5, 77, grams_per_order=40000), codec.dict, C)
codec.set_beam_search(use_tfm_pred=False, lm_panelty=2.0, len_bonus=5.8); codec.ngram = lm5
ms = timeit(lambda: codec.beam_search_indices(x), n=3, warm=1)
out["beam_T512_B256_bonus5.8_5gram"] = {"ms_total": ms, "sequences_per_s": B / ms * 1e3, "n_grams": lm5.n_grams,
                                          "table_build_s": time.time() - t0}
codec.ngram = None
# spot-check 4 sequences against the oracle (pinned to the reference)
codec.set_beam_search(use_tfm_pred=False, lm_panelty=2.0, len_bonus=5.8); codec.lm_table = None
idx, ln = codec.beam_search_indices(x[:, :4].contiguous())
t0 = time.time(); oi, ol, _ = oracle.beam_search(x[:, :4].cpu().numpy(), 10, 10, 2.0, 5.8, None); t_or = time.time() - t0
out["beam_oracle_check"] = {"equal": bool(np.array_equal(ln.cpu().numpy(), ol) and all(np.array_equal(idx[b, :ol[b]].cpu().numpy(), oi[b, :ol[b]]) for b in range(4))),
                            "oracle_c_port_s_per_sequence": t_or / 4}
del x
# ---- config 2: greedy
T, B = 2048, 64
lg = torch.randn(B, T, 7376, device=dev).to(torch.bfloat16)
view = lg[:, :, :C].permute(1, 0, 2)
codec.use_beam_search = False
ms = timeit(lambda: codec.greedy_indices(view))
out["greedy_bf16_T2048_B64"] = {"ms": ms, "GBs": 2.0 * T * B * C / ms / 1e6, "frac_hbm": 2.0 * T * B * C / ms / 1e6 / HBM}
del lg, view
# ---- config 4: CTC loss fwd+bwd on [T=2048, B=16, C]
T, B = 2048, 16
for dt in (torch.bfloat16, torch.float32):
    buf = (torch.randn(B, T, 7376, device=dev) * 2).to(dt)
    lgv = buf[:, :, :C].permute(1, 0, 2).detach().requires_grad_(True)
    tg, tl = synth.ctc_targets(B, C, 20, 60, 0, repeat_frac=0.1)
    tgt, tlt, il = torch.from_numpy(tg), torch.from_numpy(tl), torch.IntTensor([T] * B)
    def run():
        lgv.grad = None
        loss = CTCLoss.from_logits(lgv, tgt, il, tlt); loss.backward()
    ms = timeit(run, n=5, warm=2)
    es = buf.element_size(); nbytes = 3.0 * es * T * B * C
    out["ctc_loss_fwd_bwd_%s_T2048_B16" % str(dt).split(".")[1]] = {"ms": ms, "algorithmic_GBs": nbytes / ms / 1e6, "frac_hbm": nbytes / ms / 1e6 / HBM}
    # torch's own GPU path for the same tensor (library yardstick)
    def run_t():
        a = buf[:, :, :C].permute(1, 0, 2).detach().float().requires_grad_(True)
        l = torch.nn.CTCLoss(zero_infinity=True)(a.log_softmax(2), tgt.to(dev), il.to(dev), tlt.to(dev)); l.backward()
    if dt == torch.float32:
        out["torch_ctc_loss_fp32_T2048_B16_ms"] = timeit(run_t, n=3, warm=1)
print(json.dumps(out, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/codec_bench_r1.json", "w"), indent=1)
