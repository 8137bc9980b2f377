import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import hctr_b200, synth, oracle
from hctr_b200.ctc_loss import CTCLoss, _CtcFromLogits

def carve(T, B, Sp, Sa):
    off = 0; out = {}
    def take(name, n):
        nonlocal off
        out[name] = off; off = (off + n + 255) // 256 * 256
    take("lse", 4*B*T); take("lpg", 4*B*T*Sp); take("pg", 8*B*T*Sp); take("alpha", 8*B*T*Sa); take("beta", 8*B*T*Sa)
    take("ea", 4*B*T*32); take("eb", 4*B*T*32); take("ll", 8*B); take("pfin", 8*B); take("efin", 4*B); take("flag", 4*B)
    return out

sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_gpu_ctc_loss import _aligned_logits
T, B, C = 1024, 4, 500
tg, tl = synth.ctc_targets(B, C, 40, 60, 21, repeat_frac=0.15)
x = _aligned_logits(T, B, C, tg, tl, peak=45.0, seed=22)
_CtcFromLogits.record_fallback = True
xt = torch.from_numpy(x).cuda().requires_grad_(True)
loss = CTCLoss.from_logits(xt, torch.from_numpy(tg), torch.IntTensor([T]*B), torch.from_numpy(tl))
loss.backward(); torch.cuda.synchronize()
ws = _CtcFromLogits.last_workspace
maxL = int(tl.max()); Sp = (2*maxL+1+3)//4*4
K = 4 if Sp <= 128 else 8 if Sp <= 256 else 16
Sa = 32 * K
o = carve(T, B, Sp, Sa)
def arr(name, dt, n):
    return ws[o[name]:o[name] + n * torch.tensor([], dtype=dt).element_size()].clone().view(dt).cpu().numpy()
print("tl", tl, "Sp", Sp, "flag offset", hctr_b200.native.lib().hctr_ctc_loss_flag_offset(T, B, maxL), o["flag"])
print("flags", arr("flag", torch.int32, B))
ea = arr("ea", torch.int32, B*T*32).reshape(B, T*32); eb = arr("eb", torch.int32, B*T*32).reshape(B, T*32)
print("ea drops", (ea & 1).sum(1), "first", [np.nonzero(r & 1)[0][:5] for r in ea], "eb drops", (eb & 1).sum(1), [np.nonzero(r & 1)[0][:5] for r in eb])
print("ea[0,:12]", ea[0,:12], "ea[0,-5:]", ea[0,-5:])
#print("ll", arr("ll", torch.float64, B), "pfin", arr("pfin", torch.float64, B), "efin", arr("efin", torch.int32, B))
al = arr("alpha", torch.float64, B*T*Sa).reshape(B, T, Sa); be = arr("beta", torch.float64, B*T*Sa).reshape(B, T, Sa)
pg = arr("pg", torch.float64, B*T*Sp).reshape(B, T, Sp)
np.set_printoptions(linewidth=200, precision=3)

def ex(v):
    m, e_ = np.frexp(v); out = (e_ - 1).astype(np.int64); out[v <= 0] = -(1 << 40); return out
NEG = -(1 << 30)
for b_ in range(B):
    S = 2*int(tl[b_])+1
    eaL = ea[b_].reshape(T, 32); ebL = eb[b_].reshape(T, 32)
    bad = 0; shown = 0
    for t in range(T):
        if not ((eaL[t] | ebL[t]) & 1).any(): continue
        s_ = np.arange(S); sm = S - 1 - s_
        ma = ex(al[b_, t, s_]); mb = ex(be[b_, t, sm])
        ok = (ma > NEG) & (mb > NEG)
        xx = ma + (eaL[t][s_ // K] >> 1) + mb + (ebL[t][sm // K] >> 1)
        m = xx[ok].max() if ok.any() else NEG
        tiny = ok & ((ma < -760) | (mb < -760))
        cand = xx[tiny].max() if tiny.any() else NEG
        if m <= NEG or cand + 176 > m:
            bad += 1
            if shown < 2:
                shown += 1
                print("b", b_, "t", t, "m", m, "cand", cand)
                print("  states ok", np.nonzero(ok)[0]); print("  ma", ma[ok]); print("  mb", mb[ok]); print("  x", xx[ok])
                print("  ea lanes", (eaL[t][:S//K+1])); print("  eb lanes", ebL[t][:S//K+1])
    print("seq", b_, "bad rows", bad)
