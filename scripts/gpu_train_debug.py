import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import hctr_b200, synth
from oracle import hctr_forward
from hctr_b200.models.handwritten_ctr_model import hctr_model
from hctr_b200.ctc_loss import CTCLoss
torch.backends.cuda.matmul.allow_tf32 = False; torch.backends.cudnn.allow_tf32 = False
NC, B, W = 61, int(os.environ.get("DBG_B", "2")), int(os.environ.get("DBG_W", "96"))
torch.manual_seed(11); m = hctr_model(NC).cuda().train(); m.dropout_enabled = False
sd0 = {k: v.detach().clone() for k, v in m.state_dict().items()}
x = torch.from_numpy(synth.text_lines(B, W, 71)).cuda()
tg, tl = synth.ctc_targets(B, NC, 4, 9, 72)
def oracle(autocast):
    sd = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point and "running" not in k else v) for k, v in sd0.items()}
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        logits = hctr_forward.forward(x, sd, train_stats={})
    loss = torch.nn.CTCLoss(zero_infinity=True)(logits.float().log_softmax(2), torch.from_numpy(tg).cuda(), torch.IntTensor([W] * B).cuda(), torch.from_numpy(tl).cuda())
    loss.backward()
    return loss.item(), logits.detach().float(), {k: v.grad for k, v in sd.items() if getattr(v, "grad", None) is not None}
l32, lg32, g32 = oracle(False)
l16, lg16, g16 = oracle(True)
logits = m(x)
loss = CTCLoss(zero_infinity=True)(logits, torch.from_numpy(tg), torch.IntTensor([W] * B), torch.from_numpy(tl)); loss.backward()
print("loss ours %.5f fp32 %.5f autocast-bf16 %.5f" % (loss.item(), l32, l16))
print("logits err: ours max %.4f mean %.4f | torch-autocast max %.4f mean %.4f" % ((logits.detach().float() - lg32).abs().max(), (logits.detach().float() - lg32).abs().mean(), (lg16 - lg32).abs().max(), (lg16 - lg32).abs().mean()))
rows = []
for name, p in m.named_parameters():
    ref = g32[name]; n = ref.norm().item()
    rows.append((name, (p.grad.float() - ref).norm().item() / max(n, 1e-12), (g16[name].float() - ref).norm().item() / max(n, 1e-12), n))
for r in rows[::-1]:
    print("%-40s ours %.4f  autocast %.4f  |ref| %.3e" % r)
