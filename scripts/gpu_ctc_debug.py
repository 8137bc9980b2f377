import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import hctr_b200, oracle, synth
from hctr_b200.ctc_loss import CTCLoss
for (T,B,C,peak,seed) in [(160,2,101,0.0,1),(160,2,101,4.0,2),(30,3,20,4.0,41),(512,2,7375,0.0,3),(2048,2,7375,1.0,4)]:
    x = synth.ctc_like_logits(T,B,C,seed,peak=peak, noise=0.02 if peak==0 else 2.0)
    tg, tl = synth.ctc_targets(B,C,5,12,seed+1)
    xt = torch.from_numpy(x).cuda().requires_grad_(True)
    loss = CTCLoss(zero_infinity=True)(xt, torch.from_numpy(tg), torch.IntTensor([T]*B), torch.from_numpy(tl)); loss.backward()
    ol, _, og = oracle.ctc_loss(x, tg, [T]*B, tl)
    # torch reference on GPU fp32
    xr = torch.from_numpy(x).cuda().requires_grad_(True)
    lr = torch.nn.CTCLoss(zero_infinity=True)(xr.log_softmax(2), torch.from_numpy(tg).cuda(), torch.IntTensor([T]*B).cuda(), torch.from_numpy(tl).cuda()); lr.backward()
    g = xt.grad.cpu().numpy(); gr = xr.grad.cpu().numpy()
    print("T%d C%d peak%g: loss %.6f oracle %.6f torch %.6f | grad err ours %.3e torch-gpu %.3e | max|g| %.3e" % (T,C,peak,loss.item(),ol,lr.item(),np.abs(g-og).max(),np.abs(gr-og).max(),np.abs(og).max()))
