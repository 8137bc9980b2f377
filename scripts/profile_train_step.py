"""One full training step (config 4 per-GPU shape: 2 lines of 128x2048) bracketed by cudaProfilerStart/Stop."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import hctr_b200, synth
from hctr_b200.models.handwritten_ctr_model import hctr_model
from hctr_b200.train_step import TrainStep
B = int(os.environ.get("PROFILE_B", "2")); W = int(os.environ.get("PROFILE_W", "2048")); C = 7375
dev = torch.device("cuda:0")
torch.manual_seed(1234)
model = hctr_model(C).to(dev).train()
ts = TrainStep(model)
x = torch.from_numpy(synth.text_lines(B, W, 2000)).to(dev)
tg, tl = synth.ctc_targets(B, C, 20, 60, 3000)
for _ in range(3): ts.step(x, tg, tl)
torch.cuda.synchronize()
t0 = time.time()
for _ in range(5): ts.step(x, tg, tl)
cpu_enqueue = (time.time() - t0) / 5
torch.cuda.synchronize()
wall = (time.time() - t0) / 5
torch.cuda.profiler.start()
loss = ts.step(x, tg, tl)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok loss %.4f  cpu enqueue %.2f ms/step  wall %.2f ms/step" % (loss.item(), cpu_enqueue * 1e3, wall * 1e3))
