"""One full-size step (B=64, 128x2048, bf16 logits, greedy decode) bracketed by cudaProfilerStart/Stop so that
`ncu --profile-from-start off` sees exactly the 73 kernels of the hot path (B200_PROFILING.md recipe)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import hctr_b200, synth
from hctr_b200.models.handwritten_ctr_model import hctr_model
from hctr_b200.utils.ctc_codec import ctc_codec

B = int(os.environ.get("PROFILE_B", "64")); W = int(os.environ.get("PROFILE_W", "2048"))
dev = torch.device("cuda:0")
torch.manual_seed(1234)
model = hctr_model(7375).to(dev).eval(); model.logits_dtype = torch.bfloat16
codec = ctc_codec(synth.charset(7373))
x = torch.from_numpy(synth.text_lines(8, W, 1000)).repeat(B // 8, 1, 1, 1).contiguous().to(dev)
with torch.no_grad():
    codec.greedy_indices(model(x))            # warm-up (plan build, module load)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    idx, ln = codec.greedy_indices(model(x))
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("ok", int(ln.sum()))
