"""Three launches of the log-softmax + top-k entry point on the config-5 tensor (argument 'bf16' for bf16 rows): the
target of the ncu captures in scripts/gpu_run_c.sh."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, hctr_b200
from hctr_b200 import native as nat
import bench_extras as bx
lib = nat.lib(); dev = torch.device("cuda", 0)
T, B, C, k = 512, 256, 7375, 10
x = bx.beam_logits_device(T, B, C, 0, dev)
if len(sys.argv) > 1: x = x.to(torch.bfloat16)
code = nat.HCTR_BF16 if len(sys.argv) > 1 else nat.HCTR_F32
ti = torch.empty((T, B, k), dtype=torch.int32, device=dev); tp = torch.empty((T, B, k), dtype=torch.float32, device=dev)
lse = torch.empty((T, B), dtype=torch.float32, device=dev)
for _ in range(3):
    nat.check(lib.hctr_ctc_topk_logsoftmax(nat.ptr(x), code, T, B, C, x.stride(0), x.stride(1), k, nat.ptr(ti), nat.ptr(tp), nat.ptr(lse), nat.stream_ptr()))
torch.cuda.synchronize()
