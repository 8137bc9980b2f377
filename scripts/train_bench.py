"""BASELINE config 4: HCTR training step, bf16, CTC loss fwd/bwd, NCCL gradient all-reduce.
torchrun --nproc-per-node N scripts/train_bench.py [--lines-per-gpu 2] [--width 2048]   (global batch 16 at N=8, L=2)"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, torch.distributed as dist
import hctr_b200, synth
from hctr_b200.models.handwritten_ctr_model import hctr_model
from hctr_b200.train_step import TrainStep

ap = argparse.ArgumentParser()
ap.add_argument("--lines-per-gpu", type=int, default=2)
ap.add_argument("--width", type=int, default=2048)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--warmup", type=int, default=3)
args = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    dist.init_process_group("nccl", device_id=dev)
B, W, C = args.lines_per_gpu, args.width, 7375
torch.manual_seed(1234)
model = hctr_model(C).to(dev).train()
ts = TrainStep(model, lr=1e-3, momentum=0.9, weight_decay=1e-4, max_norm=5.0)
x = torch.from_numpy(synth.text_lines(B, W, 2000 + rank)).to(dev)
tg, tl = synth.ctc_targets(B, C, 20, 60, 3000 + rank, repeat_frac=0.1)
for _ in range(args.warmup):
    loss = ts.step(x, tg, tl)
if world > 1: dist.barrier()
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
t0 = time.time(); e0.record()
for _ in range(args.steps):
    loss = ts.step(x, tg, tl)
e1.record()
if world > 1: dist.barrier()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / args.steps
wall = (time.time() - t0) / args.steps * 1e3
t = torch.tensor([ms], device=dev, dtype=torch.float64)
if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
ms = float(t.item())
flops = 3 * 1358901248 * W * B * world            # fwd + dgrad + wgrad
if rank == 0:
    print(json.dumps({"config": "configs[3]: HCTR training step bf16, CTC loss fwd/bwd, NCCL gradient all-reduce",
                      "n_gpus": world, "lines_per_gpu": B, "global_batch": B * world, "width": W, "ms_per_step": ms, "wall_ms_per_step": wall,
                      "lines_per_sec": B * world / (ms * 1e-3), "model_tflops_total": flops / (ms * 1e-3) / 1e12,
                      "loss": float(loss.item()), "grad_norm": float(ts.norm[0].item()),
                      "allreduce_bytes_per_step": int(ts.flat_grads.numel() * 4), "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}))
if world > 1: dist.destroy_process_group()
