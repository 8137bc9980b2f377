"""BASELINE config 3: inference batch-sharded across N B200s with bucketed variable widths 256..4096.
4096 synthetic lines, widths 64*randint(4,64) (seed 0), buckets of 256 columns, column budget 131072 per batch,
batches dealt to ranks by greedy LPT; per rank: uint8 lines on the host -> device NormalizePAD -> model -> greedy decode.
torchrun --nproc-per-node N scripts/config3_bench.py [--lines 4096]"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, torch.distributed as dist
import hctr_b200, synth
from hctr_b200.models.handwritten_ctr_model import hctr_model
from hctr_b200.utils.ctc_codec import ctc_codec
from hctr_b200.pipeline import bucket_lines, shard_batches, make_batch

ap = argparse.ArgumentParser(); ap.add_argument("--lines", type=int, default=4096); args = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("NCCL_DEBUG", "WARN"); dist.init_process_group("nccl", device_id=dev)
rs = np.random.RandomState(0)
widths = (64 * rs.randint(4, 65, size=args.lines)).tolist()
batches = bucket_lines(widths, 256, 131072)
mine = shard_batches(batches, world)[rank]
# synthetic uint8 lines: one random strip per width class, sliced (content does not affect timing)
strip = ((synth.text_lines(1, 4096, 7)[0, 0] * 0.5 + 0.5) * 255).round().astype(np.uint8)
images = {i: strip[:, :widths[i]] for bi in mine for i in batches[bi][1]}
torch.manual_seed(1234)
model = hctr_model(7375).to(dev).eval(); model.logits_dtype = torch.bfloat16
codec = ctc_codec(synth.charset(7373))

def run():
    n = 0
    with torch.no_grad():
        for bi in mine:
            wb, idx = batches[bi]
            x = make_batch(images, idx, wb, dev)
            texts = codec.decode(model(x))
            n += len(texts)
    return n

with torch.no_grad():                                  # warm-up on the two largest batches
    for bi in mine[:2]:
        wb, idx = batches[bi]; codec.decode(model(make_batch(images, idx, wb, dev)))
if world > 1: dist.barrier()
torch.cuda.synchronize(); t0 = time.time()
n = run()
torch.cuda.synchronize()
dt = torch.tensor([time.time() - t0], dtype=torch.float64, device=dev)
cnt = torch.tensor([n], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(dt, op=dist.ReduceOp.MAX); dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
if rank == 0:
    useful = float(sum(widths)); padded = float(sum(wb * len(idx) for wb, idx in batches))
    secs = float(dt.item())
    print(json.dumps({"config": "configs[2]: inference batch-sharded across N B200, bucketed variable widths 256-4096",
                      "n_gpus": world, "lines": int(cnt.item()), "batches": len(batches), "seconds": secs,
                      "lines_per_sec": cnt.item() / secs, "useful_columns_per_sec": useful / secs, "padded_columns_per_sec": padded / secs,
                      "padding_overhead": padded / useful - 1.0,
                      "equivalent_2048_lines_per_sec_useful": useful / 2048 / secs,
                      "model_tflops_on_padded_columns": 1358901248 * padded / secs / 1e12}))
if world > 1: dist.destroy_process_group()
