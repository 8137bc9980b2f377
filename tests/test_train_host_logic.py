"""CPU tests of the training host logic: flat layout / bucket plan, and the N>1 gradient exchange over gloo
(world_size 2, two real processes) - the data-parallel path of main.py:222-237 without a GPU."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _named_numels():
    from hctr_b200.models.handwritten_ctr_model import hctr_model
    m = hctr_model(53)
    return [(k, p.numel()) for k, p in m.named_parameters()]


def test_bucket_plan_covers_all_parameters_in_backward_order():
    from hctr_b200.train_step import plan_buckets
    named = _named_numels()
    offsets, total, buckets = plan_buckets(named)
    assert [b[0] for b in buckets] == ["linear", "stage4", "stage3", "stage2", "stage1", "stage0"]
    # buckets tile [0, total) without gaps or overlap, parameters are 16-byte aligned and inside their bucket
    spans = sorted((s, e) for _, s, e in buckets)
    assert spans[0][0] == 0 and spans[-1][1] == total
    for (s0, e0), (s1, e1) in zip(spans, spans[1:]):
        assert e0 == s1
    for name, (start, n) in offsets.items():
        assert start % 4 == 0
        key = "linear" if name.startswith("linear.") else "stage" + ([c for c in name.split(".")[1] if c.isdigit()] or ["0"])[0]
        s, e = next((s, e) for k, s, e in buckets if k == key)
        assert s <= start and start + n <= e, name
    assert sum(n for _, n in named) <= total < sum(n for _, n in named) + 4 * len(named)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hctr_b200.train_step import plan_buckets, allreduce_bucket
    named = [("cnn.conv0_1.weight", 10), ("cnn.bn0_1.weight", 3), ("cnn.block1.0.conv1.weight", 17), ("cnn.conv1.weight", 5),
             ("cnn.block4.0.conv1.weight", 9), ("linear.weight", 21), ("linear.bias", 7)]
    offsets, total, buckets = plan_buckets(named)
    flat = torch.zeros(total)
    for i, (name, (s, n)) in enumerate(offsets.items()):
        flat[s:s + n] = float(rank + 1) * (i + 1)
    works = [allreduce_bucket(flat, s, e, async_op=True) for _, s, e in buckets]
    for w in works:
        w.wait()
    avg = flat / world                                   # the 1/world factor is folded into the optimizer (grad_scale)
    ok = True
    for i, (name, (s, n)) in enumerate(offsets.items()):
        ok &= bool(torch.allclose(avg[s:s + n], torch.full((n,), (i + 1) * (1 + 2) / 2.0)))
    q.put((rank, ok, [b[0] for b in buckets]))
    dist.destroy_process_group()


def test_gradient_exchange_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29650 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res)
    assert res[0][2] == ["linear", "stage4", "stage1", "stage0"]
