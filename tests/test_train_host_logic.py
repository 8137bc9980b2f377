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
    # DDP's construction-time sync (main.py:237): unseeded replicas start from rank 0's parameters and buffers
    from hctr_b200.train_step import broadcast_state
    from hctr_b200.models.handwritten_ctr_model import hctr_model
    torch.manual_seed(100 + rank)                       # the reference's --seed defaults to None: ranks differ
    m = hctr_model(11)
    for b in m.buffers():
        if b.dtype.is_floating_point:
            b.add_(float(rank))
    params = torch.cat([p.detach().reshape(-1) for p in m.parameters()])
    n = broadcast_state([params] + list(m.buffers()), src=0)
    torch.manual_seed(100)
    m0 = hctr_model(11)
    want = torch.cat([p.detach().reshape(-1) for p in m0.parameters()])
    ok &= n == 1 + len(list(m.buffers())) and bool(torch.equal(params, want))
    ok &= all(bool(torch.equal(a, b)) for a, b in zip(m.buffers(), m0.buffers()))
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


# ---------------------------------------------------------------------------------------------------------------------
# optimizer state / LR schedule / checkpoint files in the reference's format (main.py:251-265, :343-349, :540-555, :579-584)

def _small_named():
    torch.manual_seed(3)
    named = [("cnn.conv0_1.weight", torch.nn.Parameter(torch.randn(4, 1, 3, 3))), ("cnn.bn0_1.weight", torch.nn.Parameter(torch.randn(4))),
             ("cnn.block1.0.conv1.weight", torch.nn.Parameter(torch.randn(5, 4, 3, 3))), ("linear.weight", torch.nn.Parameter(torch.randn(7, 6))),
             ("linear.bias", torch.nn.Parameter(torch.randn(7)))]
    return named


def test_sgd_state_dict_round_trips_with_torch_sgd():
    from hctr_b200.train_step import plan_buckets, sgd_state_dict, load_sgd_state_dict
    named = _small_named()
    names = [k for k, _ in named]
    shapes = {k: tuple(p.shape) for k, p in named}
    offsets, total, _ = plan_buckets([(k, p.numel()) for k, p in named])
    opt = torch.optim.SGD([p for _, p in named], lr=0.05, momentum=0.9, weight_decay=1e-4)
    # before the first step torch holds no momentum buffers: the flat side must report "first step" too
    flat = torch.full((total,), 7.0)
    group = {"lr": 1.0, "momentum": 0.0, "dampening": 0, "weight_decay": 0.0, "nesterov": False, "params": None}
    assert load_sgd_state_dict(opt.state_dict(), names, offsets, shapes, flat, group) is False
    assert float(flat.abs().sum()) == 0.0 and group["lr"] == 0.05 and group["momentum"] == 0.9 and group["weight_decay"] == 1e-4
    for _, p in named:
        p.grad = torch.randn_like(p)
    opt.step()
    ref_sd = opt.state_dict()
    assert load_sgd_state_dict(ref_sd, names, offsets, shapes, flat, group) is True
    for i, k in enumerate(names):
        s, n = offsets[k]
        assert torch.equal(flat[s:s + n].view(shapes[k]), ref_sd["state"][i]["momentum_buffer"])
    ours = sgd_state_dict(names, offsets, shapes, flat, group, True)
    # what we emit loads into a fresh torch.optim.SGD over the same parameters, and equals torch's own state
    opt2 = torch.optim.SGD([p for _, p in named], lr=1.0)
    opt2.load_state_dict(ours)
    sd2 = opt2.state_dict()
    assert sd2["param_groups"][0]["lr"] == 0.05 and sd2["param_groups"][0]["momentum"] == 0.9
    assert sd2["param_groups"][0]["params"] == ref_sd["param_groups"][0]["params"]
    for i in range(len(names)):
        assert torch.equal(sd2["state"][i]["momentum_buffer"], ref_sd["state"][i]["momentum_buffer"])
    assert sgd_state_dict(names, offsets, shapes, flat, group, False)["state"] == {}


def test_load_sgd_state_dict_rejects_foreign_layouts():
    from hctr_b200.train_step import plan_buckets, load_sgd_state_dict
    named = _small_named()
    names = [k for k, _ in named]
    shapes = {k: tuple(p.shape) for k, p in named}
    offsets, total, _ = plan_buckets([(k, p.numel()) for k, p in named])
    flat = torch.zeros(total)
    group = {"lr": 1.0, "momentum": 0.0, "weight_decay": 0.0}
    opt = torch.optim.SGD([p for _, p in named[:-1]], lr=0.1, momentum=0.9)
    with pytest.raises(ValueError):
        load_sgd_state_dict(opt.state_dict(), names, offsets, shapes, flat, group)
    opt = torch.optim.SGD([p for _, p in named], lr=0.1, momentum=0.9, nesterov=True)
    with pytest.raises(ValueError):
        load_sgd_state_dict(opt.state_dict(), names, offsets, shapes, flat, group)
    opt = torch.optim.SGD([p for _, p in reversed(named)], lr=0.1, momentum=0.9)
    for _, p in named:
        p.grad = torch.ones_like(p)
    opt.step()
    with pytest.raises(ValueError):                      # same count, other order: shapes give it away
        load_sgd_state_dict(opt.state_dict(), names, offsets, shapes, flat, group)


def test_adjust_learning_rate_is_the_reference_schedule():
    from hctr_b200.train_step import adjust_learning_rate

    class A(object):
        lr = 0.1
    named = _small_named()
    opt = torch.optim.SGD([p for _, p in named], lr=0.1)
    for epoch, want in [(0, 0.1), (29, 0.1), (30, 0.01), (59, 0.01), (60, 0.001), (95, 0.0001)]:
        adjust_learning_rate(opt, epoch, A)
        assert opt.param_groups[0]["lr"] == pytest.approx(want, rel=1e-12)


def test_checkpoint_files_follow_the_reference_format(tmp_path):
    from hctr_b200 import checkpoint as ck
    from hctr_b200.models.handwritten_ctr_model import hctr_model

    class Args(object):
        model_type = "hctr"
        multiprocessing_distributed = False
        rank = 0
    torch.manual_seed(11)
    m = hctr_model(37)
    opt = torch.optim.SGD(m.parameters(), lr=0.1, momentum=0.9, weight_decay=1e-4)      # what the reference stores
    for p in m.parameters():
        p.grad = torch.randn_like(p) * 1e-3
    opt.step()
    state = ck.make_state(4, m, 0.875, opt)
    assert sorted(state.keys()) == ["best_acc", "epoch", "optimizer", "state_dict"] and state["epoch"] == 5
    written = ck.save_checkpoint(state, Args, is_best=True, directory=str(tmp_path))
    assert [os.path.basename(w) for w in written] == ["hctr_checkpoint.pth.tar", "hctr_05ep_0.8750acc_checkpoint.pth.tar"]
    written_val = ck.save_checkpoint(state, Args, is_best=False, is_val=True, directory=str(tmp_path))
    assert [os.path.basename(w) for w in written_val] == ["hctr_val_checkpoint.pth.tar"]

    class Rank1(Args):
        multiprocessing_distributed = True
        rank = 1
    assert ck.save_checkpoint(state, Rank1, is_best=True, directory=str(tmp_path / "none")) == []

    # the reference's own readers: torch.load(...)['state_dict'] into the module, ['optimizer'] into torch SGD
    raw = torch.load(written[0], map_location="cpu", weights_only=False)
    assert len(raw["state_dict"]) == 254
    torch.manual_seed(12)
    m2 = hctr_model(37)
    opt2 = torch.optim.SGD(m2.parameters(), lr=1.0)
    epoch, best = ck.load_checkpoint(written[1], m2, opt2)
    assert (epoch, best) == (5, 0.875)
    for (k, a), (_, b) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert torch.equal(a, b), k
    assert opt2.param_groups[0]["lr"] == 0.1
    assert torch.equal(opt2.state_dict()["state"][0]["momentum_buffer"], opt.state_dict()["state"][0]["momentum_buffer"])

    # `module.`-prefixed keys and wrapped models
    torch.save({"state_dict": {"module." + k: v for k, v in m.state_dict().items()}}, str(tmp_path / "wrapped.pth.tar"))

    class Wrapper(object):
        def __init__(self, module):
            self.module = module
    m3 = hctr_model(37)
    assert ck.load_checkpoint(str(tmp_path / "wrapped.pth.tar"), Wrapper(m3)) == (0, 0.0)
    assert torch.equal(m3.state_dict()["linear.weight"], m.state_dict()["linear.weight"])
    with pytest.raises(FileNotFoundError):
        ck.load_checkpoint(str(tmp_path / "missing.pth.tar"), m3)
