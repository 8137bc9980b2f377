"""Width bucketing / batch sharding (host logic, CPU) and the device-side NormalizePAD + sharded recognition (GPU)."""
import os
import sys

import numpy as np
import pytest
import torch

import synth


def _widths(n, seed):
    rs = np.random.RandomState(seed)
    return (64 * rs.randint(4, 65, size=n)).tolist()            # BASELINE config 3: multiples of 64 in [256, 4096]


def test_bucketing_properties():
    from hctr_b200.pipeline import bucket_lines
    widths = _widths(500, 0)
    batches = bucket_lines(widths, multiple=256, column_budget=131072)
    seen = sorted(i for _, idx in batches for i in idx)
    assert seen == list(range(500))                              # every line exactly once
    for wb, idx in batches:
        assert wb % 256 == 0 and all(widths[i] <= wb < widths[i] + 256 for i in idx)
        assert len(idx) * wb <= 131072
    costs = [wb * len(idx) for wb, idx in batches]
    assert costs == sorted(costs, reverse=True)                  # heaviest first
    padded = sum(costs)
    assert padded <= 1.12 * sum(widths)                          # padding overhead of 256-wide buckets stays small
    assert bucket_lines([300000], 256, 131072) == [(300032, [0])]
    with pytest.raises(ValueError):
        bucket_lines([0, 5])


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_sharding_is_balanced_and_disjoint(world):
    from hctr_b200.pipeline import bucket_lines, shard_batches
    widths = _widths(4096, 0)
    batches = bucket_lines(widths)
    shards = shard_batches(batches, world)
    flat = sorted(b for s in shards for b in s)
    assert flat == list(range(len(batches)))
    loads = [sum(batches[b][0] * len(batches[b][1]) for b in s) for s in shards]
    assert max(loads) <= 1.05 * (sum(loads) / world) + 131072    # LPT: within one batch of the mean
    assert shard_batches(batches, world) == shards               # deterministic


@pytest.mark.gpu
def test_normalize_pad_bit_exact_vs_reference(golden):
    from hctr_b200.pipeline import make_batch
    g = golden("pad")
    imgs = [g["img%d" % i] for i in range(3)]
    x = make_batch(imgs, [0, 1, 2], 64, torch.device("cuda"))
    for i in range(3):
        assert torch.equal(x[i].cpu(), torch.from_numpy(g["pad%d" % i])), i


@pytest.mark.gpu
def test_sharded_recognition_matches_per_batch_reference_padding():
    """Two emulated ranks decode disjoint batches; every line's text equals the one obtained by padding its batch on the
    host exactly as NormalizePAD does and running the model on that tensor (and the union covers all lines)."""
    from hctr_b200.models.handwritten_ctr_model import hctr_model
    from hctr_b200.utils.ctc_codec import ctc_codec
    from hctr_b200.pipeline import bucket_lines, recognize_lines
    torch.manual_seed(21)
    m = hctr_model(101).cuda().eval()
    codec = ctc_codec(synth.charset(99))
    rs = np.random.RandomState(5)
    widths = (32 * rs.randint(2, 12, size=14)).tolist()
    imgs = [((synth.text_lines(1, w, 100 + i)[0, 0] * 0.5 + 0.5) * 255).round().astype(np.uint8) for i, w in enumerate(widths)]
    parts = [recognize_lines(m, codec, imgs, rank=r, world=2, multiple=128, column_budget=768) for r in range(2)]
    assert not (set(parts[0]) & set(parts[1])) and sorted(set(parts[0]) | set(parts[1])) == list(range(14))
    got = {**parts[0], **parts[1]}
    for wb, idx in bucket_lines(widths, 128, 768):
        batch = []
        for i in idx:
            t = torch.from_numpy(imgs[i]).float().div(255).sub(0.5).div(0.5)
            pad = t[:, -1:].expand(128, wb - t.shape[1])
            batch.append(torch.cat([t, pad], 1))
        x = torch.stack(batch).unsqueeze(1).cuda()
        texts = codec.decode(m(x))
        assert [got[i] for i in idx] == texts


@pytest.mark.gpu
def test_recognition_from_raw_lines_resizes_on_the_device(golden):
    """Raw grayscale lines of arbitrary height (the five bundled images, 48..77 px high): device resize + pad + model + decode
    gives the texts obtained from cv2-resized lines (fixture outputs of cv2 itself)."""
    from hctr_b200.models.handwritten_ctr_model import hctr_model
    from hctr_b200.utils.ctc_codec import ctc_codec
    from hctr_b200.pipeline import recognize_lines
    g = golden("resize")
    raw = [g["img_src%d" % i] for i in range(5)]
    resized = [g["img_dst%d" % i] for i in range(5)]
    torch.manual_seed(23)
    m = hctr_model(101).cuda().eval()
    codec = ctc_codec(synth.charset(99))
    a = recognize_lines(m, codec, raw, multiple=256, column_budget=8192, resize_height=128, resize_rule="test")
    b = recognize_lines(m, codec, resized, multiple=256, column_budget=8192)
    assert a == b and sorted(a) == list(range(5))
