"""Batch-sharded recognition of ragged text lines (BASELINE config 3).

Lines are independent given their padded width, so N GPUs = N replicas, each taking whole batches; no collective runs
on the data path (only an optional host-side gather of the strings). Batches are built by width bucketing:
a line of width w goes to the bucket ceil(w / multiple) * multiple and is right-padded by border replication on the
device (the reference's NormalizePAD, utils/dataset.py:78-93) - i.e. exactly what the reference computes when the
batch's maximum width equals the bucket width. The column budget bounds B * W_bucket per batch (131072 = 64 x 2048).
"""
import numpy as np
import torch

from . import native as nat


def bucket_lines(widths, multiple=256, column_budget=131072, max_batch=None):
    """Host logic. Returns a list of (bucket_width, [line indices]) batches, heaviest first.
    Every line appears exactly once; bucket_width >= every member's width; len(idx) * bucket_width <= column_budget
    (a single line wider than the budget still gets its own batch)."""
    widths = [int(w) for w in widths]
    if any(w <= 0 for w in widths):
        raise ValueError("line widths must be positive")
    by_bucket = {}
    for i, w in enumerate(widths):
        wb = (w + multiple - 1) // multiple * multiple
        by_bucket.setdefault(wb, []).append(i)
    batches = []
    for wb in sorted(by_bucket, reverse=True):
        per = max(1, column_budget // wb)
        if max_batch:
            per = min(per, max_batch)
        idx = by_bucket[wb]
        for s in range(0, len(idx), per):
            batches.append((wb, idx[s:s + per]))
    batches.sort(key=lambda b: (-b[0] * len(b[1]), b[1][0]))
    return batches


def shard_batches(batches, world):
    """Host logic. Greedy longest-processing-time assignment of batches (cost = padded columns) to `world` ranks.
    Returns a list of `world` lists of batch indices; deterministic, every batch assigned exactly once."""
    load = [0] * world
    out = [[] for _ in range(world)]
    for bi, (wb, idx) in enumerate(batches):            # batches arrive heaviest first
        r = min(range(world), key=lambda q: (load[q], q))
        out[r].append(bi)
        load[r] += wb * len(idx)
    return out


def resized_width(src_h, src_w, height=128, rule="dataset"):
    """Width of a line after the reference's resize to `height`: utils/dataset.py:54-55 (`int(width * (img_h / height))`,
    rule "dataset") or test.py:210-212 (`int(height * (float(w) / float(h)))`, rule "test")."""
    if rule == "dataset":
        return int(src_w * (height / src_h))
    if rule == "test":
        return int(height * (float(src_w) / float(src_h)))
    raise ValueError("rule must be 'dataset' or 'test'")


def resize_line(image, height=128, rule="dataset", device=None):
    """cv2.resize(image, (new_width, height), interpolation=cv2.INTER_AREA) on the GPU, bit-identical to OpenCV
    (reference: utils/dataset.py:53-57, test.py:206-214). image: uint8 [h,w] numpy array or CUDA tensor (any row stride).
    Returns a uint8 CUDA tensor [height, new_width]."""
    if isinstance(image, np.ndarray):
        if device is None:
            raise ValueError("resize_line: a device is needed for a host image")
        image = torch.from_numpy(np.ascontiguousarray(image, dtype=np.uint8)).to(device, non_blocking=True)
    if image.dtype != torch.uint8 or image.dim() != 2 or not image.is_cuda or image.stride(1) != 1:
        raise ValueError("resize_line: expected a uint8 [h,w] CUDA tensor with unit column stride")
    h, w = image.shape
    new_w = resized_width(h, w, height, rule)
    if new_w <= 0:
        raise ValueError("resize_line: the line collapses to zero width")
    with torch.cuda.device(image.device):
        out = torch.empty((height, new_w), dtype=torch.uint8, device=image.device)
        nat.check(nat.lib().hctr_resize_area_u8(nat.ptr(image), h, w, image.stride(0), nat.ptr(out), height, new_w, new_w,
                                                nat.stream_ptr()), "resize_area")
    return out


def make_batch(images, indices, width, device):
    """uint8 [128,w] lines (numpy arrays, or CUDA tensors as `resize_line` returns them) -> device fp32 [B,1,128,width],
    normalised and border-padded on the GPU."""
    H = images[indices[0]].shape[0]
    ws = np.array([images[i].shape[1] for i in indices], dtype=np.int32)
    if int(ws.max()) > width:
        raise ValueError("a line is wider than its bucket")
    offs = np.zeros(len(indices), dtype=np.int64)
    offs[1:] = np.cumsum(ws[:-1].astype(np.int64) * H)
    on_device = isinstance(images[indices[0]], torch.Tensor)
    if not on_device:
        flat = np.concatenate([np.ascontiguousarray(images[i], dtype=np.uint8).reshape(-1) for i in indices])
    with torch.cuda.device(device):
        if on_device:
            pix = torch.cat([images[i].to(device).contiguous().reshape(-1) for i in indices])
        else:
            pix = torch.from_numpy(flat).to(device, non_blocking=True)
        d_off = torch.from_numpy(offs).to(device, non_blocking=True)
        d_w = torch.from_numpy(ws).to(device, non_blocking=True)
        out = torch.empty((len(indices), 1, H, width), dtype=torch.float32, device=device)
        nat.check(nat.lib().hctr_normalize_pad(nat.ptr(pix), nat.ptr(d_off), nat.ptr(d_w), nat.ptr(out), len(indices), H, width,
                                               nat.stream_ptr()), "normalize_pad")
    return out


def recognize_lines(model, codec, images, rank=0, world=1, multiple=256, column_budget=131072, device=None,
                    resize_height=None, resize_rule="test"):
    """Decode this rank's share of `images` (list of uint8 [128,w] arrays). Returns {line index: text}.
    With `resize_height` (128 for the reference model) the images are raw grayscale lines of any height: widths are
    planned with the reference's rule (`resized_width`), and each of this rank's lines is resized on the device
    (`resize_line`, cv2.INTER_AREA bit-exact) right before its batch is padded - test.py:204-227 without the host resize."""
    device = device if device is not None else next(model.parameters()).device
    if resize_height is None:
        widths = [im.shape[1] for im in images]
    else:
        widths = [resized_width(im.shape[0], im.shape[1], resize_height, resize_rule) for im in images]
    batches = bucket_lines(widths, multiple, column_budget)
    mine = shard_batches(batches, world)[rank]
    result = {}
    with torch.no_grad():
        for bi in mine:
            wb, idx = batches[bi]
            if resize_height is None:
                x = make_batch(images, idx, wb, device)
            else:
                lines = {i: resize_line(images[i], resize_height, resize_rule, device=device) for i in idx}
                x = make_batch(lines, idx, wb, device)
            if codec.use_beam_search or model.training or len(codec.characters) != model.noutput:
                texts = codec.decode(model(x))
            else:                       # greedy: the arg-max runs in the classifier epilogue, the logits are never written
                texts = codec.indices_to_text(*model.greedy_decode(x))
            for i, t in zip(idx, texts):
                result[i] = t
    return result
