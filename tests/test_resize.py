"""cv2.resize(..., INTER_AREA) to height 128 (the reference's resize in front of NormalizePAD, utils/dataset.py:53-57,
test.py:206-214): the numpy oracle against outputs of cv2 itself (tests/golden/resize.npz, CPU), and the CUDA kernel against
both (GPU). Bit-exact: the results are bytes."""
import numpy as np
import pytest
import torch

import synth
from oracle import resize as oresize


def _cases(g):
    for i, (sh, sw, seed) in enumerate(synth.RESIZE_CASES):
        yield "case%d_%dx%d" % (i, sh, sw), synth.resize_source(sh, sw, seed), g["dst%d" % i], "dataset"
    i = 0
    while "img_src%d" % i in g.files:
        yield "bundled%d" % i, g["img_src%d" % i], g["img_dst%d" % i], "test"
        i += 1


def test_oracle_matches_cv2_outputs(golden):
    g = golden("resize")
    n = 0
    for name, src, want, rule in _cases(g):
        got = oresize.resize_area(src, want.shape[1], want.shape[0])
        assert got.shape == want.shape and np.array_equal(got, want), name
        n += 1
    assert n == len(synth.RESIZE_CASES) + 5


def test_resized_width_rules():
    from hctr_b200.pipeline import resized_width
    assert resized_width(48, 1318, 128, "test") == 3514 and resized_width(53, 376, 128, "test") == 908
    assert resized_width(359, 500, 128, "dataset") == 178
    with pytest.raises(ValueError):
        resized_width(10, 10, 128, "other")


@pytest.mark.gpu
def test_device_resize_bit_exact_vs_cv2_and_oracle(golden):
    from hctr_b200.pipeline import resize_line, resized_width
    g = golden("resize")
    for name, src, want, rule in _cases(g):
        got = resize_line(src, 128, rule, device="cuda")
        assert tuple(got.shape) == want.shape, name
        assert np.array_equal(got.cpu().numpy(), want), name
    # more sizes against the oracle, a pitched (non-contiguous) source, and another target height
    rs = np.random.RandomState(77)
    for _ in range(12):
        sh, sw = int(rs.randint(6, 400)), int(rs.randint(6, 1200))
        src = rs.randint(0, 256, size=(sh, sw)).astype(np.uint8)
        got = resize_line(src, 128, "dataset", device="cuda").cpu().numpy()
        assert np.array_equal(got, oresize.resize_area(src, resized_width(sh, sw, 128, "dataset"), 128)), (sh, sw)
    big = torch.from_numpy(rs.randint(0, 256, size=(200, 900)).astype(np.uint8)).cuda()
    view = big[10:170, 33:700]                                                     # row stride 900, unit column stride
    got = resize_line(view, 64, "dataset").cpu().numpy()
    assert np.array_equal(got, oresize.resize_area(view.cpu().numpy(), resized_width(160, 667, 64), 64))
    with pytest.raises(ValueError):
        resize_line(big.float(), 128)


def test_oracle_matches_cv2_live_where_cv2_is_installed():
    """Beyond the committed fixtures: wherever opencv-python is importable (the build container and the GPU image have it),
    the restatement is compared with cv2.resize itself on fresh random shapes, including exact integer factors."""
    cv2 = pytest.importorskip("cv2")
    rs = np.random.RandomState(2024)
    shapes = [(int(rs.randint(6, 420)), int(rs.randint(6, 900))) for _ in range(30)] + [(384, 301), (512, 260), (128, 77), (64, 640)]
    for sh, sw in shapes:
        src = rs.randint(0, 256, size=(sh, sw)).astype(np.uint8)
        dw = max(1, int(sw * (128 / sh)))
        want = cv2.resize(src, (dw, 128), interpolation=cv2.INTER_AREA)
        assert np.array_equal(oresize.resize_area(src, dw, 128), want), (sh, sw, dw)
