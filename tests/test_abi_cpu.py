"""CPU-side checks of the boundary: the C-ABI library loads and exports exactly what include/hctr_b200.h
declares, the drop-in modules keep the reference's surface, and the product never routes through oracle/."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import hctr_b200
from hctr_b200 import native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols(header="hctr_b200.h"):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hctr_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    if not os.path.exists(native.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    lib = ctypes.CDLL(native.LIB_PATH)
    syms = _declared_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), "libhctr_b200.so does not export %s" % s
    # and the Python binding table covers the header one to one
    assert sorted(native.SIGNATURES) == syms
    assert not [s for s in syms if "debug" in s or "testing" in s]          # no diagnostic switches in the product ABI
    hooks = _declared_symbols("hctr_b200_testing.h")
    assert sorted(native.TESTING_SIGNATURES) == hooks and all(hasattr(lib, s) for s in hooks)
    os.environ.pop("HCTR_TEST_HOOKS", None)
    assert native.lib().hctr_testing_set_conv_variant(1, 1) == native.HCTR_ERR_UNSUPPORTED
    assert native.lib().hctr_abi_version() == 1


def test_argument_errors_are_reported_without_a_gpu():
    lib = native.lib()
    rc = lib.hctr_conv_bn_act_fwd(None, None, None, None, None, 1, 2, 3, 64, 64, 3, 1, 0, None)
    assert rc == native.HCTR_ERR_INVALID and "null" in native.last_error()
    rc = lib.hctr_ctc_topk_logsoftmax(None, 0, 4, 1, 10, 10, 10, 99, ctypes.c_void_p(8), ctypes.c_void_p(8), ctypes.c_void_p(8), None)
    assert rc == native.HCTR_ERR_INVALID and "search depth" in native.last_error()
    with pytest.raises(RuntimeError):
        native.check(rc, "topk")
    with pytest.raises(IndexError):
        native.check(native.HCTR_ERR_INDEX)


def test_model_surface_and_state_dict_layout(golden):
    from hctr_b200.models.handwritten_ctr_model import hctr_model
    g = golden("model")
    torch.manual_seed(1234)
    m = hctr_model(7375)
    sd = m.state_dict()
    assert list(sd.keys()) == list(g["sd_keys"])
    assert [str(tuple(v.shape)) for v in sd.values()] == list(g["sd_shapes"])
    assert len(sd) == 254 and sum(p.numel() for p in m.parameters()) == 53114383
    import hashlib
    h = hashlib.sha256()
    for k, v in sd.items():
        h.update(k.encode()); h.update(v.numpy().tobytes())
    assert h.hexdigest() == str(g["sd_hash_7375_seed1234"][0])      # seed-identical to the reference constructor
    assert (m.img_height, m.PAD, m.optimizer, m.pred, m.noutput) == (128, 'NormalizePAD', 'SGD', 'CTC', 7375)
    m.load_state_dict(sd, strict=True)
    m.eval()
    with pytest.raises(RuntimeError):                                # no CPU fallback
        m(torch.zeros(1, 1, 128, 32))


def test_codec_surface_and_encode(golden):
    from hctr_b200.utils.ctc_codec import ctc_codec
    g = golden("greedy")
    c = ctc_codec(str(g["enc_chars"][0]))
    idx, ln = c.encode(list(g["enc_texts"]))
    assert idx.dtype == np.int32 and ln.dtype == np.int32
    assert np.array_equal(idx, g["enc_idx"]) and np.array_equal(ln, g["enc_len"])
    assert c.characters[0] == '<blank>' and c.characters[-1] == '<unknown>'
    assert c.dict['<blank>'] == 0 and c.dict['<unknown>'] == len(c.characters) - 1
    assert (c.lm_panelty, c.len_bonus, c.search_depth, c.beam_size) == (2, 5.8, 10, 10)
    assert c.use_beam_search is False and c.use_tfm_pred is True and c.skip_search is False
    with pytest.raises(ImportError):
        c.set_beam_search()                                          # transformer LM deps are absent, as in the reference
    c.set_beam_search(use_tfm_pred=False, lm_panelty=0.8, len_bonus=4.8, beam_size=5, search_depth=7)
    assert c.use_beam_search and (c.lm_panelty, c.len_bonus, c.beam_size, c.search_depth) == (0.8, 4.8, 5, 7)


def test_drop_in_import_names():
    """With the package directory on sys.path the reference's own import lines resolve to the B200 modules."""
    import subprocess, sys
    pkg = os.path.join(ROOT, "handwritten-chinese-ocr-samples_b200")
    code = ("from models.handwritten_ctr_model import hctr_model; from utils.ctc_codec import ctc_codec; "
            "import sys; print(hctr_model.__module__, 'hctr_b200' in sys.modules)")
    out = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, PYTHONPATH=pkg), capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert out.stdout.split() == ["models.handwritten_ctr_model", "False"] or "models.handwritten_ctr_model" in out.stdout


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "handwritten-chinese-ocr-samples_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "liboracle" not in text, f
