"""Pin the CPU oracle (oracle/) against outputs of the reference itself (tests/golden/*.npz, produced by
tests/golden/make_golden.py from /root/reference). CPU-only."""
import numpy as np
import pytest

import oracle
import synth
from oracle.codec import CodecTables


def _texts(C, idx, ln):
    return CodecTables(synth.charset(C - 2)).to_text(idx, ln)


@pytest.mark.parametrize("name", ["small", "mid", "wide", "t1"])
def test_greedy_matches_reference(golden, name):
    g = golden("greedy")
    T, B, C, seed = [int(v) for v in g[name + "_shape"]]
    x = synth.ctc_like_logits(T, B, C, seed)
    _, idx, ln = oracle.greedy_decode(x)
    assert _texts(C, idx, ln) == list(g[name + "_text"])


def test_greedy_edge_cases(golden):
    g = golden("greedy")
    x = g["edge_logits"]
    raw, idx, ln = oracle.greedy_decode(x)
    assert _texts(x.shape[2], idx, ln) == list(g["edge_text"])
    assert raw[0, 0] == 3 and raw[0, 1] == 3          # exact ties -> lowest index
    assert raw[2, 0] == 4 and raw[2, 2] == 2          # NaN beats everything, first NaN wins
    assert ln[1] == 0                                 # all-blank sequence -> empty string


def test_encode_matches_reference(golden):
    g = golden("greedy")
    tab = CodecTables(str(g["enc_chars"][0]))
    idx, ln = tab.encode(list(g["enc_texts"]))
    assert idx.dtype == np.int32 and ln.dtype == np.int32
    assert np.array_equal(idx, g["enc_idx"]) and np.array_equal(ln, g["enc_len"])


BEAM_CASES = [(c, s) for c in ("small", "mid", "wide") for s in ("zero_b0", "zero_b58", "tab_p2", "tab_p08")]


@pytest.mark.parametrize("case,setting", BEAM_CASES)
def test_beam_matches_reference(golden, case, setting):
    g = golden("beam")
    T, B, C, seed, period = [int(v) for v in g[case + "_shape"]]
    tseed, pen, bonus = g["%s_%s_cfg" % (case, setting)]
    x = synth.beam_logits(T, B, C, seed, period)
    table = None if tseed < 0 else synth.lm_table(C, int(tseed))
    idx, ln, st = oracle.beam_search(x, 10, 10, pen, bonus, table)
    assert (st == 0).all()
    assert _texts(C, idx, ln) == list(g["%s_%s_text" % (case, setting)])


def test_beam_narrow(golden):
    g = golden("beam")
    T, B, C, seed, period = [int(v) for v in g["narrow_shape"]]
    x = synth.beam_logits(T, B, C, seed, period)
    idx, ln, st = oracle.beam_search(x, 4, 6, 2.0, 1.5, synth.lm_table(C, 33))
    assert _texts(C, idx, ln) == list(g["narrow_text"])


def test_beam_empty_greedy_path_is_index_error():
    x = np.zeros((5, 1, 9), np.float32)
    x[:, 0, 0] = 10.0                                  # all blank: the reference raises IndexError (:198)
    _, _, st = oracle.beam_search(x, 10, 5, 2.0, 5.8, None)
    assert st[0] == -4


@pytest.mark.parametrize("name", ["small", "mid", "wide", "len1", "infeasible"])
def test_ctc_loss_matches_torch_reference(golden, name):
    g = golden("ctc_loss")
    T, B, C, seed, Lmin, Lmax = [int(v) for v in g[name + "_shape"]]
    x = synth.ctc_like_logits(T, B, C, seed, peak=4.0)
    tg, tl = synth.ctc_targets(B, C, Lmin, Lmax, seed + 100, repeat_frac=0.3)
    loss, nll, grad = oracle.ctc_loss(x, tg, [T] * B, tl)
    ref = float(g[name + "_loss"])
    assert abs(loss - ref) <= 2e-5 * max(1.0, abs(ref))      # torch computes in fp32; oracle in fp64
    if name + "_grad" in g:
        assert np.abs(grad - g[name + "_grad"]).max() <= 1e-5
    else:
        assert np.abs(grad.reshape(-1)[g[name + "_grad_pick"]] - g[name + "_grad_vals"]).max() <= 1e-5
        assert np.abs(np.abs(grad).sum(2) - g[name + "_grad_rowabs"]).max() <= 1e-4
        assert np.abs(grad[:, :, g[name + "_grad_tgt_cls"]] - g[name + "_grad_tgt"]).max() <= 1e-5
    if name == "infeasible":
        assert loss == 0.0 and np.abs(grad).max() == 0.0     # zero_infinity=True


SKIP_CASES = [(c, s) for c in ("small", "mid", "wide", "flat", "big") for s in ("zero_b0", "zero_b58", "tab_p2")]


@pytest.mark.parametrize("case,setting", SKIP_CASES)
def test_skip_search_matches_reference(golden, case, setting):
    g = golden("beam_skip")
    T, B, C, seed, period = [int(v) for v in g[case + "_shape"]]
    noise, boost = float(g[case + "_noise"]), float(g[case + "_boost"])
    x = synth.beam_logits(T, B, C, seed, period)
    if noise != 2.0:
        x = (x * (noise / 2.0)).astype(np.float32)
    x = synth.peakier(x, boost)
    tseed, pen, bonus = g["%s_%s_cfg" % (case, setting)]
    table = None if tseed < 0 else synth.lm_table(C, int(tseed))
    idx, ln, st = oracle.beam_search_skip(x, 10, pen, bonus, table)
    assert (st == 0).all()
    assert _texts(C, idx, ln) == list(g["%s_%s_text" % (case, setting)])


def test_edit_distance_known_answers():
    # classic known answers (the `editdistance` package's own doctest values)
    assert oracle.edit_distance("kitten", "sitting") == 3
    assert oracle.edit_distance("flaw", "lawn") == 2
    assert oracle.edit_distance("", "abc") == 3 and oracle.edit_distance("abc", "") == 3 and oracle.edit_distance("", "") == 0
    assert oracle.edit_distance("intention", "execution") == 5
    assert oracle.edit_distance([1, 2, 3], [1, 2, 3]) == 0
