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


# ---- back-off n-gram LM (oracle/ngram.py; kenlm is absent: pinned to hand-computed ARPA values only) -----------------
_TINY_ARPA = """\\data\\
ngram 1=6
ngram 2=4
ngram 3=2

\\1-grams:
-1.0\t<unk>
-99\t<s>\t-0.5
-1.5\t</s>
-0.7\ta\t-0.25
-0.9\tb\t-0.125
-1.1\tc

\\2-grams:
-0.3\t<s> a\t-0.0625
-0.4\ta b\t-0.03125
-0.6\tb a
-0.2\tb c

\\3-grams:
-0.1\t<s> a b
-0.05\ta b c

\\end\\
"""


def _f32(*terms):
    t = np.float32(0.0)
    for v in terms:
        t = np.float32(t + np.float32(v))
    return float(t)


def test_ngram_oracle_known_answers():
    """Katz back-off by hand: p(w|ctx) = longest match + back-offs of the longer contexts that exist."""
    from oracle.ngram import ArpaLM
    lm = ArpaLM(_TINY_ARPA)
    assert lm.order == 3
    # "a b c": p(a|<s>) = -0.3;  p(b|<s> a) = -0.1 (trigram);  p(c|a b) = -0.05 (trigram)
    assert lm.score("a b c") == _f32(-0.3, -0.1, -0.05)
    # "b a": p(b|<s>): no "<s> b" -> unigram -0.9 + backoff(<s>) -0.5;  p(a|<s> b): "<s> b a" missing, "b a" = -0.6, and the
    # longer context "<s> b" does not exist -> no back-off
    assert lm.score("b a") == _f32(np.float32(np.float32(-0.9) + np.float32(-0.5)), -0.6)
    # "a c": p(c|<s> a): no "<s> a c", no "a c" -> unigram -1.1 + backoff(a) -0.25 + backoff(<s> a) -0.0625
    assert lm.score("a c") == _f32(-0.3, np.float32(np.float32(np.float32(-1.1) + np.float32(-0.25)) + np.float32(-0.0625)))
    # unknown word: <unk> unigram, backed off from the context; then it becomes context itself (no n-gram continues it)
    assert lm.score("z a") == _f32(np.float32(np.float32(-1.0) + np.float32(-0.5)), -0.7)
    assert lm.score("") == 0.0
    assert lm.score("a", eos=True) == _f32(-0.3, np.float32(np.float32(np.float32(-1.5) + np.float32(-0.25)) + np.float32(-0.0625)))


def test_ngram_table_builder_matches_oracle():
    """The product's ARPA reader + hash table (host side of csrc/ngram_lm.cuh) holds exactly the oracle's n-grams."""
    from oracle.ngram import ArpaLM
    from hctr_b200.ngram_lm import NgramLM
    C = 64
    chars = synth.charset(C - 2)
    text = synth.arpa_text(chars[:40], 5, 3, grams_per_order=250)
    o = ArpaLM(text)
    cd = {ch: i + 1 for i, ch in enumerate(chars)}
    lm = NgramLM.from_arpa_text(text, cd, C)
    assert lm.order == 5 and lm.n_grams == len(o.grams)
    w2i = dict(cd); w2i["<s>"] = C; w2i["</s>"] = C + 1; w2i["<unk>"] = C + 2
    for g, (p, b) in o.grams.items():
        assert lm.host_find([w2i[w] for w in g]) == (float(p), float(b)), g
    assert lm.host_find([w2i[chars[0]], w2i[chars[41]]]) is None
    # characters without a unigram map to <unk>, the others to themselves
    have = {g[0] for g in o.grams if len(g) == 1}
    for ch in chars:
        assert lm.host["vocab"][cd[ch]] == (cd[ch] if ch in have else C + 2)
    with pytest.raises(ValueError):
        NgramLM.from_arpa_text(text.replace("\\5-grams:", "\\6-grams:"), cd, C)


# ---------------------------------------------------------------------------------------------------------------------
# model forward: oracle/hctr_forward.py pinned on the CPU to logits the REFERENCE module produced (make_golden.make_model)

def _seeded_state_dict(num_classes, seed):
    """Seed-identical parameters of the reference constructor (hash-checked against the reference's in test_abi_cpu /
    below); only the parameter HOLDER of the product is used here - the arithmetic under test is the oracle's."""
    import torch
    from hctr_b200.models.handwritten_ctr_model import hctr_model
    torch.manual_seed(seed)
    return {k: v.detach().clone() for k, v in hctr_model(num_classes).state_dict().items()}


def _sd_hash(sd):
    import hashlib
    h = hashlib.sha256()
    for k, v in sd.items():
        h.update(k.encode()); h.update(v.numpy().tobytes())
    return h.hexdigest()


def test_forward_oracle_matches_reference_logits_small_model(golden):
    import torch
    from oracle import hctr_forward
    g = golden("model")
    sd = _seeded_state_dict(37, 4321)
    assert _sd_hash(sd) == str(g["small_sd_hash"][0])
    x = torch.from_numpy(synth.text_lines(2, 72, 51))
    with torch.no_grad():
        y = hctr_forward.forward(x, sd)
    ref = torch.from_numpy(g["small_default_logits"])
    assert tuple(y.shape) == tuple(ref.shape) == (72, 2, 37)
    assert (y - ref).abs().max().item() <= 1e-5
    # BN-calibrated regime (logits O(1), the only regime where they are not ~ linear.bias): running statistics from the
    # reference's own train-mode pass
    cal = dict(sd)
    for k in sd:
        if "small_cal." + k in g:
            cal[k] = torch.from_numpy(g["small_cal." + k])
    with torch.no_grad():
        yc = hctr_forward.forward(x, cal)
    refc = torch.from_numpy(g["small_cal_logits"])
    assert refc.abs().max().item() > 1.0
    assert (yc - refc).abs().max().item() <= 1e-3 * refc.abs().max().item()
    assert (yc.argmax(2) == refc.argmax(2)).float().mean().item() >= 0.99
    # and the oracle's own calibration pass reproduces the reference's running statistics
    own = hctr_forward.calibrate_bn(sd, torch.from_numpy(synth.text_lines(3, 96, 52)))
    for k in sd:
        if "small_cal." + k in g:
            want = torch.from_numpy(g["small_cal." + k])
            assert torch.allclose(own[k], want, rtol=1e-3, atol=1e-6), k


def test_forward_oracle_matches_reference_full_charset(golden):
    import torch
    from oracle import hctr_forward
    from oracle.codec import CodecTables
    g = golden("model")
    sd = _seeded_state_dict(7375, 1234)
    assert _sd_hash(sd) == str(g["sd_hash_7375_seed1234"][0])
    x = torch.from_numpy(synth.text_lines(1, 136, 53))
    with torch.no_grad():
        y = hctr_forward.forward(x, sd)
    assert (y[0, 0] - torch.from_numpy(g["full_default_logits_t0"])).abs().max().item() <= 1e-5
    assert abs(y.abs().max().item() - float(g["full_default_absmax"])) <= 1e-5
    _, idx, ln = oracle.greedy_decode(y.contiguous().numpy())
    assert CodecTables(synth.charset(7373)).to_text(idx, ln) == list(g["full_default_text"])


def test_ngram_oracle_matches_real_kenlm():
    """oracle/ngram.py against scores of the real KenLM library (tests/golden/make_kenlm_golden.py; needs `pip install kenlm`,
    which the build container cannot do). Skipped - and the n-gram scorer stays 'parity-unpinned against the real library' in
    DESIGN.md - until tests/golden/kenlm_scores.npz is generated."""
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kenlm_scores.npz")
    if not os.path.exists(path):
        pytest.skip("tests/golden/kenlm_scores.npz not generated (kenlm is not installable here: no network)")
    from oracle.ngram import ArpaLM
    g = np.load(path, allow_pickle=True)
    for name in ("bi", "tri", "five"):
        lm = ArpaLM(str(g[name + "_arpa"][0]))
        got = np.array([lm.score(s) for s in g[name + "_sentences"]])
        assert np.array_equal(got.astype(np.float32), g[name + "_scores"].astype(np.float32)), name
        if name + "_scores_binary" in g:
            assert np.array_equal(g[name + "_scores"], g[name + "_scores_binary"]), name      # ARPA and build_binary agree
