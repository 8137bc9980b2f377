"""GPU parity of the CTC codec kernels (through the C ABI) against the oracle and the reference's golden outputs.
Bit-exact for indices / strings; log-probs within 2e-6 (fp32 log-sum-exp)."""
import numpy as np
import pytest
import torch

import oracle
import synth

pytestmark = pytest.mark.gpu


def _codec(C):
    from hctr_b200.utils.ctc_codec import ctc_codec
    return ctc_codec(synth.charset(C - 2))


@pytest.mark.parametrize("name", ["small", "mid", "wide", "t1"])
def test_greedy_golden(golden, name):
    g = golden("greedy")
    T, B, C, seed = [int(v) for v in g[name + "_shape"]]
    x = synth.ctc_like_logits(T, B, C, seed)
    c = _codec(C)
    assert c.decode(x) == list(g[name + "_text"])                       # numpy in, as the reference's callers do
    assert c.decode(torch.from_numpy(x).cuda()) == list(g[name + "_text"])   # device tensor: logits stay on the GPU


def test_greedy_edge_cases(golden):
    g = golden("greedy")
    x = g["edge_logits"]
    c = _codec(x.shape[2])
    assert c.decode(x) == list(g["edge_text"])
    idx, ln, raw = c.greedy_indices(torch.from_numpy(x).cuda(), return_argmax=True)
    r, i, l = oracle.greedy_decode(x)
    assert np.array_equal(raw.cpu().numpy(), r) and np.array_equal(ln.cpu().numpy(), l)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("T,B,C", [(257, 5, 7375), (33, 3, 101), (64, 2, 8)])
def test_greedy_vs_oracle_layouts(dtype, T, B, C):
    x = synth.ctc_like_logits(T, B, C, 7 + T)
    xt = torch.from_numpy(x).cuda().to(dtype)
    xf = xt.float().cpu().numpy()                                     # exact upcast: argmax is invariant
    r, i, l = oracle.greedy_decode(xf)
    c = _codec(C)
    # [T,B,C] contiguous, and the model's layout: a (1,0,2) view of a padded-pitch [B,T,Cpad] tensor
    pitch = (C + 7) // 8 * 8
    padded = torch.full((B, T, pitch), 1e9, dtype=dtype, device="cuda")
    padded[:, :, :C] = xt.permute(1, 0, 2)
    for view in (xt, padded[:, :, :C].permute(1, 0, 2)):
        idx, ln, raw = c.greedy_indices(view, return_argmax=True)
        assert np.array_equal(raw.cpu().numpy(), r)
        assert np.array_equal(ln.cpu().numpy(), l)
        for b in range(B):
            assert np.array_equal(idx[b, :l[b]].cpu().numpy(), i[b, :l[b]])


def test_greedy_full_size_properties():
    """BASELINE config 2 size (T=2048, B=64, C=7375, bf16): size-independent checks - the planted path comes back,
    and decoding is idempotent under a monotone transform of the logits."""
    T, B, C = 2048, 64, 7375
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn(B, T, 7376, generator=g, device="cuda", dtype=torch.float32).to(torch.bfloat16)
    path = torch.randint(0, C, (B, T), generator=g, device="cuda")
    x.scatter_(2, path.unsqueeze(2), 30.0)
    c = _codec(C)
    view = x[:, :, :C].permute(1, 0, 2)
    idx, ln, raw = c.greedy_indices(view, return_argmax=True)
    assert torch.equal(raw.long(), path)
    idx2, ln2 = c.greedy_indices((view.float() * 0.5 + 3.0).to(torch.bfloat16))
    assert torch.equal(ln, ln2) and torch.equal(idx, idx2)
    keep = (path != 0) & (path != C - 1)
    keep[:, 1:] &= path[:, 1:] != path[:, :-1]
    assert torch.equal(ln.long(), keep.sum(1))


def test_greedy_empty_inputs():
    c = _codec(30)
    assert c.decode(np.zeros((0, 3, 30), np.float32)) == []          # T == 0: the reference appends nothing (:85-86)
    assert c.decode(np.zeros((4, 0, 30), np.float32)) == []


def test_topk_logsoftmax_vs_oracle():
    from hctr_b200 import native as nat
    T, B, C, k = 40, 3, 7375, 10
    x = synth.beam_logits(T, B, C, 77, 4)
    lp = oracle.log_softmax(x)
    tk = oracle.topk(lp, k)
    for dt in (torch.float32, torch.bfloat16):
        xt = torch.from_numpy(x).cuda().to(dt)
        ti = torch.empty((T, B, k), dtype=torch.int32, device="cuda")
        tp = torch.empty((T, B, k), dtype=torch.float32, device="cuda")
        lse = torch.empty((T, B), dtype=torch.float32, device="cuda")
        nat.check(nat.lib().hctr_ctc_topk_logsoftmax(nat.ptr(xt), nat.HCTR_F32 if dt == torch.float32 else nat.HCTR_BF16,
                                                     T, B, C, xt.stride(0), xt.stride(1), k, nat.ptr(ti), nat.ptr(tp),
                                                     nat.ptr(lse), nat.stream_ptr()))
        if dt == torch.float32:
            assert np.array_equal(ti.cpu().numpy(), tk)
            ref = np.take_along_axis(lp, tk, axis=2)
            assert np.abs(tp.cpu().numpy() - ref).max() <= 2e-6 * max(1.0, np.abs(ref).max())
        else:
            xf = xt.float().cpu().numpy()
            lpb = oracle.log_softmax(xf)
            got = ti.cpu().numpy()
            # bf16 rounding creates ties; compare the VALUES selected, and the order only where they differ
            assert np.array_equal(np.take_along_axis(xf, got, 2), np.take_along_axis(xf, oracle.topk(lpb, k), 2))


def _run_topk(xt, k):
    from hctr_b200 import native as nat
    T, B, C = xt.shape
    ti = torch.empty((T, B, k), dtype=torch.int32, device="cuda")
    tp = torch.empty((T, B, k), dtype=torch.float32, device="cuda")
    lse = torch.empty((T, B), dtype=torch.float32, device="cuda")
    nat.check(nat.lib().hctr_ctc_topk_logsoftmax(nat.ptr(xt), nat.HCTR_F32 if xt.dtype == torch.float32 else nat.HCTR_BF16,
                                                 T, B, C, xt.stride(0), xt.stride(1), k, nat.ptr(ti), nat.ptr(tp),
                                                 nat.ptr(lse), nat.stream_ptr()))
    torch.cuda.synchronize()
    return ti.cpu().numpy(), tp.cpu().numpy(), lse.cpu().numpy()


@pytest.mark.parametrize("T,B,C,k", [(700, 5, 301, 10), (97, 11, 7375, 10), (33, 2, 5, 3), (64, 3, 40001, 16), (300, 4, 17, 1)])
def test_topk_logsoftmax_many_rows_per_cta_and_odd_shapes(T, B, C, k):
    """More rows than resident CTAs (every CTA walks several rows through both shared-memory buffers), rows that are
    not 16-byte aligned (odd C), tiny rows (all elements through registers), rows too large for two buffers."""
    rng = np.random.default_rng(C * 7 + T)
    x = (3.0 * rng.standard_normal((T, B, C))).astype(np.float32)
    lp = oracle.log_softmax(x)
    tk = oracle.topk(x, k)              # ranked on the raw values (log_softmax rounding can merge neighbours)
    ti, tp, lse = _run_topk(torch.from_numpy(x).cuda(), k)
    assert np.array_equal(ti, tk)
    ref = np.take_along_axis(lp, tk, axis=2)
    assert np.abs(tp - ref).max() <= 2e-6 * max(1.0, np.abs(ref).max())
    m = x.max(axis=2)
    ref_lse = m + np.log(np.exp(x - m[..., None]).sum(axis=2, dtype=np.float64))
    assert np.abs(lse - ref_lse).max() <= 1e-5 * max(1.0, np.abs(ref_lse).max())
    # bf16 rows (2-byte aligned starts): compare the selected VALUES (rounding creates ties)
    xb = torch.from_numpy(x).cuda().to(torch.bfloat16)
    tib, _, _ = _run_topk(xb, k)
    xf = xb.float().cpu().numpy()
    assert np.array_equal(np.take_along_axis(xf, tib, 2), np.take_along_axis(xf, oracle.topk(oracle.log_softmax(xf), k), 2))


def test_topk_logsoftmax_strided_views_and_constant_rows():
    """The model's logits are a permuted, pitched view ([B,W,pitch] -> [W,B,C]); constant rows overflow the candidate
    list and take the exact arg-max fallback (ties -> lowest index, numpy argsort of the reversed stable order aside)."""
    T, B, C, k = 50, 6, 1001, 10
    rng = np.random.default_rng(5)
    base = torch.from_numpy((2.0 * rng.standard_normal((B, T, C + 3))).astype(np.float32)).cuda()
    view = base[:, :, :C].permute(1, 0, 2)                       # [T,B,C], strides (C+3, T*(C+3), 1)
    ti, tp, _ = _run_topk(view, k)
    x = view.cpu().numpy()
    lp = oracle.log_softmax(np.ascontiguousarray(x))
    tk = oracle.topk(x, k)
    assert np.array_equal(ti, tk)
    assert np.abs(tp - np.take_along_axis(lp, tk, 2)).max() <= 2e-6 * np.abs(lp).max()
    const = torch.zeros((3, 2, 777), dtype=torch.float32, device="cuda")
    const[1, 1, 5] = 1.0
    ti, tp, _ = _run_topk(const, 4)
    assert ti[0, 0].tolist() == [0, 1, 2, 3] and ti[1, 1].tolist() == [5, 0, 1, 2]
    assert np.allclose(tp[0, 0], -np.log(777.0), rtol=1e-6)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_topk_logsoftmax_adversarial_rows(dtype):
    """Rows that defeat the running top-k bound of the one-read kernel: ascending rows (every vector of every chunk reaches the
    bound of the chunks before it: the marked-vector list overflows -> exact fallback), descending rows (the first chunk holds
    the whole top k), a late spike after a flat start, -inf-padded rows, rows whose top k sit in one lane's vectors, heavy ties.
    Exact indices for any of them (value desc, index asc)."""
    T, B, C, k = 9, 2, 7375, 10
    rng = np.random.default_rng(11)
    x = np.zeros((T, B, C), dtype=np.float32)
    ramp = np.linspace(-8.0, 8.0, C, dtype=np.float32)
    x[0] = ramp                                        # ascending
    x[1] = ramp[::-1]                                  # descending
    x[2] = -3.0; x[2, :, C - 5:] = np.arange(5, dtype=np.float32)[None, :] + 4.0     # flat, spike in the last vector
    x[3] = rng.standard_normal((B, C)).astype(np.float32); x[3, :, 100:] = -np.inf    # mostly -inf
    x[4] = rng.standard_normal((B, C)).astype(np.float32); x[4, :, :4096] = -np.inf   # the first chunks are all -inf
    x[5] = rng.standard_normal((B, C)).astype(np.float32)
    x[5, :, 8 * 32 * 3: 8 * 32 * 3 + 16] += 20.0       # 16 large elements in two adjacent vectors of one lane pair
    x[6] = np.round(rng.standard_normal((B, C)) * 2).astype(np.float32)               # heavy ties (few distinct values)
    x[7] = rng.standard_normal((B, C)).astype(np.float32) * 30.0                       # large dynamic range
    x[8] = ramp; x[8, :, ::2] = ramp[::-1][::2]        # interleaved up / down
    xt = torch.from_numpy(x).cuda().to(dtype)
    xf = xt.float().cpu().numpy()
    ti, tp, lse = _run_topk(xt, k)
    want = oracle.topk(xf, k)
    assert np.array_equal(ti, want)
    m = xf.max(axis=2)
    ref_lse = m + np.log(np.exp((xf - m[..., None]).astype(np.float64)).sum(axis=2))
    assert np.abs(lse - ref_lse).max() <= 1e-5 * max(1.0, np.abs(ref_lse).max())
    ref = np.take_along_axis(xf, want, 2).astype(np.float64) - ref_lse[..., None]
    fin = np.isfinite(ref)
    assert np.array_equal(np.isfinite(tp), fin)
    assert np.abs(tp[fin] - ref[fin]).max() <= 4e-6 * max(1.0, np.abs(ref[fin]).max())


def test_topk_logsoftmax_kernel_variants_agree(monkeypatch):
    """The default one-read chunk kernel against the kernels it replaced (CTA per row, two-pass warp per row), misaligned rows
    and every search depth: identical indices, log-probs to rounding."""
    T, B, C = 23, 3, 7375
    rng = np.random.default_rng(3)
    base = torch.from_numpy((2.5 * rng.standard_normal((T, B, C + 1))).astype(np.float32)).cuda()
    for dtype in (torch.float32, torch.bfloat16):
        buf = base.to(dtype)
        view = buf[:, :, 1:]                                      # rows start 1 element off the 16-byte grid
        for k in (1, 5, 10, 16):
            res = {}
            for name, env in (("chunk", {}), ("cta", {"HCTR_TOPK_CHUNK": "0", "HCTR_TOPK_WARP": "0"}),
                              ("warp", {"HCTR_TOPK_CHUNK": "0", "HCTR_TOPK_WARP": "1"})):
                for key in ("HCTR_TOPK_CHUNK", "HCTR_TOPK_WARP"):
                    monkeypatch.delenv(key, raising=False)
                for key, val in env.items():
                    monkeypatch.setenv(key, val)
                res[name] = _run_topk(view, k)
            for name in ("cta", "warp"):
                assert np.array_equal(res["chunk"][0], res[name][0]), (dtype, k, name)
                assert np.abs(res["chunk"][1] - res[name][1]).max() <= 4e-6
                assert np.abs(res["chunk"][2] - res[name][2]).max() <= 4e-6
            xf = view.float().cpu().numpy()
            assert np.array_equal(res["chunk"][0], oracle.topk(xf, k))


BEAM_CASES = [(c, s) for c in ("small", "mid", "wide") for s in ("zero_b0", "zero_b58", "tab_p2", "tab_p08")]


@pytest.mark.parametrize("case,setting", BEAM_CASES)
def test_beam_golden(golden, case, setting):
    g = golden("beam")
    T, B, C, seed, period = [int(v) for v in g[case + "_shape"]]
    tseed, pen, bonus = g["%s_%s_cfg" % (case, setting)]
    x = synth.beam_logits(T, B, C, seed, period)
    c = _codec(C)
    c.set_beam_search(use_tfm_pred=False, lm_panelty=float(pen), len_bonus=float(bonus))
    c.lm_table = None if tseed < 0 else synth.lm_table(C, int(tseed))
    assert c.decode(x) == list(g["%s_%s_text" % (case, setting)])


def test_beam_narrow_golden(golden):
    g = golden("beam")
    T, B, C, seed, period = [int(v) for v in g["narrow_shape"]]
    x = synth.beam_logits(T, B, C, seed, period)
    c = _codec(C)
    c.set_beam_search(use_tfm_pred=False, lm_panelty=2.0, len_bonus=1.5, beam_size=4, search_depth=6)
    c.lm_table = synth.lm_table(C, 33)
    assert c.decode(torch.from_numpy(x).cuda()) == list(g["narrow_text"])


def test_beam_vs_oracle_config5_sample():
    """BASELINE config 5 shape (T=512, C=7375) on a sample of sequences; oracle = C restatement pinned to the reference."""
    T, B, C = 512, 6, 7375
    x = synth.beam_logits(T, B, C, 0, 8)
    for bonus, tab in ((0.0, None), (5.8, None), (5.8, synth.lm_table(C, 9))):
        c = _codec(C)
        c.set_beam_search(use_tfm_pred=False, lm_panelty=2.0, len_bonus=bonus)
        c.lm_table = tab
        idx, ln = c.beam_search_indices(torch.from_numpy(x).cuda())
        oi, ol, st = oracle.beam_search(x, 10, 10, 2.0, bonus, tab)
        assert np.array_equal(ln.cpu().numpy(), ol)
        for b in range(B):
            assert np.array_equal(idx[b, :ol[b]].cpu().numpy(), oi[b, :ol[b]])


def test_beam_equals_greedy_on_peaky_input_without_length_bonus():
    T, B, C = 200, 4, 500
    x = synth.beam_logits(T, B, C, 3, 5)
    c = _codec(C)
    greedy = c.decode(x)
    c.set_beam_search(use_tfm_pred=False, lm_panelty=2.0, len_bonus=0.0)
    assert c.decode(x) == greedy


def test_beam_empty_greedy_path_raises_index_error():
    x = np.zeros((5, 2, 9), np.float32)
    x[:, :, 0] = 10.0
    c = _codec(9)
    c.set_beam_search(use_tfm_pred=False, search_depth=5)
    with pytest.raises(IndexError):
        c.decode(x)


# ------------------------------------------------------------------ __cbs_skip__ (utils/ctc_codec.py:124-181)
def _skip_inputs(g, case):
    T, B, C, seed, period = [int(v) for v in g[case + "_shape"]]
    noise, boost = float(g[case + "_noise"]), float(g[case + "_boost"])
    x = synth.beam_logits(T, B, C, seed, period)
    if noise != 2.0:
        x = (x * (noise / 2.0)).astype(np.float32)
    return synth.peakier(x, boost), C


# "big": up to 144 classes above the prune threshold in one step -> the call is repeated with the 1024-candidate tables
SKIP_CASES = [(c, s) for c in ("small", "mid", "wide", "flat", "big") for s in ("zero_b0", "zero_b58", "tab_p2")]


@pytest.mark.parametrize("case,setting", SKIP_CASES)
def test_skip_search_golden(golden, case, setting):
    g = golden("beam_skip")
    x, C = _skip_inputs(g, case)
    tseed, pen, bonus = g["%s_%s_cfg" % (case, setting)]
    c = _codec(C)
    c.set_beam_search(skip_search=True, use_tfm_pred=False, lm_panelty=float(pen), len_bonus=float(bonus))
    c.lm_table = None if tseed < 0 else synth.lm_table(C, int(tseed))
    assert c.decode(x) == list(g["%s_%s_text" % (case, setting)])


def test_skip_search_many_candidates_and_errors(golden):
    """A nearly flat distribution: several hundred classes above the 0.001 prune threshold at every step (the reference takes
    up to 999, utils/ctc_codec.py:144). Device (1024-candidate variant) vs the oracle restatement of __cbs_skip__."""
    from oracle.codec import CodecTables
    T, B, C = 24, 2, 900
    rs = np.random.RandomState(12)
    x = (0.3 * rs.randn(T, B, C)).astype(np.float32)       # p ~ 1/900 each: about half the classes pass the threshold
    x[np.arange(0, T, 3), :, 7] += 6.0                     # a non-empty greedy path
    x[np.arange(1, T, 3), :, 0] += 6.0
    lp = oracle.log_softmax(x)
    counts = (lp > np.log(0.001)).sum(axis=2)
    assert counts.max() > 300 and counts.min() >= 1
    c = _codec(C)
    c.set_beam_search(skip_search=True, use_tfm_pred=False, lm_panelty=2.0, len_bonus=0.0)
    idx, ln, st = oracle.beam_search_skip(x, 10, 2.0, 0.0, None)
    assert (st == 0).all()
    assert c.decode(x) == CodecTables(synth.charset(C - 2)).to_text(idx, ln)
    flat = np.zeros((6, 1, 2000), np.float32)              # uniform over 2000 classes: nothing above 0.001 -> reference IndexError
    flat[:, 0, 5] = 0.5
    c2 = _codec(2000)
    c2.set_beam_search(skip_search=True, use_tfm_pred=False)
    with pytest.raises(IndexError):
        c2.decode(flat)


def test_edit_distance_and_cer_on_device():
    """CER of decoded label arrays vs ground-truth strings (main.py:506-517) - bit-exact integer distances."""
    T, B, C = 700, 9, 300
    rs = np.random.RandomState(3)
    c = _codec(C)
    x = synth.ctc_like_logits(T, B, C, 31, period=2)
    x[:, 3, :] = 0.0; x[:, 3, 0] = 5.0                      # an empty prediction
    idx, ln = c.greedy_indices(torch.from_numpy(x).cuda())
    preds = c.indices_to_text(idx, ln)
    truths = []
    for b in range(B):
        s = list(preds[b])
        for _ in range(rs.randint(0, 12)):                  # random edits
            op = rs.randint(3)
            pos = rs.randint(0, len(s) + 1)
            if op == 0: s.insert(pos, c.characters[1 + rs.randint(C - 2)])
            elif op == 1 and s: s.pop(min(pos, len(s) - 1))
            elif s: s[min(pos, len(s) - 1)] = c.characters[1 + rs.randint(C - 2)]
        truths.append(''.join(s))
    truths[5] = ""                                           # an empty ground truth
    truths[6] = truths[6] + "éx"                        # characters outside the charset
    dist, nchars = c.error_counts(idx, ln, truths)
    want = [oracle.edit_distance(p, t) for p, t in zip(preds, truths)]
    assert dist.cpu().tolist() == want
    assert nchars == sum(len(t) for t in truths)


# ---- back-off n-gram language model inside the beam search (reference: kenlm, utils/ctc_codec.py:120-122,276-279) ------
def _ngram_codec(C, order, lmseed):
    from hctr_b200.ngram_lm import NgramLM
    c = _codec(C)
    text = synth.arpa_text(synth.charset(C - 2)[:200], order, lmseed, grams_per_order=500)
    c.set_beam_search(use_tfm_pred=False)
    c.ngram = NgramLM.from_arpa_text(text, c.dict, C)
    return c, text


def test_ngram_score_on_device_matches_oracle():
    """hctr_ngram_score == kenlm score(bos=True, eos=False) as restated in oracle/ngram.py, bit for bit (float32)."""
    from oracle.ngram import ArpaLM
    C = 300
    c, text = _ngram_codec(C, 5, 17)
    o = ArpaLM(text)
    rs = np.random.RandomState(5)
    seqs = [rs.randint(1, C - 1, size=rs.randint(0, 30)).tolist() for _ in range(200)]
    seqs += [rs.randint(1, 201, size=rs.randint(1, 40)).tolist() for _ in range(200)]      # in-vocabulary heavy
    got = c.ngram.score_ids(seqs)
    want = np.array([o.score(" ".join(c.characters[i] for i in s)) for s in seqs], np.float32)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("case", ["tri", "five", "five_wide"])
@pytest.mark.parametrize("setting,pen,bonus", [("p2_b58", 2.0, 5.8), ("p1_b2", 1.0, 2.0)])
def test_beam_ngram_golden(golden, case, setting, pen, bonus):
    """The reference's own beam search driven by the ARPA scorer vs the device search with the n-gram table."""
    g = golden("beam_ngram")
    T, B, C, seed, period, order, lmseed = [int(v) for v in g[case + "_shape"]]
    x = synth.beam_logits(T, B, C, seed, period)
    c, _ = _ngram_codec(C, order, lmseed)
    c.lm_panelty, c.len_bonus = pen, bonus
    assert c.decode(torch.from_numpy(x).cuda()) == list(g["%s_%s_text" % (case, setting)])


def test_set_beam_search_reads_an_arpa_file(tmp_path):
    C = 120
    p = tmp_path / "lm.arpa"
    p.write_text(synth.arpa_text(synth.charset(C - 2)[:60], 3, 4, grams_per_order=100), encoding="utf-8")
    c = _codec(C)
    c.set_beam_search(ngram_path=str(p), use_tfm_pred=False)
    assert c.ngram is not None and c.ngram.order == 3 and c.lm_table is None
    x = synth.beam_logits(40, 2, C, 8, 5)
    out = c.decode(torch.from_numpy(x).cuda())
    assert len(out) == 2 and all(isinstance(t, str) for t in out)
    with pytest.raises(NotImplementedError):
        c.set_beam_search(ngram_path="model.bin", use_tfm_pred=False)
    # a KenLM binary next to the ARPA file it was built from (third-party/README.md:40-42): the ARPA file is read instead
    (tmp_path / "lm.bin").write_bytes(b"mmap lm http://kheafield.com/code format version 5\n\0")
    c2 = _codec(C)
    c2.set_beam_search(ngram_path=str(tmp_path / "lm.bin"), use_tfm_pred=False)
    assert c2.ngram is not None and c2.ngram.order == 3
    assert c2.decode(torch.from_numpy(x).cuda()) == out


@pytest.mark.parametrize("case", ["small", "mid", "flat"])
def test_skip_search_ngram_golden(golden, case):
    """__cbs_skip__ of the reference driven by the ARPA scorer vs the device skip search with the n-gram table."""
    g = golden("beam_skip_ngram")
    T, B, C, seed, period, order, lmseed = [int(v) for v in g[case + "_shape"]]
    noise, boost = float(g[case + "_noise"]), float(g[case + "_boost"])
    x = synth.beam_logits(T, B, C, seed, period)
    if noise != 2.0:
        x = (x * (noise / 2.0)).astype(np.float32)
    x = synth.peakier(x, boost)
    c, _ = _ngram_codec(C, order, lmseed)
    c.skip_search = True
    c.lm_panelty, c.len_bonus = 2.0, 5.8
    assert c.decode(torch.from_numpy(x).cuda()) == list(g[case + "_text"])
