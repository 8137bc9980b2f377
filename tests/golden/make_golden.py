"""Generate tests/golden/*.npz by running the REFERENCE itself (imported from /root/reference) in the build
container. The reference cannot travel to the GPU box, so its outputs are committed as small fixtures and this
script is the record of how they were made:   python tests/golden/make_golden.py
Inputs are regenerated from seeds by tests/synth.py; fixtures hold the reference's outputs."""
import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")
sys.path.insert(0, os.path.join(ROOT, "tests"))
from models.handwritten_ctr_model import hctr_model          # noqa: E402  (reference)
from utils.ctc_codec import ctc_codec                        # noqa: E402  (reference)
from utils.dataset import NormalizePAD                       # noqa: E402  (reference)
import synth                                                  # noqa: E402

torch.set_num_threads(8)


def strs(lst):
    return np.array(lst, dtype=object)


# ------------------------------------------------------------------ codec: encode + greedy
def make_greedy():
    out = {}
    cases = [("small", 40, 3, 30, 11), ("mid", 64, 4, 200, 12), ("wide", 6, 2, 7375, 13), ("t1", 1, 2, 30, 14)]
    for name, T, B, C, seed in cases:
        codec = ctc_codec(synth.charset(C - 2))
        x = synth.ctc_like_logits(T, B, C, seed)
        out[name + "_shape"] = np.array([T, B, C, seed])
        out[name + "_text"] = strs(codec.decode(x))
    # exact ties -> lowest index; all-blank sequence; NaN -> first NaN (numpy.argmax semantics)
    C = 12
    codec = ctc_codec(synth.charset(C - 2))
    x = np.zeros((6, 3, C), np.float32)
    x[0, 0, [3, 5]] = 2.0; x[1, 0, [5, 3]] = 2.0; x[2, 0, 0] = 1.0; x[3, 0, [7, 8, 9]] = 4.0; x[4, 0, C - 1] = 3.0; x[5, 0, 7] = 1.0
    x[:, 1, 0] = 5.0
    x[0, 2, 4] = np.nan; x[1, 2, 6] = 9.0; x[2, 2, [2, 9]] = np.nan; x[3, 2, 9] = 1.0; x[4, 2, 9] = 1.0; x[5, 2, 1] = -1.0
    out["edge_logits"] = x
    out["edge_text"] = strs(codec.decode(x))
    # encode: known, unknown, empty strings, duplicate char in the charset
    chars = synth.charset(20) + synth.charset(3)[1]
    codec = ctc_codec(chars)
    texts = [synth.charset(20)[3:9], "", "x" + synth.charset(20)[0] + "?", synth.charset(3)[1] * 2]
    idx, ln = codec.encode(texts)
    out["enc_chars"] = strs([chars]); out["enc_texts"] = strs(texts); out["enc_idx"] = idx; out["enc_len"] = ln
    np.savez_compressed(os.path.join(HERE, "greedy.npz"), **out)


# ------------------------------------------------------------------ beam search (__cbs_full__)
class TableLM(object):
    def __init__(self, table, dict_):
        self.table, self.dict = table, dict_

    def score(self, sentence, eos=False):
        return sum(self.table[self.dict[ch]] for ch in sentence.split(' ') if ch)


class ZeroLM(object):
    def score(self, sentence, eos=False):
        return 0.0


def make_beam():
    out = {}
    cases = [("small", 48, 3, 40, 21, 4), ("mid", 96, 2, 300, 22, 8), ("wide", 24, 1, 7375, 23, 6)]
    settings = [("zero_b0", None, 2.0, 0.0), ("zero_b58", None, 2.0, 5.8), ("tab_p2", 31, 2.0, 5.8), ("tab_p08", 32, 0.8, 4.8)]
    for name, T, B, C, seed, period in cases:
        x = synth.beam_logits(T, B, C, seed, period)
        out[name + "_shape"] = np.array([T, B, C, seed, period])
        for sname, tseed, pen, bonus in settings:
            codec = ctc_codec(synth.charset(C - 2))
            codec.use_beam_search = True; codec.use_tfm_pred = False; codec.use_tfm_score = False
            codec.skip_search = False; codec.lm_panelty = pen; codec.len_bonus = bonus
            codec.ngram = ZeroLM() if tseed is None else TableLM(synth.lm_table(C, tseed), codec.dict)
            out["%s_%s_text" % (name, sname)] = strs(codec.decode(x))
            out["%s_%s_cfg" % (name, sname)] = np.array([-1 if tseed is None else tseed, pen, bonus])
    # smaller beam / depth
    T, B, C, seed, period = 40, 2, 60, 24, 5
    x = synth.beam_logits(T, B, C, seed, period)
    codec = ctc_codec(synth.charset(C - 2))
    codec.use_beam_search = True; codec.use_tfm_pred = False; codec.use_tfm_score = False; codec.skip_search = False
    codec.beam_size = 4; codec.search_depth = 6; codec.lm_panelty = 2.0; codec.len_bonus = 1.5
    codec.ngram = TableLM(synth.lm_table(C, 33), codec.dict)
    out["narrow_shape"] = np.array([T, B, C, seed, period]); out["narrow_text"] = strs(codec.decode(x))
    np.savez_compressed(os.path.join(HERE, "beam.npz"), **out)


# ------------------------------------------------------------------ beam search with a back-off n-gram LM
class ArpaStub(object):
    """What `kenlm.Model` is to the reference's beam search (utils/ctc_codec.py:120-122,279): .score(sentence, eos=False).
    kenlm itself is not installed here; the scorer is oracle/ngram.py (KenLM's published query algorithm, float32)."""

    def __init__(self, text):
        sys.path.insert(0, ROOT)
        from oracle.ngram import ArpaLM
        self.lm = ArpaLM(text)

    def score(self, sentence, eos=False):
        return self.lm.score(sentence, bos=True, eos=eos)


def make_beam_ngram():
    out = {}
    cases = [("tri", 64, 3, 60, 41, 6, 3, 51), ("five", 80, 2, 120, 42, 7, 5, 52), ("five_wide", 32, 1, 7375, 43, 6, 5, 53)]
    for name, T, B, C, seed, period, order, lmseed in cases:
        x = synth.beam_logits(T, B, C, seed, period)
        chars = synth.charset(C - 2)
        arpa = synth.arpa_text(chars[:200], order, lmseed, grams_per_order=500)
        for sname, pen, bonus in (("p2_b58", 2.0, 5.8), ("p1_b2", 1.0, 2.0)):
            codec = ctc_codec(chars)
            codec.use_beam_search = True; codec.use_tfm_pred = False; codec.use_tfm_score = False
            codec.skip_search = False; codec.lm_panelty = pen; codec.len_bonus = bonus
            codec.ngram = ArpaStub(arpa)
            out["%s_%s_text" % (name, sname)] = strs(codec.decode(x))
        out[name + "_shape"] = np.array([T, B, C, seed, period, order, lmseed])
    np.savez_compressed(os.path.join(HERE, "beam_ngram.npz"), **out)


def make_beam_skip():
    """__cbs_skip__ (utils/ctc_codec.py:124-181): pruned candidates + the single-candidate fast path."""
    out = {}
    # noise scales chosen so that steps mix the single-candidate fast path with 2..~60-candidate searches
    # (name, T, B, C, seed, period, scale, boost): boost sharpens the arg-max so that steps mix the single-candidate
    # fast path with small multi-candidate searches; "big" keeps >64 candidates per step (oracle only).
    cases = [("small", 48, 3, 40, 21, 4, 2.0, 0.0), ("mid", 96, 2, 300, 22, 8, 2.0, 2.0), ("wide", 64, 2, 7375, 23, 6, 2.0, 3.5),
             ("flat", 60, 2, 60, 25, 3, 0.7, 0.0), ("big", 40, 2, 7375, 26, 6, 2.0, 0.0)]
    settings = [("zero_b0", None, 2.0, 0.0), ("zero_b58", None, 2.0, 5.8), ("tab_p2", 31, 2.0, 5.8)]
    for name, T, B, C, seed, period, noise, boost in cases:
        x = synth.beam_logits(T, B, C, seed, period)
        if noise != 2.0:
            x = (x * (noise / 2.0)).astype(np.float32)          # weaker peaks -> more steps with several candidates
        x = synth.peakier(x, boost)
        out[name + "_shape"] = np.array([T, B, C, seed, period]); out[name + "_noise"] = np.array(noise)
        out[name + "_boost"] = np.array(boost)
        for sname, tseed, pen, bonus in settings:
            codec = ctc_codec(synth.charset(C - 2))
            codec.use_beam_search = True; codec.use_tfm_pred = False; codec.use_tfm_score = False
            codec.skip_search = True; codec.lm_panelty = pen; codec.len_bonus = bonus
            codec.ngram = ZeroLM() if tseed is None else TableLM(synth.lm_table(C, tseed), codec.dict)
            out["%s_%s_text" % (name, sname)] = strs(codec.decode(x))
            out["%s_%s_cfg" % (name, sname)] = np.array([-1 if tseed is None else tseed, pen, bonus])
    np.savez_compressed(os.path.join(HERE, "beam_skip.npz"), **out)


# ------------------------------------------------------------------ CTC loss (main.py:205,406-409)
def make_ctc_loss():
    out = {}
    crit = torch.nn.CTCLoss(zero_infinity=True)
    cases = [("small", 30, 3, 20, 41, 2, 8), ("mid", 80, 4, 500, 42, 5, 20), ("wide", 64, 2, 7375, 43, 10, 20),
             ("len1", 12, 2, 16, 44, 1, 1), ("infeasible", 6, 2, 16, 45, 5, 6)]
    for name, T, B, C, seed, Lmin, Lmax in cases:
        x = torch.from_numpy(synth.ctc_like_logits(T, B, C, seed, peak=4.0)).requires_grad_(True)
        tg, tl = synth.ctc_targets(B, C, Lmin, Lmax, seed + 100, repeat_frac=0.3)
        il = torch.IntTensor([T] * B)
        loss = crit(x.log_softmax(2), torch.from_numpy(tg), il, torch.from_numpy(tl))
        loss.backward()
        g = x.grad.numpy()
        out[name + "_shape"] = np.array([T, B, C, seed, Lmin, Lmax])
        out[name + "_loss"] = np.array(loss.item(), dtype=np.float64)
        if g.size <= 200000:
            out[name + "_grad"] = g
        else:   # sampled entries + per-(t,b) row sums of |grad|
            rs = np.random.RandomState(7)
            pick = rs.randint(0, g.size, size=4096)
            out[name + "_grad_pick"] = pick; out[name + "_grad_vals"] = g.reshape(-1)[pick]
            out[name + "_grad_rowabs"] = np.abs(g).sum(2)
            # and the gradient at the target classes (where the alpha/beta term lives)
            out[name + "_grad_tgt"] = g[:, :, np.unique(tg)]
            out[name + "_grad_tgt_cls"] = np.unique(tg)
    np.savez_compressed(os.path.join(HERE, "ctc_loss.npz"), **out)


# ------------------------------------------------------------------ model forward
def state_hash(sd):
    h = hashlib.sha256()
    for k, v in sd.items():
        h.update(k.encode()); h.update(v.numpy().tobytes())
    return h.hexdigest()


def make_model():
    out = {}
    torch.manual_seed(1234)
    m = hctr_model(7375)
    out["sd_hash_7375_seed1234"] = strs([state_hash(m.state_dict())])
    out["sd_keys"] = strs(list(m.state_dict().keys()))
    out["sd_shapes"] = strs([str(tuple(v.shape)) for v in m.state_dict().values()])
    # small-class model (tiny logits fixture), default init and BN-calibrated
    NC, B, W = 37, 2, 72
    torch.manual_seed(4321)
    m = hctr_model(NC).eval()
    out["small_sd_hash"] = strs([state_hash(m.state_dict())])
    x = torch.from_numpy(synth.text_lines(B, W, 51))
    with torch.no_grad():
        out["small_default_logits"] = m(x).numpy()
    for mod in m.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.momentum = 1.0
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    m.train()
    with torch.no_grad():
        m(torch.from_numpy(synth.text_lines(3, 96, 52)))
    m.eval()
    sd = m.state_dict()
    for k, v in sd.items():
        if k.endswith("running_mean") or k.endswith("running_var"):
            out["small_cal." + k] = v.numpy()
    with torch.no_grad():
        out["small_cal_logits"] = m(x).numpy()
    # full-charset model, default init, one short line: greedy text only (config-1 style check)
    torch.manual_seed(1234)
    m = hctr_model(7375).eval()
    codec = ctc_codec(synth.charset(7373))
    x = torch.from_numpy(synth.text_lines(1, 136, 53))
    with torch.no_grad():
        y = m(x).numpy()
    out["full_default_text"] = strs(codec.decode(y))
    out["full_default_logits_t0"] = y[0, 0, :]
    out["full_default_absmax"] = np.array(np.abs(y).max())
    np.savez_compressed(os.path.join(HERE, "model.npz"), **out)


# ------------------------------------------------------------------ BASELINE config 1: the 5 bundled images
def make_config1():
    import cv2
    out = {}
    names = ["000000.jpg", "000001.jpg", "000002.jpg", "000003.jpg", "000004.jpg"]
    torch.manual_seed(1234)
    m = hctr_model(7375).eval()
    codec = ctc_codec(synth.charset(7373))
    texts, widths = [], []
    for i, n in enumerate(names):
        src = cv2.imread(os.path.join("/root/reference/images", n))        # test.py:207-215 preprocess_input
        src = cv2.cvtColor(src, cv2.COLOR_BGR2GRAY)
        tw = int(128 * float(src.shape[1]) / float(src.shape[0]))
        rsz = cv2.resize(src, (tw, 128), fx=0, fy=0, interpolation=cv2.INTER_AREA)
        out["img%d" % i] = rsz                                             # uint8 [128, W] after preprocessing
        widths.append(tw)
        t = NormalizePAD((1, 128, tw))(rsz[:, :, None]).unsqueeze(0)       # test.py:181-186
        with torch.no_grad():
            texts.append(codec.decode(m(t).numpy())[0])                    # test.py:191-194, -b 1
    out["widths"] = np.array(widths); out["text"] = strs(texts)
    np.savez_compressed(os.path.join(HERE, "config1.npz"), **out)


def make_beam_skip_ngram():
    """__cbs_skip__ with the ARPA scorer: fast-path steps keep the LM total only implicitly (the prefix is re-scored)."""
    out = {}
    cases = [("small", 48, 3, 40, 21, 4, 2.0, 0.0, 3, 61), ("mid", 96, 2, 300, 22, 8, 2.0, 2.0, 5, 62), ("flat", 60, 2, 60, 25, 3, 0.7, 0.0, 4, 63)]
    for name, T, B, C, seed, period, noise, boost, order, lmseed in cases:
        x = synth.beam_logits(T, B, C, seed, period)
        if noise != 2.0:
            x = (x * (noise / 2.0)).astype(np.float32)
        x = synth.peakier(x, boost)
        chars = synth.charset(C - 2)
        arpa = synth.arpa_text(chars[:200], order, lmseed, grams_per_order=500)
        codec = ctc_codec(chars)
        codec.use_beam_search = True; codec.use_tfm_pred = False; codec.use_tfm_score = False
        codec.skip_search = True; codec.lm_panelty = 2.0; codec.len_bonus = 5.8
        codec.ngram = ArpaStub(arpa)
        out[name + "_shape"] = np.array([T, B, C, seed, period, order, lmseed]); out[name + "_noise"] = np.array(noise)
        out[name + "_boost"] = np.array(boost)
        out[name + "_text"] = strs(codec.decode(x))
    np.savez_compressed(os.path.join(HERE, "beam_skip_ngram.npz"), **out)


# ------------------------------------------------------------------ NormalizePAD (utils/dataset.py:78-93)
def make_pad():
    out = {}
    rs = np.random.RandomState(61)
    widths = [37, 64, 50]
    maxw = 64
    trans = NormalizePAD((1, 16, maxw))
    for i, w in enumerate(widths):
        img = rs.randint(0, 256, size=(16, w, 1)).astype(np.uint8)
        out["img%d" % i] = img[:, :, 0]
        out["pad%d" % i] = trans(img).numpy()
    np.savez_compressed(os.path.join(HERE, "pad.npz"), **out)


def make_resize():
    """cv2.resize(..., INTER_AREA) to height 128 exactly as the reference calls it (utils/dataset.py:53-57, test.py:206-214):
    random lines of the listed sizes plus the five bundled images (sources stored: the GPU box has no /root/reference)."""
    import cv2
    out = {"cv2_version": np.array(cv2.__version__)}
    for i, (sh, sw, seed) in enumerate(synth.RESIZE_CASES):
        src = synth.resize_source(sh, sw, seed)
        ratio = 128 / sh
        new_w = int(sw * ratio)                                            # utils/dataset.py:54-55
        out["dst%d" % i] = cv2.resize(src, (new_w, 128), interpolation=cv2.INTER_AREA)
    names = sorted(f for f in os.listdir("/root/reference/images") if f.startswith("0000") and f.endswith(".jpg"))
    for i, f in enumerate(names):
        src = cv2.imread(os.path.join("/root/reference/images", f))
        src = cv2.cvtColor(src, cv2.COLOR_BGR2GRAY)                        # test.py:207-208
        ratio = float(src.shape[1]) / float(src.shape[0])
        tw = int(128 * ratio)                                              # test.py:210-212
        out["img_src%d" % i] = src
        out["img_dst%d" % i] = cv2.resize(src, (tw, 128), fx=0, fy=0, interpolation=cv2.INTER_AREA)
    np.savez_compressed(os.path.join(HERE, "resize.npz"), **out)


if __name__ == "__main__":
    which = sys.argv[1:] or ["greedy", "beam", "beam_ngram", "beam_skip", "beam_skip_ngram", "ctc_loss", "model", "config1", "pad",
                             "resize"]
    for w in which:
        print("making", w, flush=True)
        globals()["make_" + w]()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))
