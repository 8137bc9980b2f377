"""Seeded synthetic inputs shared by tests/golden/make_golden.py, the tests and bench.py.
Everything uses numpy's legacy RandomState (bit-stable across numpy versions), so fixtures only need to
store the reference's OUTPUTS; the inputs are regenerated from the seed on any box."""
import numpy as np


def charset(n):
    """Synthetic charset of n distinct CJK characters (BASELINE config 1: chr(0x4E00+i))."""
    return ''.join(chr(0x4E00 + i) for i in range(n))


def ctc_like_logits(T, B, C, seed, peak=12.0, period=3, noise=2.0):
    """Peaky CTC-like logits [T,B,C] fp32: a planted path that emits a random class every `period` steps
    (blank otherwise) on top of noise*randn; includes forced repeats and unknown-class hits."""
    rs = np.random.RandomState(seed)
    x = (noise * rs.randn(T, B, C)).astype(np.float32)
    for b in range(B):
        last = 0
        for t in range(T):
            r = rs.rand()
            if t % period == 0:
                if r < 0.15 and last not in (0,):
                    c = last                          # forced repeat of the previous emission
                elif r < 0.25:
                    c = C - 1                         # unknown-class hit
                else:
                    c = 1 + rs.randint(C - 2)
                last = c
            elif r < 0.2 and last != 0:
                c = last                              # run of the same class (collapsed by the decoder)
            else:
                c = 0
            x[t, b, c] += peak
    return x


def beam_logits(T, B, C, seed, period=8):
    """BASELINE config 5 input: 2*randn with +12 on a random class at t%period==0 and on blank otherwise
    (guarantees a non-empty greedy path and no ties inside the top 11)."""
    rs = np.random.RandomState(seed)
    x = (2.0 * rs.randn(T, B, C)).astype(np.float32)
    for b in range(B):
        for t in range(T):
            c = 1 + rs.randint(C - 2) if t % period == 0 else 0
            x[t, b, c] += 12.0
    return x


def lm_table(C, seed):
    """Deterministic per-class unigram 'log10 probabilities' for the table LM stub."""
    rs = np.random.RandomState(seed)
    t = -(0.5 + 3.0 * rs.rand(C))
    t[0] = 0.0
    return t.astype(np.float64)


def ctc_targets(B, C, Lmin, Lmax, seed, repeat_frac=0.1):
    """Random label sequences over classes 1..C-2 with forced repeats -> (targets int32 [sum L], lengths int32 [B])."""
    rs = np.random.RandomState(seed)
    lens = rs.randint(Lmin, Lmax + 1, size=B).astype(np.int32)
    out = []
    for L in lens:
        seq = []
        for i in range(L):
            if i > 0 and rs.rand() < repeat_frac:
                seq.append(seq[-1])
            else:
                seq.append(1 + rs.randint(C - 2))
        out.extend(seq)
    return np.array(out, dtype=np.int32), lens


def text_lines(B, W, seed, H=128):
    """Synthetic 'ink on paper' text lines [B,1,H,W] fp32 in [-1,1]: background +1, random dark strokes."""
    rs = np.random.RandomState(seed)
    x = np.ones((B, 1, H, W), np.float32)
    for b in range(B):
        n = max(4, W // 6)
        cx = rs.randint(0, W, size=n); cy = rs.randint(16, H - 16, size=n)
        ln = rs.randint(4, 28, size=n); th = rs.randint(1, 4, size=n); hor = rs.rand(n) < 0.5
        val = -1.0 + 0.6 * rs.rand(n)
        for i in range(n):
            if hor[i]:
                x[b, 0, cy[i]:cy[i] + th[i], max(0, cx[i] - ln[i]):cx[i] + ln[i]] = val[i]
            else:
                x[b, 0, max(0, cy[i] - ln[i]):cy[i] + ln[i], cx[i]:cx[i] + th[i]] = val[i]
    x += (0.05 * rs.randn(B, 1, H, W)).astype(np.float32)
    return np.clip(x, -1.0, 1.0).astype(np.float32)


def peakier(x, boost):
    """Add `boost` to the arg-max class of every (t, b) row of logits [T,B,C] (returns a copy)."""
    y = np.array(x, copy=True)
    if boost:
        am = y.argmax(axis=2)
        t, b = np.meshgrid(np.arange(y.shape[0]), np.arange(y.shape[1]), indexing="ij")
        y[t, b, am] += np.float32(boost)
    return y


def arpa_text(chars, order, seed, vocab_frac=0.8, grams_per_order=400):
    """A small, structurally valid ARPA model over single-character words: every n-gram's prefix and suffix exist,
    random log10 probabilities / back-offs with at most 6 significant digits (what lmplz prints), <unk>, <s>, </s>."""
    rs = np.random.RandomState(seed)
    vocab = [c for c in chars if rs.rand() < vocab_frac]
    words = ["<unk>", "<s>", "</s>"] + vocab
    levels = [set((w,) for w in words)]
    for n in range(2, order + 1):
        prev = sorted(levels[-1])
        cur = set()
        tries = 0
        while len(cur) < grams_per_order and tries < 20 * grams_per_order:
            tries += 1
            g = prev[rs.randint(len(prev))]
            if g[-1] == "</s>":
                continue
            w = words[2 + rs.randint(len(words) - 2)]                  # never <unk> or <s> as the predicted word
            cand = g + (w,)
            if cand[1:] in levels[-1] and "<unk>" not in cand:         # the suffix must exist too (KenLM requires it)
                cur.add(cand)
        levels.append(cur)
    lines = ["\\data\\"] + ["ngram %d=%d" % (n + 1, len(l)) for n, l in enumerate(levels)] + [""]
    for n, level in enumerate(levels, start=1):
        lines.append("\\%d-grams:" % n)
        for g in sorted(level):
            p = -round(0.2 + 4.0 * rs.rand(), 5)
            if g == ("<s>",):
                p = -99.0
            if n < order and g[-1] != "</s>":
                lines.append("%g\t%s\t%g" % (p, " ".join(g), -round(1.5 * rs.rand(), 5)))
            else:
                lines.append("%g\t%s" % (p, " ".join(g)))
        lines.append("")
    lines.append("\\end\\")
    return "\n".join(lines) + "\n"


# cv2.resize(INTER_AREA) parity cases (tests/golden/resize.npz): (source height, source width, seed) - shrinking (area-weighted
# mean), enlarging (bilinear with INTER_AREA coefficients), integer factors (fast path), mixed factors, tiny sources
RESIZE_CASES = [
    (359, 500, 1), (250, 371, 2), (183, 378, 3), (130, 696, 4), (347, 179, 5),
    (110, 272, 6), (57, 427, 7), (30, 634, 8), (99, 635, 9),
    (384, 300, 10), (512, 512, 11), (640, 333, 12), (128, 300, 13), (256, 512, 14),
    (127, 100, 15), (129, 100, 16), (300, 50, 17), (9, 40, 18),
]


def resize_source(sh, sw, seed):
    """Random uint8 line of the given size (legacy RandomState, as the fixture generator uses)."""
    return np.random.RandomState(1000 + seed).randint(0, 256, size=(sh, sw)).astype(np.uint8)
