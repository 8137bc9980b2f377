"""TEST INFRASTRUCTURE ONLY - CPU restatement of the n-gram language model the reference's beam search scores with.

The reference calls `kenlm.Model(ngram_path).score(sentence, eos=False)` (utils/ctc_codec.py:120-122,276-279; the model
is a 5-gram trained with `lmplz -o 5`, third-party/README.md:28-42). `kenlm` (github.com/kpu/kenlm, unpinned "master" in
the reference) is not installed in this image, so this file restates its PUBLISHED query algorithm - PARITY UNPINNED
against the real library; it is pinned only against hand-computed ARPA back-off values (tests/test_oracle_golden.py):

  * ARPA file: sections `\\N-grams:` with lines `log10(p) <TAB> w1 .. wN [<TAB> log10(backoff)]`; values are stored as
    float32 (KenLM's ProbBackoff); a missing back-off is 0.
  * score(sentence, bos=True, eos=False): state = [<s>]; for each space-separated word (unknown words map to <unk>)
      p(w | ctx) = prob(ctx[-j:] + w) for the longest j such that the n-gram exists, plus the back-off weights of the
      longer contexts ctx[-i:], i = j+1 .. len(ctx), added in increasing i, all in float32 (lm/model.cc, FullScore);
    the running total is a C `float` (python/kenlm.pyx: `cdef float total`), returned as a Python float.
"""
import numpy as np


class ArpaLM(object):
    def __init__(self, text):
        self.order = 0
        self.grams = {}                       # tuple(words) -> (np.float32 prob, np.float32 backoff)
        section = 0
        for raw in text.splitlines():
            line = raw.strip()
            if not line:
                continue
            if line.startswith("\\"):
                if line.endswith("-grams:"):
                    section = int(line[1:line.index("-")])
                    self.order = max(self.order, section)
                elif line == "\\end\\":
                    break
                else:
                    section = 0
                continue
            if section == 0:
                continue                      # the ngram counts of the \data\ header
            cols = line.split("\t")
            if len(cols) < 2:
                cols = line.split()
                words = tuple(cols[1:1 + section])
                rest = cols[1 + section:]
            else:
                words = tuple(cols[1].split(" "))
                rest = cols[2:]
            assert len(words) == section, line
            self.grams[words] = (np.float32(cols[0]), np.float32(rest[0]) if rest else np.float32(0.0))
        if ("<unk>",) not in self.grams:      # lm/vocab.cc: a model without <unk> gets one with probability 10^-100
            self.grams[("<unk>",)] = (np.float32(-100.0), np.float32(0.0))

    def word_score(self, ctx, w):
        """log10 p(w | ctx) in float32; ctx = previous words, oldest first (only the last order-1 matter)."""
        ctx = tuple(ctx)[-(self.order - 1):] if self.order > 1 else ()
        if (w,) not in self.grams:
            w = "<unk>"
        ctx = tuple(c if (c,) in self.grams else "<unk>" for c in ctx)
        j = len(ctx)
        while j > 0 and ctx[len(ctx) - j:] + (w,) not in self.grams:
            j -= 1
        p = self.grams[ctx[len(ctx) - j:] + (w,)][0]
        for i in range(j + 1, len(ctx) + 1):
            g = self.grams.get(ctx[len(ctx) - i:])
            if g is not None:
                p = np.float32(p + g[1])
        return np.float32(p)

    def score(self, sentence, bos=True, eos=False):
        words = [w for w in sentence.split(" ") if w] if isinstance(sentence, str) else list(sentence)
        ctx = ["<s>"] if bos else []
        total = np.float32(0.0)
        for w in words:
            total = np.float32(total + self.word_score(ctx, w))
            ctx.append(w)
        if eos:
            total = np.float32(total + self.word_score(ctx, "</s>"))
        return float(total)
