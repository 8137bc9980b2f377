"""Pin oracle/ngram.py (and through it the device scorer) to the REAL KenLM library - to be run where `kenlm` can be
installed (the reference installs it from GitHub master, third-party/README.md:23-26; there is no network in the build
container, so tests/golden/kenlm_scores.npz is absent until someone runs this):

    pip install https://github.com/kpu/kenlm/archive/master.zip
    python tests/golden/make_kenlm_golden.py            # writes tests/golden/kenlm_scores.npz

It writes the ARPA text of three synthetic models (orders 2, 3, 5; the generator is tests/synth.arpa_text, the same one the
GPU tests use), a few hundred sentences over their vocabularies (incl. out-of-vocabulary words), and
`kenlm.Model(path).score(sentence, bos=True, eos=False)` for each - the call of utils/ctc_codec.py:279 - both from the ARPA
file and from the binary `build_binary` makes of it when that tool is on PATH (third-party/README.md:40-42).
tests/test_oracle_golden.py::test_ngram_oracle_matches_real_kenlm replays them (skipped while the file is absent)."""
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import synth  # noqa: E402


def sentences(chars, n, seed):
    rs = np.random.RandomState(seed)
    out = []
    for _ in range(n):
        L = int(rs.randint(1, 12))
        words = [chars[int(rs.randint(len(chars)))] for _ in range(L)]
        if rs.rand() < 0.2:
            words[int(rs.randint(L))] = "☃"                    # a word outside the vocabulary -> <unk>
        out.append(" ".join(words))
    return out


def main():
    import kenlm
    out = {}
    chars = synth.charset(200)
    for name, order, seed in (("bi", 2, 21), ("tri", 3, 22), ("five", 5, 23)):
        text = synth.arpa_text(chars[:120], order, seed, grams_per_order=600)
        sents = sentences(chars[:150], 300, seed + 100)
        with tempfile.TemporaryDirectory() as tmp:
            arpa = os.path.join(tmp, name + ".arpa")
            with open(arpa, "w", encoding="utf-8") as fh:
                fh.write(text)
            model = kenlm.Model(arpa)
            scores = np.array([model.score(s, bos=True, eos=False) for s in sents], dtype=np.float64)
            out[name + "_arpa"] = np.array([text], dtype=object)
            out[name + "_sentences"] = np.array(sents, dtype=object)
            out[name + "_scores"] = scores
            tool = shutil.which("build_binary")
            if tool:
                binary = os.path.join(tmp, name + ".bin")
                subprocess.run([tool, arpa, binary], check=True, capture_output=True)
                mb = kenlm.Model(binary)
                out[name + "_scores_binary"] = np.array([mb.score(s, bos=True, eos=False) for s in sents], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "kenlm_scores.npz"), **out)
    print("wrote kenlm_scores.npz:", sorted(out))


if __name__ == "__main__":
    main()
