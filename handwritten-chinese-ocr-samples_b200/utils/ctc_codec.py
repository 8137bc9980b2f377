"""B200-native drop-in for the reference CTC codec (reference: utils/ctc_codec.py).

Same surface: `ctc_codec(characters_str)`, `.encode(list[str])`, `.decode(preds[T,B,C])`,
`.set_beam_search(...)`, public attributes `characters`, `dict`, `chars_list` and the beam knobs.
`decode` accepts what the reference's callers pass (a NumPy [T,B,C] array, main.py:495, test.py:194)
*and* a CUDA tensor, in which case the logits never leave the device: arg-max + blank/repeat collapse
(greedy) or log-softmax + top-k + prefix beam search run as sm_100a kernels and only the compact
[B, <=T] int32 label array is copied back; the host does the index -> character mapping.
There is no CPU fallback: a NumPy input is uploaded and decoded on the GPU as well.
"""
import importlib.util
import os
import sys

import numpy as np
import torch

_PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _core():
    pkg = sys.modules.get("hctr_b200")
    if pkg is None:
        spec = importlib.util.spec_from_file_location(
            "hctr_b200", os.path.join(_PKG_DIR, "__init__.py"), submodule_search_locations=[_PKG_DIR])
        pkg = importlib.util.module_from_spec(spec)
        sys.modules["hctr_b200"] = pkg
        spec.loader.exec_module(pkg)
    return pkg


class ctc_codec(object):
    """Convert between text-label and text-index (reference: utils/ctc_codec.py:14-41)."""

    def __init__(self, characters_str):
        self.chars_list = list(characters_str)
        # index 0 is the CTC blank, the last index stands for characters outside the charset
        self.dict = {ch: i + 1 for i, ch in enumerate(self.chars_list)}
        self.characters = ['<blank>'] + self.chars_list + ['<unknown>']
        self.dict['<blank>'] = 0
        self.dict['<unknown>'] = len(self.characters) - 1
        self._char_array = np.array(self.characters, dtype=object)

        self.ngram = None
        self.transformer = None
        self.lm_panelty = 2          # (sic) spelling kept: it is part of the reference's attribute surface
        self.len_bonus = 5.8
        self.search_depth = 10
        self.beam_size = 10
        self.use_tfm_score = False
        self.use_tfm_pred = True
        self.skip_search = False
        self.use_beam_search = False
        # device language model for the beam path: per-class unigram log10 scores (None = zero LM)
        self.lm_table = None
        self.ngram = None               # NgramLM (back-off n-gram on the device) set by set_beam_search(ngram_path='*.arpa')
        self.device = None           # CUDA device used when decode() is handed a NumPy array

    # ------------------------------------------------------------------ encode (reference :43-61)
    def encode(self, text):
        """list[str] -> (int32 [sum L] concatenated label indices, int32 [B] lengths)."""
        length = [len(s) for s in text]
        unknown = len(self.characters) - 1
        single = self.dict
        # `char in chars_list` in the reference == membership among single-character keys of the dict
        index = [single[ch] if (ch in single and len(ch) == 1) else unknown for ch in ''.join(text)]
        return (np.array(index, dtype=np.int32), np.array(length, dtype=np.int32))

    # ------------------------------------------------------------------ decode (reference :63-68)
    def decode(self, preds):
        if preds.shape[0] == 0 and not self.use_beam_search:
            return []               # reference: zero-length samples are skipped (utils/ctc_codec.py:85-86)
        logits = self._as_device_logits(preds)
        if self.use_beam_search:
            idx, ln = self.skip_search_indices(logits) if self.skip_search else self.beam_search_indices(logits)
        else:
            idx, ln = self.greedy_indices(logits)
        return self.indices_to_text(idx, ln)

    def indices_to_text(self, idx, ln):
        idx = idx.cpu().numpy() if isinstance(idx, torch.Tensor) else np.asarray(idx)
        ln = ln.cpu().numpy() if isinstance(ln, torch.Tensor) else np.asarray(ln)
        table = self._char_array
        return [''.join(table[idx[b, :ln[b]]]) for b in range(idx.shape[0])]

    def _as_device_logits(self, preds):
        if isinstance(preds, np.ndarray):
            if preds.ndim != 3:
                raise ValueError("decode expects [T,B,C] predictions, got shape %s" % (preds.shape,))
            dev = self.device if self.device is not None else torch.device("cuda", torch.cuda.current_device())
            t = torch.from_numpy(np.ascontiguousarray(preds))
            if t.dtype not in (torch.float32, torch.bfloat16):
                t = t.float()
            return t.to(dev, non_blocking=False)
        if not isinstance(preds, torch.Tensor) or preds.dim() != 3:
            raise ValueError("decode expects a [T,B,C] numpy array or torch tensor")
        if not preds.is_cuda:
            dev = self.device if self.device is not None else torch.device("cuda", torch.cuda.current_device())
            preds = preds.to(dev)
        if preds.dtype not in (torch.float32, torch.bfloat16):
            preds = preds.float()
        if preds.stride(2) != 1:
            preds = preds.contiguous()
        return preds.detach()

    @staticmethod
    def _dtype_code(nat, t):
        return nat.HCTR_F32 if t.dtype == torch.float32 else nat.HCTR_BF16

    # ------------------------------------------------------------------ greedy (reference :70-99)
    def greedy_indices(self, logits, return_argmax=False):
        """logits: CUDA [T,B,C] (any T/B strides, class stride 1) -> (int32 [B,T] labels, int32 [B] lengths)
        on the device; blank(0)/unknown(C-1)/repeat collapse done on the GPU."""
        nat = _core().native
        T, B, C = logits.shape
        if C != len(self.characters):
            raise ValueError("logits have %d classes but the codec has %d" % (C, len(self.characters)))
        dev = logits.device
        with torch.cuda.device(dev):
            raw = torch.empty((B, T), dtype=torch.int32, device=dev)
            idx = torch.zeros((B, T), dtype=torch.int32, device=dev)
            ln = torch.zeros((B,), dtype=torch.int32, device=dev)
            nat.check(nat.lib().hctr_ctc_greedy_decode(
                nat.ptr(logits), self._dtype_code(nat, logits), T, B, C, logits.stride(0), logits.stride(1),
                nat.ptr(raw), nat.ptr(idx), nat.ptr(ln), nat.stream_ptr()), "ctc_greedy_decode")
        return (idx, ln, raw) if return_argmax else (idx, ln)

    # ------------------------------------------------------------------ CER on device (main.py:497-517)
    def error_counts(self, idx, ln, truths):
        """Edit distances between decoded label arrays (device, from greedy_/beam_search_indices) and ground-truth strings.
        Returns (int32 [B] distances on the device, total characters). CER = dist.sum() / nchars, as main.py:506-517."""
        nat = _core().native
        tg, tl = self.encode(list(truths))
        dev = idx.device
        B, T = idx.shape
        with torch.cuda.device(dev):
            d_tg = torch.from_numpy(tg).to(dev) if tg.size else torch.zeros((1,), dtype=torch.int32, device=dev)
            d_tl = torch.from_numpy(tl).to(dev)
            dist = torch.empty((B,), dtype=torch.int32, device=dev)
            nat.check(nat.lib().hctr_edit_distance(nat.ptr(idx), nat.ptr(ln), T, T, nat.ptr(d_tg), nat.ptr(d_tl), B, nat.ptr(dist),
                                                   nat.stream_ptr()), "edit_distance")
        return dist, int(tl.sum())

    # ------------------------------------------------------------------ beam search (reference :101-122,183-285)
    def set_beam_search(self, skip_search=False, ngram_path='', tfm_path='',
                        lm_panelty=2, len_bonus=5.8, beam_size=10, search_depth=10,
                        use_tfm_score=False, use_tfm_pred=True,
                        use_openvino=False):
        self.use_beam_search = True
        self.lm_panelty = lm_panelty
        self.len_bonus = len_bonus
        self.beam_size = beam_size
        self.search_depth = search_depth
        self.use_tfm_pred = use_tfm_pred
        self.use_tfm_score = use_tfm_score
        self.skip_search = skip_search
        if use_tfm_pred or use_tfm_score:
            # the reference imports fairseq / openvino here (utils/ctc_codec.py:113-119); neither is part of
            # this path (out of scope) -> same exception type as a missing dependency
            raise ImportError("hctr_b200: transformer language models (fairseq/OpenVINO) are outside the "
                              "B200 hot path; pass use_tfm_pred=False, use_tfm_score=False")
        self.lm_table = None            # per-class unigram table (float64 log10 scores), or
        self.ngram = None               # NgramLM: back-off n-gram model scored inside the kernel; neither = zero LM
        if ngram_path:
            if ngram_path.endswith(".npy"):
                table = np.load(ngram_path).astype(np.float64)
                if table.shape != (len(self.characters),):
                    raise ValueError("unigram table must have shape (%d,)" % len(self.characters))
                self.lm_table = table
            elif ngram_path.endswith((".arpa", ".arpa.txt", ".lm")):
                # the reference loads this file with kenlm.Model (utils/ctc_codec.py:120-122); here the ARPA text is
                # turned into a device hash table and queried inside the beam-search kernel
                self.ngram = _core().ngram_lm.NgramLM.from_arpa(ngram_path, self)
            else:
                # build_binary's output hashes its n-grams (probing) or bit-packs them (trie) and cannot be turned back into
                # n-grams without the library; the ARPA file it was made FROM holds the same model (third-party/README.md:
                # `lmplz -o 5 <text >text.arpa` then `build_binary text.arpa text.bin`) - point ngram_path at that file
                arpa = os.path.splitext(ngram_path)[0] + ".arpa"
                if os.path.exists(arpa):
                    self.ngram = _core().ngram_lm.NgramLM.from_arpa(arpa, self)
                else:
                    raise NotImplementedError("hctr_b200: KenLM binary files are not read; pass the ARPA file build_binary was "
                                              "given (looked for %s) or a per-class unigram table (*.npy)" % arpa)

    def skip_search_indices(self, logits):
        """__cbs_skip__ on the device (reference: utils/ctc_codec.py:124-181)."""
        nat = _core().native
        lib = nat.lib()
        T, B, C = logits.shape
        if C != len(self.characters):
            raise ValueError("logits have %d classes but the codec has %d" % (C, len(self.characters)))
        dev = logits.device
        with torch.cuda.device(dev):
            idx = torch.zeros((B, T), dtype=torch.int32, device=dev)
            ln = torch.zeros((B,), dtype=torch.int32, device=dev)
            status = torch.zeros((B,), dtype=torch.int32, device=dev)
            table = None
            if self.lm_table is not None:
                table = torch.from_numpy(np.ascontiguousarray(self.lm_table, dtype=np.float64)).to(dev)
            ngram = getattr(self, "ngram", None)
            lm_ref = None
            if ngram is not None and hasattr(ngram, "struct"):
                import ctypes
                lm = ngram.struct(dev)
                lm_ref = ctypes.byref(lm)
                table = None

            def run(max_candidates):
                nb = lib.hctr_ctc_skip_workspace_bytes_ex(T, B, max_candidates)
                ws = torch.empty((max(nb, 8) + 256,), dtype=torch.uint8, device=dev)
                off = (-ws.data_ptr()) % 256
                nat.check(lib.hctr_ctc_skip_beam_search_ex(
                    nat.ptr(logits), self._dtype_code(nat, logits), T, B, C, logits.stride(0), logits.stride(1),
                    int(self.beam_size), float(self.lm_panelty), float(self.len_bonus), nat.ptr(table), lm_ref, max_candidates,
                    nat.ptr(idx), nat.ptr(ln), nat.ptr(status), nat.c_void_p(ws.data_ptr() + off), nb, nat.stream_ptr()),
                    "ctc_skip_beam_search")
                return status.cpu()

            st = run(128)
            if T > 0 and bool((st == nat.HCTR_ERR_UNSUPPORTED).any()):
                # a step with more than 128 classes above the 0.001 prune threshold (the reference takes up to 999,
                # utils/ctc_codec.py:144): repeat with the large candidate tables
                st = run(lib.hctr_ctc_skip_max_candidates())
            if T == 0 or bool((st == nat.HCTR_ERR_INDEX).any()) or bool((st == -1).any()):
                raise IndexError("list index out of range")     # reference: utils/ctc_codec.py:139,179
            if bool((st != 0).any()):
                raise RuntimeError("hctr_b200: skip search failed with status %s" % sorted(set(st.tolist())))
        return idx, ln

    def beam_search_indices(self, logits):
        """__cbs_full__ on the device: fused log-softmax + top-k, then one CTA per sequence."""
        nat = _core().native
        lib = nat.lib()
        T, B, C = logits.shape
        if C != len(self.characters):
            raise ValueError("logits have %d classes but the codec has %d" % (C, len(self.characters)))
        k, beam = int(self.search_depth), int(self.beam_size)
        dev = logits.device
        with torch.cuda.device(dev):
            st = nat.stream_ptr()
            tk_idx = torch.empty((T, B, k), dtype=torch.int32, device=dev)
            tk_lp = torch.empty((T, B, k), dtype=torch.float32, device=dev)
            lse = torch.empty((T, B), dtype=torch.float32, device=dev)
            nat.check(lib.hctr_ctc_topk_logsoftmax(
                nat.ptr(logits), self._dtype_code(nat, logits), T, B, C, logits.stride(0), logits.stride(1), k,
                nat.ptr(tk_idx), nat.ptr(tk_lp), nat.ptr(lse), st), "ctc_topk_logsoftmax")
            idx = torch.zeros((B, T), dtype=torch.int32, device=dev)
            ln = torch.zeros((B,), dtype=torch.int32, device=dev)
            status = torch.zeros((B,), dtype=torch.int32, device=dev)
            table = None
            if self.lm_table is not None:
                table = torch.from_numpy(np.ascontiguousarray(self.lm_table, dtype=np.float64)).to(dev)
            ws_bytes = lib.hctr_ctc_beam_workspace_bytes(T, B, beam)
            ws = torch.empty((max(ws_bytes, 8),), dtype=torch.uint8, device=dev)
            ngram = getattr(self, "ngram", None)
            if ngram is not None and hasattr(ngram, "struct"):
                import ctypes
                lm = ngram.struct(dev)
                nat.check(lib.hctr_ctc_prefix_beam_search_lm(
                    nat.ptr(tk_idx), nat.ptr(tk_lp), T, B, C, k, beam, float(self.lm_panelty), float(self.len_bonus),
                    None, ctypes.byref(lm), nat.ptr(idx), nat.ptr(ln), nat.ptr(status), nat.ptr(ws), ws_bytes, st),
                    "ctc_prefix_beam_search_lm")
            else:
                nat.check(lib.hctr_ctc_prefix_beam_search(
                    nat.ptr(tk_idx), nat.ptr(tk_lp), T, B, C, k, beam, float(self.lm_panelty), float(self.len_bonus),
                    nat.ptr(table), nat.ptr(idx), nat.ptr(ln), nat.ptr(status), nat.ptr(ws), ws_bytes, st),
                    "ctc_prefix_beam_search")
            if T == 0 or bool((status != 0).any().item()):
                # reference: top_line[-1] on an empty greedy path raises IndexError (utils/ctc_codec.py:198)
                raise IndexError("list index out of range")
        return idx, ln
