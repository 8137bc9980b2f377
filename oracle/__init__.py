"""oracle/ - TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement of the reference's HCTR hot path (AndrewCullacino/handwritten-chinese-ocr-samples):
  * ctc_oracle.c      plain C: greedy decode, log-softmax/top-k, prefix beam search, CTC loss fwd/bwd
                      (utils/ctc_codec.py:63-99,183-307; main.py:205,406-409)
  * hctr_forward.py   torch fp32 functional restatement of hctr_model.forward
                      (models/handwritten_ctr_model.py:26-30,47-60,115-153,171-178)
  * codec.py          pure-Python/numpy restatement of ctc_codec.encode and the index->string mapping

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
package, and only as the checker / the timed CPU baseline. The product package
(handwritten-chinese-ocr-samples_b200/) never imports it and has no CPU fallback.

Parity pin: the reference has no tests or golden vectors of its own ("parity unpinned"
by the reference's tests), so the oracle is pinned against outputs of the reference itself: the Python
reference was imported in the build container by tests/golden/make_golden.py and its outputs are
committed under tests/golden/*.npz; tests/test_oracle_golden.py replays them against this package.
The reference is pure Python, so there is nothing to compile into oracle/_ref/.
"""
import ctypes
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_DIR, "ctc_oracle.c")
_OUT_DIR = os.path.join(_DIR, "_build")
_LIB = os.path.join(_OUT_DIR, "liboracle.so")
_lib = None


def build(force=False):
    """gcc -O2 -shared ctc_oracle.c -> oracle/_build/liboracle.so (no -ffast-math: IEEE semantics matter)."""
    os.makedirs(_OUT_DIR, exist_ok=True)
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(_SRC):
        subprocess.run(["gcc", "-O2", "-fPIC", "-shared", "-fno-fast-math", "-ffp-contract=off", "-o", _LIB, _SRC, "-lm"],
                       check=True)
    return _LIB


def _c():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB)
        P, I, D = ctypes.c_void_p, ctypes.c_int, ctypes.c_double
        L.oracle_ctc_greedy.argtypes = [P, I, I, I, P, P, P]
        L.oracle_ctc_greedy.restype = None
        L.oracle_log_softmax.argtypes = [P, I, I, P]
        L.oracle_log_softmax.restype = None
        L.oracle_topk_row.argtypes = [P, I, I, P]
        L.oracle_topk_row.restype = None
        L.oracle_ctc_beam_search.argtypes = [P, P, I, I, I, I, I, D, D, P, P, P, P]
        L.oracle_ctc_beam_search.restype = I
        L.oracle_ctc_beam_search_skip.argtypes = [P, P, I, I, I, I, D, D, P, P, P, P]
        L.oracle_ctc_beam_search_skip.restype = I
        L.oracle_edit_distance.argtypes = [P, I, P, I]
        L.oracle_edit_distance.restype = I
        L.oracle_ctc_loss.argtypes = [P, I, I, I, P, P, P, P, P]
        L.oracle_ctc_loss.restype = D
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def greedy_decode(preds):
    """preds: np.float32 [T,B,C] -> (raw_argmax int32 [B,T], idx int32 [B,T], length int32 [B])."""
    preds = np.ascontiguousarray(preds, dtype=np.float32)
    T, B, C = preds.shape
    raw = np.zeros((B, T), np.int32)
    idx = np.zeros((B, T), np.int32)
    ln = np.zeros((B,), np.int32)
    if T > 0 and B > 0:
        _c().oracle_ctc_greedy(_p(preds), T, B, C, _p(raw), _p(idx), _p(ln))
    return raw, idx, ln


def log_softmax(x):
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty_like(x)
    C = x.shape[-1]
    _c().oracle_log_softmax(_p(x), x.size // C, C, _p(out))
    return out


def topk(logp, k):
    """[T,B,C] -> int32 [T,B,k], descending (ties -> lower index)."""
    logp = np.ascontiguousarray(logp, dtype=np.float32)
    T, B, C = logp.shape
    out = np.zeros((T, B, k), np.int32)
    L = _c()
    for t in range(T):
        for b in range(B):
            L.oracle_topk_row(_p(logp[t, b]), C, k, _p(out[t, b]))
    return out


def beam_search(logits, beam_size=10, search_depth=10, lm_penalty=2.0, len_bonus=5.8, lm_table=None):
    """Full restatement of decode() in beam mode: log_softmax -> top-k -> __cbs_full__.
    Returns (idx int32 [B,T], length int32 [B], status int32 [B])."""
    logits = np.ascontiguousarray(logits, dtype=np.float32)
    T, B, C = logits.shape
    logp = log_softmax(logits)
    tk = topk(logp, search_depth)
    idx = np.zeros((B, T), np.int32)
    ln = np.zeros((B,), np.int32)
    st = np.zeros((B,), np.int32)
    tab = None if lm_table is None else np.ascontiguousarray(lm_table, dtype=np.float64)
    _c().oracle_ctc_beam_search(_p(logp), _p(tk), T, B, C, search_depth, beam_size, float(lm_penalty), float(len_bonus),
                                _p(tab), _p(idx), _p(ln), _p(st))
    return idx, ln, st


def beam_search_skip(logits, beam_size=10, lm_penalty=2.0, len_bonus=5.8, lm_table=None):
    """decode() with skip_search=True: log_softmax -> __cbs_skip__. Returns (idx [B,T], length [B], status [B])."""
    logits = np.ascontiguousarray(logits, dtype=np.float32)
    T, B, C = logits.shape
    logp = log_softmax(logits)
    top1 = np.ascontiguousarray(np.argmax(logp, axis=2).astype(np.int32))
    idx = np.zeros((B, T), np.int32)
    ln = np.zeros((B,), np.int32)
    st = np.zeros((B,), np.int32)
    tab = None if lm_table is None else np.ascontiguousarray(lm_table, dtype=np.float64)
    _c().oracle_ctc_beam_search_skip(_p(logp), _p(top1), T, B, C, beam_size, float(lm_penalty), float(len_bonus), _p(tab),
                                     _p(idx), _p(ln), _p(st))
    return idx, ln, st


def edit_distance(a, b):
    """Levenshtein distance between two label sequences (or strings, compared by code point)."""
    if isinstance(a, str):
        a = [ord(c) for c in a]
    if isinstance(b, str):
        b = [ord(c) for c in b]
    a = np.ascontiguousarray(a, dtype=np.int32)
    b = np.ascontiguousarray(b, dtype=np.int32)
    return int(_c().oracle_edit_distance(_p(a), len(a), _p(b), len(b)))


def ctc_loss(logits, targets, input_lengths, target_lengths, need_grad=True):
    """CTCLoss(zero_infinity=True, reduction='mean') on log_softmax(logits) and d loss / d logits.
    logits np.float32 [T,B,C]. Returns (loss float, nll float64 [B], grad float64 [T,B,C] or None)."""
    logits = np.ascontiguousarray(logits, dtype=np.float32)
    T, B, C = logits.shape
    targets = np.ascontiguousarray(targets, dtype=np.int32)
    tl = np.ascontiguousarray(target_lengths, dtype=np.int32)
    il = np.ascontiguousarray(input_lengths, dtype=np.int32)
    nll = np.zeros((B,), np.float64)
    grad = np.zeros((T, B, C), np.float64) if need_grad else None
    loss = _c().oracle_ctc_loss(_p(logits), T, B, C, _p(targets), _p(tl), _p(il), _p(nll), _p(grad))
    return loss, nll, grad
