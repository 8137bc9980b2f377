"""Back-off n-gram language model for the device beam search.

The reference scores beams with `kenlm.Model(ngram_path).score(' '.join(prefix + suffix), eos=False)`
(utils/ctc_codec.py:120-122,276-279) on a 5-gram trained by `lmplz -o 5` over space-separated characters
(third-party/README.md:8-35). Here the ARPA text file is read once on the host and turned into the hash table that
`include/hctr_b200.h:hctr_ngram_lm` describes; every query then runs on the GPU inside the beam-search kernel
(csrc/ngram_lm.cuh). KenLM's binary format is not read - convert with the ARPA file `lmplz` wrote.

Word ids: class indices of the codec (1 .. C-2 are characters), then <s> = C, </s> = C+1, <unk> = C+2. N-grams that
contain a word outside the charset can never be queried by the decoder and are dropped.
"""
import ctypes

import numpy as np
import torch

MAX_ORDER = 5
_M64 = (1 << 64) - 1


class HctrNgramLm(ctypes.Structure):
    """ctypes mirror of `hctr_ngram_lm` (include/hctr_b200.h)."""
    _fields_ = [("entries", ctypes.c_void_p), ("backoff", ctypes.c_void_p), ("vocab", ctypes.c_void_p),
                ("mask", ctypes.c_ulonglong), ("order", ctypes.c_int), ("bos_id", ctypes.c_int),
                ("unk_id", ctypes.c_int), ("num_ids", ctypes.c_int)]


def _splitmix(z):
    z = z.astype(np.uint64)
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def _keys(ids):
    """ids: int array [n_grams, order] oldest word first -> (key_lo uint64, key_hi uint32), most recent word in the low bits."""
    n = ids.shape[1]
    rev = ids[:, ::-1].astype(np.uint64)
    lo = np.zeros(ids.shape[0], np.uint64)
    for i in range(min(n, 4)):
        lo |= rev[:, i] << np.uint64(16 * i)
    hi = np.full(ids.shape[0], n << 16, np.uint64)
    if n > 4:
        hi |= rev[:, 4]
    return lo, hi.astype(np.uint32)


def _slots(lo, hi, mask):
    with np.errstate(over="ignore"):
        return _splitmix(lo ^ (hi.astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15))) & np.uint64(mask)


def parse_arpa(text, word_to_id, bos_id, eos_id, unk_id):
    """-> (order, [per order: (ids int32 [n, order], prob float32 [n], backoff float32 [n])])."""
    special = {"<s>": bos_id, "</s>": eos_id, "<unk>": unk_id}
    per = {}
    section = 0
    for raw in text.splitlines():
        line = raw.strip()
        if not line:
            continue
        if line[0] == "\\":
            if line.endswith("-grams:"):
                section = int(line[1:line.index("-")])
                if section > MAX_ORDER:
                    raise ValueError("hctr_b200: n-gram order %d exceeds the device limit of %d" % (section, MAX_ORDER))
                per.setdefault(section, ([], [], []))
            elif line == "\\end\\":
                break
            else:
                section = 0
            continue
        if section == 0:
            continue
        cols = line.split("\t")
        if len(cols) >= 2:
            words, rest = cols[1].split(" "), cols[2:]
        else:
            cols = line.split()
            words, rest = cols[1:1 + section], cols[1 + section:]
        if len(words) != section:
            raise ValueError("hctr_b200: malformed ARPA line in the %d-gram section: %r" % (section, raw))
        ids = []
        for w in words:
            i = special.get(w)
            if i is None:
                i = word_to_id.get(w)
            if i is None:
                break
            ids.append(i)
        else:
            g = per[section]
            g[0].append(ids); g[1].append(float(cols[0])); g[2].append(float(rest[0]) if rest else 0.0)
    if not per or 1 not in per:
        raise ValueError("hctr_b200: no \\1-grams: section found (is this an ARPA file?)")
    order = max(per)
    out = []
    for n in range(1, order + 1):
        ids, p, b = per.get(n, ([], [], []))
        out.append((np.asarray(ids, np.int32).reshape(-1, n), np.asarray(p, np.float32), np.asarray(b, np.float32)))
    return order, out


class NgramLM(object):
    """Host-built hash table + device copy. `NgramLM.from_arpa(path, codec)`; `.struct()` is what the C ABI takes."""

    def __init__(self, order, grams, num_classes):
        self.order = order
        self.num_classes = num_classes
        self.bos_id, self.eos_id, self.unk_id = num_classes, num_classes + 1, num_classes + 2
        self.num_ids = num_classes + 3
        total = sum(g[0].shape[0] for g in grams)
        cap = 64
        while cap < 2 * total + 2:
            cap *= 2
        self.mask = cap - 1
        key_lo = np.zeros(cap, np.uint64); key_hi = np.zeros(cap, np.uint32)
        prob = np.zeros(cap, np.float32); backoff = np.zeros(cap, np.float32)
        uni = grams[0][0][:, 0] if grams[0][0].size else np.zeros(0, np.int32)
        if self.unk_id not in set(uni.tolist()):
            # lm/vocab.cc: a model without <unk> gets one with log10 p = -100
            grams = [(np.concatenate([grams[0][0], [[self.unk_id]]]).astype(np.int32),
                      np.concatenate([grams[0][1], [-100.0]]).astype(np.float32),
                      np.concatenate([grams[0][2], [0.0]]).astype(np.float32))] + list(grams[1:])
        for ids, p, b in grams:
            if ids.shape[0] == 0:
                continue
            lo, hi = _keys(ids)
            slot = _slots(lo, hi, self.mask).astype(np.int64)
            pending = np.arange(ids.shape[0])
            while pending.size:                                   # vectorised linear probing
                s = slot[pending]
                free = key_hi[s] == 0
                # several pending keys may want the same free slot: the first one (in file order) takes it
                order_idx = np.argsort(s, kind="stable")
                s_sorted = s[order_idx]
                first = np.ones(s_sorted.size, bool); first[1:] = s_sorted[1:] != s_sorted[:-1]
                winner = np.zeros(s.size, bool); winner[order_idx[first]] = True
                take = free & winner
                idx = pending[take]
                key_lo[s[take]] = lo[idx]; key_hi[s[take]] = hi[idx]; prob[s[take]] = p[idx]; backoff[s[take]] = b[idx]
                pending = pending[~take]
                slot[pending] = (slot[pending] + 1) & self.mask
        vocab = np.full(self.num_ids, self.unk_id, np.int32)
        have = grams[0][0][:, 0]
        vocab[have] = have
        self.host = {"key_lo": key_lo, "key_hi": key_hi, "prob": prob, "backoff": backoff, "vocab": vocab}
        self.n_grams = total
        self._dev = {}

    @classmethod
    def from_arpa_text(cls, text, characters_dict, num_classes):
        words = {ch: i for ch, i in characters_dict.items() if 0 < i < num_classes - 1 and len(ch) == 1}
        order, grams = parse_arpa(text, words, num_classes, num_classes + 1, num_classes + 2)
        return cls(order, grams, num_classes)

    @classmethod
    def from_arpa(cls, path, codec):
        with open(path, encoding="utf-8") as fh:
            return cls.from_arpa_text(fh.read(), codec.dict, len(codec.characters))

    # ---- device side
    def to(self, device):
        device = torch.device(device)
        d = self._dev.get(device)
        if d is None:
            h = self.host
            ent = np.empty((h["key_lo"].size, 4), np.uint32)
            ent[:, 0] = (h["key_lo"] & np.uint64(0xffffffff)).astype(np.uint32)
            ent[:, 1] = (h["key_lo"] >> np.uint64(32)).astype(np.uint32)
            ent[:, 2] = h["key_hi"]
            ent[:, 3] = h["prob"].view(np.uint32)
            d = {"entries": torch.from_numpy(ent.view(np.int32)).to(device),
                 "backoff": torch.from_numpy(h["backoff"]).to(device),
                 "vocab": torch.from_numpy(h["vocab"]).to(device)}
            self._dev[device] = d
        return d

    def struct(self, device):
        d = self.to(device)
        return HctrNgramLm(d["entries"].data_ptr(), d["backoff"].data_ptr(), d["vocab"].data_ptr(), self.mask, self.order,
                           self.bos_id, self.unk_id, self.num_ids)

    def score_ids(self, sequences, device="cuda"):
        """kenlm score(bos=True, eos=False) of class-index sequences on the device -> float32 numpy [n]."""
        from . import native as nat
        lib = nat.lib()
        dev = torch.device(device)
        off = np.zeros(len(sequences) + 1, np.int32)
        off[1:] = np.cumsum([len(s) for s in sequences])
        flat = np.concatenate([np.asarray(s, np.int32).reshape(-1) for s in sequences]) if off[-1] else np.zeros(1, np.int32)
        with torch.cuda.device(dev):
            ids = torch.from_numpy(flat).to(dev); offs = torch.from_numpy(off).to(dev)
            out = torch.empty((max(len(sequences), 1),), dtype=torch.float32, device=dev)
            st = self.struct(dev)
            nat.check(lib.hctr_ngram_score(ctypes.byref(st), nat.ptr(ids), nat.ptr(offs), len(sequences), nat.ptr(out),
                                           nat.stream_ptr()), "ngram_score")
            return out[:len(sequences)].cpu().numpy()

    # ---- host restatement of the device probe (used by CPU tests of the table builder)
    def host_find(self, ids_oldest_first):
        ids = np.asarray(ids_oldest_first, np.int32).reshape(1, -1)
        lo, hi = _keys(ids)
        s = int(_slots(lo, hi, self.mask)[0])
        h = self.host
        while h["key_hi"][s] != 0:
            if h["key_hi"][s] == hi[0] and h["key_lo"][s] == lo[0]:
                return float(h["prob"][s]), float(h["backoff"][s])
            s = (s + 1) & self.mask
        return None
