// Back-off n-gram language model on the device (order <= 5, single-character words = class indices).
// Reference: ctc_codec scores `prefix + suffix` with kenlm.Model.score(sentence, eos=False) at every beam step
// (utils/ctc_codec.py:120-122,276-279; 5-gram from lmplz, third-party/README.md:28-42). KenLM's query (lm/model.cc):
//   p(w | ctx) = prob(longest existing n-gram ctx[-j:] + w) + sum of the back-off weights of the longer contexts
//   ctx[-i:], i = j+1..len(ctx), added in increasing i; float32 throughout; unknown words -> <unk>; bos context <s>.
// Storage: one open-addressing hash table (linear probing) over all orders. Key = the n-gram's word ids, most recent
// word first, 16 bits each: lo = r0 | r1<<16 | r2<<32 | r3<<48, hi = r4 | order<<16 (hi == 0 marks an empty slot).
#pragma once
#include "common.cuh"
#include "../../include/hctr_b200.h"

namespace hctr {

constexpr int kNgramMaxOrder = 5;

int check_ngram(const hctr_ngram_lm* lm, const char* who);     // argument validation shared by the entry points (ngram_lm.cu)

__host__ __device__ __forceinline__ unsigned long long ngram_hash(unsigned long long lo, unsigned int hi) {
    unsigned long long z = lo ^ (static_cast<unsigned long long>(hi) * 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;          // splitmix64 finaliser
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// ids r[0..n-1], most recent first. Returns the slot or -1.
__device__ __forceinline__ long long ngram_find(const hctr_ngram_lm& lm, const int* r, int n, float* prob) {
    unsigned long long lo = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) if (i < n) lo |= static_cast<unsigned long long>(r[i] & 0xffff) << (16 * i);
    const unsigned int hi = (n > 4 ? (r[4] & 0xffff) : 0u) | (static_cast<unsigned int>(n) << 16);
    const uint4* tab = static_cast<const uint4*>(lm.entries);
    unsigned long long slot = ngram_hash(lo, hi) & lm.mask;
    for (;;) {
        const uint4 e = __ldg(tab + slot);
        if (e.z == 0u) return -1;
        if (e.z == hi && e.x == static_cast<unsigned int>(lo) && e.y == static_cast<unsigned int>(lo >> 32)) {
            *prob = __uint_as_float(e.w);
            return static_cast<long long>(slot);
        }
        slot = (slot + 1) & lm.mask;
    }
}

// log10 p(w | ctx) in float32. ctx[0..m-1]: LM word ids of the previous words, most recent first, m <= order-1.
__device__ inline float ngram_word_score(const hctr_ngram_lm& lm, const int* ctx, int m, int w) {
    int r[kNgramMaxOrder];
    r[0] = w;
#pragma unroll
    for (int i = 0; i < kNgramMaxOrder - 1; ++i) r[i + 1] = i < m ? ctx[i] : 0;
    float p = -100.f;
    int j = m;
    for (; j >= 0; --j)
        if (ngram_find(lm, r, j + 1, &p) >= 0) break;
    if (j < 0) { j = 0; p = -100.f; }                  // cannot happen: every id maps to a word with a unigram
    for (int i = j + 1; i <= m; ++i) {
        float dummy;
        const long long s = ngram_find(lm, ctx, i, &dummy);
        if (s >= 0) p = __fadd_rn(p, __ldg(lm.backoff + s));
    }
    return p;
}

// push a word in front of a most-recent-first context of capacity order-1
__device__ __forceinline__ void ngram_push(int* ctx, int& m, int cap, int w) {
#pragma unroll
    for (int i = kNgramMaxOrder - 2; i > 0; --i) ctx[i] = ctx[i - 1];
    ctx[0] = w;
    if (m < cap) ++m;
}

// LM context of the string (trie node `node`) + optional extra character `extra` (class index, -1 = none): the last
// order-1 words, most recent first, with <s> in front of a short string (kenlm score(..., bos=True))
__device__ inline void trie_context(const hctr_ngram_lm& lm, const int* __restrict__ n_parent, const int* __restrict__ n_chr,
                                    int node, int extra, int* ctx, int& m) {
    const int cap = lm.order - 1;
    m = 0;
#pragma unroll
    for (int i = 0; i < kNgramMaxOrder - 1; ++i) ctx[i] = 0;
    if (extra >= 0 && m < cap) ctx[m++] = __ldg(lm.vocab + extra);
    while (m < cap) {
        const int c = n_chr[node];
        if (c < 0) { ctx[m++] = lm.bos_id; break; }            // root of the trie: beginning of the sentence
        ctx[m++] = __ldg(lm.vocab + c);
        node = n_parent[node];
    }
}

}  // namespace hctr
