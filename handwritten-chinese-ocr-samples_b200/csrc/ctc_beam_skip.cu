// ctc_codec.__cbs_skip__ on device (reference: utils/ctc_codec.py:124-181 with __context_beam_search__ :212-285).
//   kernel 1 (HBM-bound, one CTA per (t,b) row): log-softmax, the arg-max class (for the greedy look-ahead), the blank
//            log-prob, and the "pruned" candidate list = every class with log-prob > log(0.001), in index order (:144).
//   kernel 2 (latency-bound, one CTA per sequence): exactly one candidate -> the reference's in-place fast path
//            (:147-171, quirks included: a blank step leaves pnb untouched, equal prefixes are NOT merged); otherwise a
//            context beam search over the candidates. Because the fast path can leave several kept beams with the SAME
//            prefix, the search here is duplicate-aware: beams are grouped by string, dict entries are keyed by the group,
//            and every entry accumulates its contributions in ascending beam order, which is the reference's order.
// float64 accumulators, np.logaddexp branch structure, stable ranking by insertion order.
#include <cfloat>

#include <cstring>

#include "common.cuh"
#include "ngram_lm.cuh"
#include "../../include/hctr_b200.h"

namespace hctr {

// Candidates per step: the reference takes every class with p > 0.001 (utils/ctc_codec.py:144), i.e. at most 999. The
// common case (<= 128) keeps the candidate tables small and the dict entries of a step in shared memory; a step with more
// makes the call report HCTR_ERR_UNSUPPORTED for that sequence, and the caller repeats the call with max_candidates = 1024:
// same kernels, 8 KB of candidate table per row and the dict entries of a step in the global workspace.
constexpr int kSkMaxC = 128;
constexpr int kSkBigC = 1024;
constexpr int kSkMaxBeam = 16;
constexpr int kSkThreads = 128;
constexpr int kPruneThreads = 256;

template <typename T> struct SkLoad;
template <> struct SkLoad<float> {
    static constexpr int N = 4;
    static __device__ __forceinline__ void load(const float* p, float (&o)[4]) {
        const uint4 q = ld_nc_v4(p);
        o[0] = __uint_as_float(q.x); o[1] = __uint_as_float(q.y); o[2] = __uint_as_float(q.z); o[3] = __uint_as_float(q.w);
    }
    static __device__ __forceinline__ float one(const float* p) { return __ldg(p); }
};
template <> struct SkLoad<__nv_bfloat16> {
    static constexpr int N = 8;
    static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&o)[8]) {
        const uint4 q = ld_nc_v4(p);
        o[0] = bf16_lo(q.x); o[1] = bf16_hi(q.x); o[2] = bf16_lo(q.y); o[3] = bf16_hi(q.y);
        o[4] = bf16_lo(q.z); o[5] = bf16_hi(q.z); o[6] = bf16_lo(q.w); o[7] = bf16_hi(q.w);
    }
    static __device__ __forceinline__ float one(const __nv_bfloat16* p) {
        return __uint_as_float(static_cast<uint32_t>(*reinterpret_cast<const unsigned short*>(p)) << 16);
    }
};

// per row outputs: meta[row] = {count, top1}, blank_lp[row], cand_idx[row][MAXC] (ascending), cand_lp[row][MAXC]
template <typename T, int MAXC>
__global__ void __launch_bounds__(kPruneThreads)
ctc_prune_logsoftmax_kernel(const T* __restrict__ logits, int Tn, int Bn, int C, long long stride_t, long long stride_b,
                            int2* __restrict__ meta, float* __restrict__ blank_lp, int32_t* __restrict__ cand_idx,
                            float* __restrict__ cand_lp) {
    constexpr int V = SkLoad<T>::N;
    extern __shared__ float rowbuf[];
    __shared__ float red[8];
    __shared__ int redi[8];
    __shared__ float s_max, s_logs;
    __shared__ int s_top1, s_n;
    __shared__ int c_i[MAXC];
    __shared__ float c_v[MAXC];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long row = blockIdx.x;                              // row = t*B + b
    const int t = (int)(row / Bn), b = (int)(row - (long long)t * Bn);
    const T* p = logits + (long long)t * stride_t + (long long)b * stride_b;
    float tmax = -INFINITY; int targ = 0x7fffffff;
    auto see = [&](float x, int c) { rowbuf[c] = x; if (x > tmax || (x == tmax && c < targ)) { tmax = x; targ = c; } };
    const uintptr_t addr = reinterpret_cast<uintptr_t>(p);
    int head = (int)(((16 - (addr & 15)) & 15) / sizeof(T));
    if (head > C) head = C;
    if (tid < head) see(SkLoad<T>::one(p + tid), tid);
    const int nvec = (C - head) / V;
    for (int vi = tid; vi < nvec; vi += kPruneThreads) {
        float x[V];
        SkLoad<T>::load(p + head + (long long)vi * V, x);
#pragma unroll
        for (int j = 0; j < V; ++j) see(x[j], head + vi * V + j);
    }
    const int tail0 = head + nvec * V;
    if (tail0 + tid < C) see(SkLoad<T>::one(p + tail0 + tid), tail0 + tid);
    if (tid == 0) s_n = 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, tmax, o);
        const int oi = __shfl_xor_sync(0xffffffffu, targ, o);
        if (ov > tmax || (ov == tmax && oi < targ)) { tmax = ov; targ = oi; }
    }
    if (lane == 0) { red[warp] = tmax; redi[warp] = targ; }
    __syncthreads();
    if (tid == 0) {
        float m = red[0]; int a = redi[0];
        for (int i = 1; i < kPruneThreads / 32; ++i) if (red[i] > m || (red[i] == m && redi[i] < a)) { m = red[i]; a = redi[i]; }
        s_max = m; s_top1 = a;
    }
    __syncthreads();
    const float m = s_max;
    float sum = 0.f;
    for (int c = tid; c < C; c += kPruneThreads) sum += __expf(rowbuf[c] - m);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    __syncthreads();
    if (lane == 0) red[warp] = sum;
    __syncthreads();
    if (tid == 0) {
        float sacc = 0.f;
        for (int i = 0; i < kPruneThreads / 32; ++i) sacc += red[i];
        s_logs = logf(sacc);
    }
    __syncthreads();
    const float logs = s_logs;
    const double thresh = -6.907755278982137;                      // np.log(0.001) (float64), compared in double (:129,144)
    for (int c = tid; c < C; c += kPruneThreads) {
        const float lp = (rowbuf[c] - m) - logs;                   // scipy: (x - max) - log(sum(exp(x - max)))
        if ((double)lp > thresh) {
            const int slot = atomicAdd(&s_n, 1);
            if (slot < MAXC) { c_i[slot] = c; c_v[slot] = lp; }
        }
    }
    __syncthreads();
    const int n = s_n;
    if (tid == 0) { meta[row] = make_int2(n, s_top1); blank_lp[row] = (rowbuf[0] - m) - logs; }
    if (n <= MAXC) {
        // index order: rank by counting (indices are distinct)
        for (int e = tid; e < n; e += kPruneThreads) {
            int rank = 0;
            for (int f = 0; f < n; ++f) rank += c_i[f] < c_i[e];
            cand_idx[row * MAXC + rank] = c_i[e];
            cand_lp[row * MAXC + rank] = c_v[e];
        }
    }
}

// ------------------------------------------------------------------------------------------------ search
__device__ __forceinline__ double sk_logaddexp(double x, double y) {
    // numpy npy_logaddexp, written with selects (one exp + one log1p for all lanes; same operations, same bits)
    const double d = __dsub_rn(x, y);
    const bool pos = d > 0;
    const double hi = pos ? x : y;
    const double arg = pos ? -d : d;
    const double r = __dadd_rn(hi, log1p(exp(arg)));
    if (x == y) return __dadd_rn(x, 0.693147180559945309417232121458176568);
    if (d != d) return d;
    return r;
}
__device__ __forceinline__ unsigned long long sk_mix(unsigned long long h, int c) {
    h ^= (unsigned long long)(c + 1) * 0x9E3779B97F4A7C15ull;
    h *= 0xFF51AFD7ED558CCDull;
    h ^= h >> 29;
    return h;
}
__device__ bool sk_same(const int* parent, const int* chr, int a, int b) {
    while (a != b) {
        if (chr[a] != chr[b]) return false;
        a = parent[a]; b = parent[b];
    }
    return true;
}

struct SkKept {
    int node[kSkMaxBeam], len[kSkMaxBeam], last[kSkMaxBeam];
    unsigned long long hash[kSkMaxBeam];
    double pb[kSkMaxBeam], pnb[kSkMaxBeam], lmsum[kSkMaxBeam];
};

template <int MAXC>
__global__ void __launch_bounds__(kSkThreads)
ctc_skip_beam_kernel(const int2* __restrict__ meta, const float* __restrict__ blank_lp, const int32_t* __restrict__ cand_idx,
                     const float* __restrict__ cand_lp, int Tn, int Bn, int C, int beam_size, double lm_penalty,
                     double len_bonus, const double* __restrict__ lm_table, const hctr_ngram_lm ng_lm, int32_t* __restrict__ out_idx,
                     int32_t* __restrict__ out_len, int32_t* __restrict__ status, unsigned char* __restrict__ workspace,
                     long long ws_per_seq, unsigned char* __restrict__ entry_ws, long long entry_ws_per_seq) {
    __shared__ SkKept kept[2];
    __shared__ double Pj[kSkMaxBeam];
    __shared__ int canon[kSkMaxBeam], parentc[kSkMaxBeam], ent_of_canon[kSkMaxBeam];
    __shared__ int cand[MAXC];
    __shared__ double candp[MAXC];
    // dict entries of one step, beam_size * (MAXC + 1) slots, and the new-entry map [kSkMaxBeam][MAXC]: dynamic shared memory,
    // or (entry_ws != nullptr, the 1024-candidate variant) this sequence's slice of the global workspace
    extern __shared__ double sk_dyn[];
    const int gen_cap = beam_size * (MAXC + 1);
    double* e_base = entry_ws ? reinterpret_cast<double*>(entry_ws + (long long)blockIdx.x * entry_ws_per_seq) : sk_dyn;
    double* e_pb = e_base;
    double* e_pnb = e_pb + gen_cap;
    double* e_tot = e_pnb + gen_cap;
    int* e_chr = reinterpret_cast<int*>(e_tot + gen_cap);
    short* e_kind = reinterpret_cast<short*>(e_chr + gen_cap);
    short* e_src = e_kind + gen_cap;
    short (*new_ent)[MAXC] = reinterpret_cast<short (*)[MAXC]>(e_src + gen_cap);
    __shared__ int s_ngen, s_ng, s_fail;

    const int b = blockIdx.x, tid = threadIdx.x;
    const int unknown = C - 1;
    unsigned char* ws = workspace + (long long)b * ws_per_seq;
    int* g_char = reinterpret_cast<int*>(ws);
    int* g_time = g_char + Tn;
    const int cap = Tn * kSkMaxBeam + 1;
    int* n_parent = g_time + Tn;
    int* n_chr = n_parent + cap;
    unsigned long long* n_hash = reinterpret_cast<unsigned long long*>((reinterpret_cast<uintptr_t>(n_chr + cap) + 7) & ~uintptr_t(7));

    // ---- greedy (char, t) list from the arg-max classes (:130-137); sequential compaction is fine (T <= few thousand)
    if (tid == 0) {
        int ng = 0, prev = -1;
        for (int t = 0; t < Tn; ++t) {
            const int cur = meta[(long long)t * Bn + b].y;
            if (cur != 0 && cur != unknown && !(t > 0 && prev == cur)) { g_char[ng] = cur; g_time[ng] = t; ++ng; }
            prev = cur;
        }
        s_ng = ng; s_fail = 0;
        n_parent[0] = 0; n_chr[0] = -1; n_hash[0] = 0x243F6A8885A308D3ull;
        kept[0].node[0] = 0; kept[0].len[0] = 0; kept[0].last[0] = -1; kept[0].hash[0] = n_hash[0];
        kept[0].pb[0] = 0.0; kept[0].pnb[0] = -INFINITY; kept[0].lmsum[0] = 0.0;
    }
    __syncthreads();
    const int ng = s_ng;
    if (ng == 0) {                                        // top_line[-1] -> IndexError (:139)
        if (tid == 0) { status[b] = HCTR_ERR_INDEX; out_len[b] = 0; }
        return;
    }
    int end_step = g_time[ng - 1] + 4;
    if (end_step >= Tn) end_step = Tn;
    int nkept = 1, cur_buf = 0, gptr = 0;

    for (int t = 0; t < end_step; ++t) {
        const long long row = (long long)t * Bn + b;
        const int nc = meta[row].x;
        if (nc > MAXC) {                               // more candidates than the device path holds
            if (tid == 0) { status[b] = HCTR_ERR_UNSUPPORTED; out_len[b] = 0; }
            return;
        }
        SkKept& K = kept[cur_buf];
        if (nc == 1) {
            // ---------------- in-place fast path (:147-171), one thread per kept beam
            const int pidx = cand_idx[row * MAXC];
            if (pidx >= unknown) continue;                // :150-151 (block-uniform)
            if (tid < nkept) {
                const int j = tid;
                const double p = (double)cand_lp[row * MAXC], p0 = (double)blank_lp[row];
                const double pr = sk_logaddexp(K.pb[j], K.pnb[j]);
                bool extend = false;
                if (pidx == 0) {
                    K.pb[j] = __dadd_rn(pr, p0);                                      // pnb stays (reference quirk)
                } else if (pidx != K.last[j]) {
                    extend = true; K.pnb[j] = __dadd_rn(pr, p); K.pb[j] = -INFINITY;
                } else if (K.pb[j] != -INFINITY) {
                    extend = true; K.pnb[j] = __dadd_rn(K.pb[j], p); K.pb[j] = -INFINITY;
                } else {
                    K.pb[j] = __dadd_rn(pr, p0);
                    K.pnb[j] = __dadd_rn(K.pnb[j], p);
                }
                if (extend) {
                    const int id = 1 + t * kSkMaxBeam + j;
                    const unsigned long long h = sk_mix(K.hash[j], pidx);
                    n_parent[id] = K.node[j]; n_chr[id] = pidx; n_hash[id] = h;
                    K.node[j] = id; K.len[j] += 1; K.last[j] = pidx; K.hash[j] = h;
                    if (lm_table) K.lmsum[j] = __dadd_rn(K.lmsum[j], lm_table[pidx]);
                    if (ng_lm.entries) {                                              // kenlm: float32 running total
                        int ctx[kNgramMaxOrder - 1]; int m;
                        trie_context(ng_lm, n_parent, n_chr, n_parent[id], -1, ctx, m);
                        K.lmsum[j] = (double)__fadd_rn((float)K.lmsum[j], ngram_word_score(ng_lm, ctx, m, __ldg(ng_lm.vocab + pidx)));
                    }
                }
            }
            __threadfence_block();
            __syncthreads();
            continue;
        }
        if (nc == 0) {                                    // no candidate: the beam list empties -> IndexError at the end (:179)
            if (tid == 0) { status[b] = HCTR_ERR_INDEX; out_len[b] = 0; }
            return;
        }
        // ---------------- context beam search over the pruned candidates
        SkKept& Kn = kept[cur_buf ^ 1];
        while (gptr < ng && g_time[gptr] <= t) ++gptr;
        int nsuf = ng - gptr; if (nsuf > 4) nsuf = 4;
        for (int i = tid; i < nc; i += kSkThreads) { cand[i] = cand_idx[row * MAXC + i]; candp[i] = (double)cand_lp[row * MAXC + i]; }
        if (tid < nkept) {
            Pj[tid] = sk_logaddexp(K.pb[tid], K.pnb[tid]);
            int cn = tid;                                  // first beam with the same string
            for (int j = 0; j < tid; ++j)
                if (K.len[j] == K.len[tid] && K.hash[j] == K.hash[tid] && sk_same(n_parent, n_chr, K.node[j], K.node[tid])) { cn = j; break; }
            canon[tid] = cn;
            ent_of_canon[tid] = -1;
        }
        for (int i = tid; i < kSkMaxBeam * MAXC; i += kSkThreads) new_ent[i / MAXC][i % MAXC] = -1;
        __syncthreads();
        if (tid < nkept) {
            int pk = -1;                                   // canonical beam whose string is this one minus its last char
            if (canon[tid] == tid && K.len[tid] > 0) {
                const int par = n_parent[K.node[tid]];
                for (int j = 0; j < nkept; ++j) {
                    if (canon[j] != j || K.len[j] != K.len[tid] - 1) continue;
                    if (K.node[j] == par || (K.hash[j] == n_hash[par] && sk_same(n_parent, n_chr, K.node[j], par))) { pk = j; break; }
                }
            }
            parentc[tid] = pk;
        }
        __syncthreads();
        // insertion order of the reference's gen_beams dict (:235-255)
        if (tid == 0) {
            int ngen = 0;
            for (int j = 0; j < nkept; ++j) {
                const int Kc = canon[j];
                for (int q = 0; q < nc; ++q) {
                    const int idx = cand[q];
                    if (idx >= unknown) continue;
                    if (ent_of_canon[Kc] < 0) { ent_of_canon[Kc] = ngen; e_kind[ngen] = (short)Kc; ++ngen; }
                    if (idx == 0) continue;
                    int tgt = -1;
                    for (int j2 = 0; j2 < nkept; ++j2)
                        if (canon[j2] == j2 && parentc[j2] == Kc && K.last[j2] == idx) { tgt = j2; break; }
                    if (tgt >= 0) {
                        if (ent_of_canon[tgt] < 0) { ent_of_canon[tgt] = ngen; e_kind[ngen] = (short)tgt; ++ngen; }
                    } else if (new_ent[Kc][q] < 0) {
                        new_ent[Kc][q] = (short)ngen; e_kind[ngen] = -1; e_src[ngen] = (short)Kc; e_chr[ngen] = idx; ++ngen;
                    }
                }
            }
            s_ngen = ngen;
        }
        __syncthreads();
        const int ngen = s_ngen;
        // contributions in ascending beam order (= the reference's accumulation order per field)
        for (int e = tid; e < ngen; e += kSkThreads) {
            double pb = -INFINITY, pnb = -INFINITY, lm, plen;
            if (e_kind[e] >= 0) {
                const int K2 = e_kind[e];
                const int par = parentc[K2];
                for (int j = 0; j < nkept; ++j) {
                    if (canon[j] == K2) {
                        for (int q = 0; q < nc; ++q) {
                            const int idx = cand[q];
                            if (idx >= unknown) continue;
                            if (idx == 0) pb = sk_logaddexp(pb, __dadd_rn(Pj[j], candp[q]));
                            else if (idx == K.last[j]) pnb = sk_logaddexp(pnb, __dadd_rn(K.pnb[j], candp[q]));
                        }
                    } else if (par >= 0 && canon[j] == par) {
                        for (int q = 0; q < nc; ++q) {
                            const int idx = cand[q];
                            if (idx != K.last[K2] || idx >= unknown || idx == 0) continue;
                            pnb = sk_logaddexp(pnb, (idx != K.last[j]) ? __dadd_rn(Pj[j], candp[q]) : __dadd_rn(K.pb[j], candp[q]));
                        }
                    }
                }
                lm = K.lmsum[K2]; plen = (double)K.len[K2];
            } else {
                const int Kc = e_src[e], idx = e_chr[e];
                double p = 0.0;
                for (int q = 0; q < nc; ++q) if (cand[q] == idx) p = candp[q];
                for (int j = 0; j < nkept; ++j)
                    if (canon[j] == Kc)
                        pnb = sk_logaddexp(pnb, (idx != K.last[j]) ? __dadd_rn(Pj[j], p) : __dadd_rn(K.pb[j], p));
                lm = lm_table ? __dadd_rn(K.lmsum[Kc], lm_table[idx]) : 0.0;
                if (ng_lm.entries) {
                    int ctx[kNgramMaxOrder - 1]; int m;
                    trie_context(ng_lm, n_parent, n_chr, K.node[Kc], -1, ctx, m);
                    lm = (double)__fadd_rn((float)K.lmsum[Kc], ngram_word_score(ng_lm, ctx, m, __ldg(ng_lm.vocab + idx)));
                }
                plen = (double)(K.len[Kc] + 1);
            }
            double lmt = lm;
            if (lm_table) for (int c = 0; c < nsuf; ++c) lmt = __dadd_rn(lmt, lm_table[g_char[gptr + c]]);
            if (ng_lm.entries && nsuf > 0) {
                int ctx[kNgramMaxOrder - 1]; int m;
                if (e_kind[e] >= 0) trie_context(ng_lm, n_parent, n_chr, K.node[e_kind[e]], -1, ctx, m);
                else trie_context(ng_lm, n_parent, n_chr, K.node[e_src[e]], e_chr[e], ctx, m);
                float tot = (float)lm;
                for (int c = 0; c < nsuf; ++c) {
                    const int w = __ldg(ng_lm.vocab + g_char[gptr + c]);
                    tot = __fadd_rn(tot, ngram_word_score(ng_lm, ctx, m, w));
                    ngram_push(ctx, m, ng_lm.order - 1, w);
                }
                lmt = (double)tot;
            }
            const double pt = __dadd_rn(__dmul_rn(lmt, lm_penalty), __dmul_rn(plen, len_bonus));
            e_pb[e] = pb; e_pnb[e] = pnb;
            e_tot[e] = __dadd_rn(sk_logaddexp(pb, pnb), pt);
        }
        __syncthreads();
        const int keep_n = ngen < beam_size ? ngen : beam_size;
        for (int e = tid; e < ngen; e += kSkThreads) {
            const double te = e_tot[e];
            int rank = 0;
            for (int f = 0; f < ngen; ++f) rank += (e_tot[f] > te) || (e_tot[f] == te && f < e);
            if (rank < keep_n) {
                Kn.pb[rank] = e_pb[e]; Kn.pnb[rank] = e_pnb[e];
                if (e_kind[e] >= 0) {
                    const int j2 = e_kind[e];
                    Kn.node[rank] = K.node[j2]; Kn.len[rank] = K.len[j2]; Kn.last[rank] = K.last[j2]; Kn.hash[rank] = K.hash[j2];
                    Kn.lmsum[rank] = K.lmsum[j2];
                } else {
                    const int j = e_src[e], idx = e_chr[e];
                    const int id = 1 + t * kSkMaxBeam + rank;
                    const unsigned long long h = sk_mix(K.hash[j], idx);
                    n_parent[id] = K.node[j]; n_chr[id] = idx; n_hash[id] = h;
                    Kn.node[rank] = id; Kn.len[rank] = K.len[j] + 1; Kn.last[rank] = idx; Kn.hash[rank] = h;
                    double nl = lm_table ? __dadd_rn(K.lmsum[j], lm_table[idx]) : 0.0;
                    if (ng_lm.entries) {                  // the same float32 sum the scoring loop formed for this entry
                        int ctx[kNgramMaxOrder - 1]; int m;
                        trie_context(ng_lm, n_parent, n_chr, K.node[j], -1, ctx, m);
                        nl = (double)__fadd_rn((float)K.lmsum[j], ngram_word_score(ng_lm, ctx, m, __ldg(ng_lm.vocab + idx)));
                    }
                    Kn.lmsum[rank] = nl;
                }
            }
        }
        __threadfence_block();
        __syncthreads();
        nkept = keep_n;
        cur_buf ^= 1;
        if (nkept == 0) {
            if (tid == 0) { status[b] = HCTR_ERR_INDEX; out_len[b] = 0; }
            return;
        }
    }
    if (tid == 0) {
        const SkKept& K = kept[cur_buf];
        const int L = K.len[0];
        int node = K.node[0];
        for (int c = L - 1; c >= 0; --c) { out_idx[(long long)b * Tn + c] = n_chr[node]; node = n_parent[node]; }
        out_len[b] = L;
        status[b] = 0;
    }
}

static long long sk_ws_per_seq(int T) {
    const long long cap = (long long)T * kSkMaxBeam + 1;
    long long bytes = 8ll * T + 8ll * cap + 8 + 8ll * cap;
    return (bytes + 15) & ~15ll;
}

static size_t sk_entry_bytes(int beam_size, int maxc) {
    return (size_t)beam_size * (maxc + 1) * (3 * sizeof(double) + sizeof(int) + 2 * sizeof(short)) + (size_t)kSkMaxBeam * maxc * sizeof(short) + 64;
}

template <int MAXC>
static int sk_launch(const void* logits, int dtype, int T, int B, int C, long long stride_t, long long stride_b, int beam_size,
                     double lm_penalty, double len_bonus, const double* lm_table, const hctr_ngram_lm& lm, int32_t* out_idx,
                     int32_t* out_len, int32_t* status, void* workspace, cudaStream_t s) {
    const long long rows = (long long)T * B;
    const size_t smem = (size_t)C * sizeof(float);
    char* base = static_cast<char*>(workspace);
    int2* meta = reinterpret_cast<int2*>(base);
    float* blank = reinterpret_cast<float*>(base + rows * 8);
    int32_t* cidx = reinterpret_cast<int32_t*>(base + rows * 12);
    float* clp = reinterpret_cast<float*>(base + rows * 12 + rows * MAXC * 4);
    long long tables = rows * (8 + 4 + (long long)MAXC * 8);
    tables = (tables + 255) & ~255ll;
    unsigned char* seq_ws = reinterpret_cast<unsigned char*>(base + tables);
    const long long seq_bytes = (sk_ws_per_seq(T) * B + 255) & ~255ll;
    const bool global_entries = MAXC > kSkMaxC;
    unsigned char* entry_ws = global_entries ? seq_ws + seq_bytes : nullptr;
    const long long entry_per_seq = ((long long)sk_entry_bytes(kSkMaxBeam, MAXC) + 255) & ~255ll;
    static PerDeviceOnce once;
    int dev;
    if (once.need(dev)) {
        HCTR_CUDA(cudaFuncSetAttribute(ctc_prune_logsoftmax_kernel<float, MAXC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        HCTR_CUDA(cudaFuncSetAttribute(ctc_prune_logsoftmax_kernel<__nv_bfloat16, MAXC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        if (!global_entries)
            HCTR_CUDA(cudaFuncSetAttribute(ctc_skip_beam_kernel<MAXC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)sk_entry_bytes(kSkMaxBeam, MAXC)));
        once.mark(dev);
    }
    if (dtype == HCTR_F32)
        ctc_prune_logsoftmax_kernel<float, MAXC><<<(int)rows, kPruneThreads, smem, s>>>(static_cast<const float*>(logits), T, B, C, stride_t,
                                                                                         stride_b, meta, blank, cidx, clp);
    else
        ctc_prune_logsoftmax_kernel<__nv_bfloat16, MAXC><<<(int)rows, kPruneThreads, smem, s>>>(
            static_cast<const __nv_bfloat16*>(logits), T, B, C, stride_t, stride_b, meta, blank, cidx, clp);
    HCTR_CUDA(cudaGetLastError());
    const size_t dyn = global_entries ? 0 : sk_entry_bytes(beam_size, MAXC);
    ctc_skip_beam_kernel<MAXC><<<B, kSkThreads, dyn, s>>>(meta, blank, cidx, clp, T, B, C, beam_size, lm_penalty, len_bonus, lm_table, lm,
                                                          out_idx, out_len, status, seq_ws, sk_ws_per_seq(T), entry_ws, entry_per_seq);
    HCTR_CUDA(cudaGetLastError());
    return HCTR_OK;
}

}  // namespace hctr

using namespace hctr;

extern "C" {

int hctr_ctc_skip_max_candidates(void) { return kSkBigC; }

long long hctr_ctc_skip_workspace_bytes_ex(int T, int B, int max_candidates) {
    if (T <= 0 || B <= 0) return 0;
    const int maxc = max_candidates > kSkMaxC ? kSkBigC : kSkMaxC;
    // per-row candidate tables + per-sequence greedy list and prefix trie (+ per-sequence dict entries for the large variant)
    const long long rows = (long long)T * B;
    long long tables = rows * (8 + 4 + (long long)maxc * 8);
    tables = (tables + 255) & ~255ll;
    long long entries = 0;
    if (maxc > kSkMaxC) entries = (((long long)sk_entry_bytes(kSkMaxBeam, maxc) + 255) & ~255ll) * B;
    return tables + ((sk_ws_per_seq(T) * B + 255) & ~255ll) + entries + 256;
}
long long hctr_ctc_skip_workspace_bytes(int T, int B) { return hctr_ctc_skip_workspace_bytes_ex(T, B, kSkMaxC); }

int hctr_ctc_skip_beam_search(const void* logits, int dtype, int T, int B, int C, long long stride_t, long long stride_b,
                              int beam_size, double lm_penalty, double len_bonus, const double* lm_table, int32_t* out_idx,
                              int32_t* out_len, int32_t* status, void* workspace, long long workspace_bytes, void* stream) {
    return hctr_ctc_skip_beam_search_ex(logits, dtype, T, B, C, stride_t, stride_b, beam_size, lm_penalty, len_bonus, lm_table,
                                        nullptr, kSkMaxC, out_idx, out_len, status, workspace, workspace_bytes, stream);
}

int hctr_ctc_skip_beam_search_lm(const void* logits, int dtype, int T, int B, int C, long long stride_t, long long stride_b,
                                 int beam_size, double lm_penalty, double len_bonus, const double* lm_table,
                                 const hctr_ngram_lm* ngram, int32_t* out_idx, int32_t* out_len, int32_t* status,
                                 void* workspace, long long workspace_bytes, void* stream) {
    return hctr_ctc_skip_beam_search_ex(logits, dtype, T, B, C, stride_t, stride_b, beam_size, lm_penalty, len_bonus, lm_table,
                                        ngram, kSkMaxC, out_idx, out_len, status, workspace, workspace_bytes, stream);
}

int hctr_ctc_skip_beam_search_ex(const void* logits, int dtype, int T, int B, int C, long long stride_t, long long stride_b,
                                 int beam_size, double lm_penalty, double len_bonus, const double* lm_table,
                                 const hctr_ngram_lm* ngram, int max_candidates, int32_t* out_idx, int32_t* out_len,
                                 int32_t* status, void* workspace, long long workspace_bytes, void* stream) {
    HCTR_CHECK(!(ngram && lm_table), HCTR_ERR_INVALID, "skip beam: pass either a unigram table or an n-gram model");
    hctr_ngram_lm lm;
    memset(&lm, 0, sizeof(lm));
    if (ngram) {
        int rc = hctr::check_ngram(ngram, "skip beam");
        if (rc) return rc;
        HCTR_CHECK(ngram->num_ids >= C, HCTR_ERR_INVALID, "skip beam: the n-gram vocabulary map covers %d ids, the logits have %d classes", ngram->num_ids, C);
        lm = *ngram;
    }
    HCTR_CHECK(out_idx && out_len && status, HCTR_ERR_INVALID, "skip beam: null output");
    HCTR_CHECK(dtype == HCTR_F32 || dtype == HCTR_BF16, HCTR_ERR_INVALID, "skip beam: bad dtype");
    HCTR_CHECK(beam_size >= 1 && beam_size <= kSkMaxBeam, HCTR_ERR_INVALID, "skip beam: beam size must be in [1,%d]", kSkMaxBeam);
    HCTR_CHECK(T >= 0 && B >= 0 && C > 1, HCTR_ERR_INVALID, "skip beam: bad shape");
    HCTR_CHECK(max_candidates == kSkMaxC || max_candidates == kSkBigC, HCTR_ERR_INVALID, "skip beam: max_candidates must be %d or %d", kSkMaxC, kSkBigC);
    if (B == 0) return HCTR_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (T == 0) {
        HCTR_CUDA(cudaMemsetAsync(out_len, 0, sizeof(int32_t) * B, s));
        HCTR_CUDA(cudaMemsetAsync(status, 0xff, sizeof(int32_t) * B, s));
        return HCTR_OK;
    }
    HCTR_CHECK(logits != nullptr, HCTR_ERR_INVALID, "skip beam: null logits");
    const long long need = hctr_ctc_skip_workspace_bytes_ex(T, B, max_candidates);
    HCTR_CHECK(workspace && workspace_bytes >= need, HCTR_ERR_INVALID, "skip beam: workspace too small (%lld < %lld)", workspace_bytes, need);
    HCTR_CHECK((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, HCTR_ERR_INVALID, "skip beam: workspace must be 256-byte aligned");
    const long long rows = (long long)T * B;
    HCTR_CHECK(rows < (1ll << 31) && (long long)T * kSkMaxBeam + 1 < (1ll << 31), HCTR_ERR_INVALID, "skip beam: too large");
    HCTR_CHECK((size_t)C * sizeof(float) <= 160 * 1024, HCTR_ERR_INVALID, "skip beam: %d classes do not fit the shared-memory row buffer", C);
    if (max_candidates == kSkBigC)
        return sk_launch<kSkBigC>(logits, dtype, T, B, C, stride_t, stride_b, beam_size, lm_penalty, len_bonus, lm_table, lm, out_idx,
                                  out_len, status, workspace, s);
    return sk_launch<kSkMaxC>(logits, dtype, T, B, C, stride_t, stride_b, beam_size, lm_penalty, len_bonus, lm_table, lm, out_idx,
                              out_len, status, workspace, s);
}

}  // extern "C"
