// CTC loss forward/backward fused with log-softmax (reference call: main.py:205,406-409,
//   criterion = CTCLoss(zero_infinity=True); loss = criterion(preds.log_softmax(2), targets, T*B, lengths)).
// Three passes:
//   A (HBM-bound): one warp per (t,b) row -> log-sum-exp, and the gathered log-probs of the blank-interleaved
//                  label sequence l' (S = 2L+1) for the recursion.
//   B (latency-bound): alpha and beta run CONCURRENTLY (grid (B,2)), each as ONE WARP per sequence holding the state in
//                  registers (K = 4..32 consecutive states per lane, neighbours by shuffle: no barrier, no shared-memory
//                  round trip). The recursion is evaluated in SCALED LINEAR space in fp64:
//                      a_t(s) = (a_{t-1}(s) + a_{t-1}(s-1) + [skip] a_{t-1}(s-2)) * p_t(l'_s) * 2^-d_t
//                  with an exact power-of-two rescale per step (d_t from the warp-wide maximum exponent two steps back,
//                  a dead-beat controller, so the reduction is off the dependency chain) and the per-step exponents
//                  kept as integers. No exp/log on the chain: a step is two shuffles, two adds and a multiply. The
//                  label probabilities are staged through shared memory with cp.async three chunks deep. beta is the
//                  same recursion on the reversed label sequence and reversed time.
//                  Linear space cannot hold states more than ~2^-700 below the row maximum; steps that drop such a
//                  state are marked, a verify pass proves the dropped paths carry < 2^-260 of the likelihood (the
//                  opposite recursion is not small where this one peaks) and otherwise - or when a label probability
//                  is below e^-80 - the sequence is recomputed by the log-space recursion (fp64 state, 3-way
//                  log-sum-exp as ATen's ctc_loss), which has no such limit.
//   C (HBM-bound): one CTA per row: grad = (softmax - occupancy) * scale written in one pass, where
//                  occupancy_c = sum_{s: l'_s = c} exp(alpha_t(s) + beta_t(s) - ll - lp[t, l'_s]).
// Logits are read twice and the gradient written once: (2*s_in + s_out) * T*B*C bytes.
//
// Round 2 - the ONE-PASS ROWS path (default whenever a logits row fits the registers of four warps, C <= 8192 fp32 / 16384
// bf16); its kernels run back to back on the caller's stream, or - opt-in, HCTR_CTC_OVERLAP=1, measured slower - overlapped:
//   rows  (HBM-bound, ctc_rows_kernel): ONE pass per logits row, the row held in the registers of four warps: max, exp
//         and sum, the label gather for the recursion, AND the dense part of the gradient softmax*scale - which needs no
//         alpha/beta. The logits are read once and the gradient written once: (s_in + s_out) * T*B*C bytes of DRAM
//         traffic instead of (2*s_in + s_out). Rows are dealt to the CTAs from BOTH ends of the time axis and every
//         finished row bumps a per-(sequence, 8-row granule) counter (release).
//   scan  (latency-bound, ctc_scan_kernel<K, POLL>): unchanged recursion; 16-byte cp.async.cg staging from the plain
//         (alpha) / mirrored (beta) probability rows. POLL = the overlapped schedule: launched on a helper stream at the
//         same time as the rows kernel, it waits on a granule's counter before staging its rows.
//   fix   (ctc_fix_kernel): one warp per row subtracts the occupancy at the <= L+1 distinct label classes (a few dozen
//         2-byte read-modify-writes per row; blank = a fixed warp tree, repeated labels = a precomputed chain).
// HCTR_CTC_OVERLAP=0 keeps the round-1 sequential passes A/B/C (also the fallback for rows that do not fit).
//
// Round 2 - the SPLIT schedules (default with a gradient and B <= 24; ctc_lse_chunk_kernel / ctc_gather_rel_kernel below):
//   a sparse label gather -> [scans on SMs of their own || log-sum-exp pass + dense gradient pass on a helper stream] ->
//   verify -> [log-space fallback] -> fix. In the RELATIVE form the label tables are relative to the row's largest label
//   logit instead of to the log-sum-exp (the recursions do not need normalised probabilities: every path takes one emission
//   per step, so the factor prod_t exp(lse_t - ref_t) cancels in the occupancy and is added back to the loss at the end);
//   with the classifier's row log-sum-exp passed in (the training step) the normalised form needs no full-row pass either.
#include <cfloat>
#include <cstdlib>
#include <mutex>

#include "common.cuh"
#include "row_stage.cuh"
#include "../../include/hctr_b200.h"

namespace hctr {

struct CtcWs {
    float* lse;        // [B][T]
    float* ref;        // [B][T]         relative schedule: the row's reference logit (max over the label set); lpg / pg / pgr
                       //                are then relative to it instead of to the log-sum-exp
    float* lpg;        // [B][T][Sp]     log-prob of l'_s at (t,b) (log-space fallback)
    double* pg;        // [B][T][Sp]     probability of l'_s at (t,b), zero beyond S (scaled linear recursion)
    double* pgr;       // [B][T][Sp]     the same row mirrored: pgr[q] = pg[S-1-q] (what the beta scan reads front to back)
    double* alpha;     // [B][T][Sa]     linear mode: alpha_t(s) / 2^ea[t][s/K];  log mode: log alpha_t(s)
    double* beta;      // [B][T][Sa]     linear mode: MIRRORED, beta[s'] = beta_t(S-1-s') / 2^eb[t][s'/K];  log mode: log beta_t(s)
    int* ea;           // [B][T][32]     per lane of the scan: (binary exponent << 1) | (a state was dropped at this step)
    int* eb;           // [B][T][32]     (lanes of the beta scan hold mirrored states s' = S-1-s)
    double* ll;        // [B]            log-likelihood (may be -inf), fp64
    double* pfin;      // [B]            linear mode: likelihood / 2^efin
    int* efin;         // [B]
    int* flag;         // [B]            1 = this sequence takes the log-space recursion
    int* canon;        // [B][Sp]        first s' with the same class as s
    int* nxt;          // [B][Sp]        next s' > s with the same class as s, -1 if none
    int* toff;         // [B]            offset of sequence b in the concatenated targets
    int* len;          // [B]            target length clamped to [0, max_target_len] (never trust device lengths)
    int* lab;          // [B][Lp]        labels clamped to [0, C-1]
    int* prog;         // [B][NG]        rows of granule g (8 time steps) whose label probabilities are in place
    int* err;          // [4]            bit 0: a length / label was out of range; bit 1: the scan timed out waiting for rows
    unsigned long long* dbg;   // [B][2][8]  HCTR_CTC_TIMING diagnostics of the polling scan: start, end, ns waiting, polls, quarter marks
    int Lp, NG;
};
constexpr int kGranule = 8;          // time steps per progress counter

__host__ __device__ inline long long align_up(long long v, long long a) { return (v + a - 1) / a * a; }

// row pitch of the per-state arrays: S = 2L+1 rounded up to a multiple of 4 (16-byte rows of floats)
__host__ __device__ inline int state_pitch(int max_target_len) { return (2 * max_target_len + 1 + 3) / 4 * 4; }
// states per lane of the one-warp scan (0: too many states, log-space recursion only) and the alpha/beta row pitch
inline int scan_lane_states(int Sp) { return Sp <= 128 ? 4 : Sp <= 256 ? 8 : Sp <= 512 ? 16 : 0; }
inline int alpha_pitch(int Sp) { const int k = scan_lane_states(Sp); return k ? 32 * k : Sp; }

static CtcWs carve(void* base, int T, int B, int Sp, int Sa, long long* total) {
    long long off = 0;
    auto take = [&](long long bytes) { long long o = off; off = align_up(off + bytes, 256); return o; };
    const long long o_lse = take(4ll * B * T);
    const long long o_ref = take(4ll * B * T);
    const long long o_lpg = take(4ll * B * T * Sp);
    const long long o_pg = take(8ll * B * T * Sp);
    const long long o_pgr = take(8ll * B * T * Sp + 64);
    const long long o_alpha = take(8ll * B * T * Sa);
    const long long o_beta = take(8ll * B * T * Sa);
    const long long o_ea = take(4ll * B * T * 32);
    const long long o_eb = take(4ll * B * T * 32);
    const long long o_ll = take(8ll * B);
    const long long o_pfin = take(8ll * B);
    const long long o_efin = take(4ll * B);
    const long long o_flag = take(4ll * B);
    const long long o_canon = take(4ll * B * Sp);
    const long long o_nxt = take(4ll * B * Sp);
    const long long o_toff = take(4ll * B);
    const int Lp = (Sp - 1) / 2 > 0 ? (Sp - 1) / 2 : 1;          // >= max_target_len (Sp >= 2L+1)
    const int NG = (T + kGranule - 1) / kGranule;
    const long long o_len = take(4ll * B);
    const long long o_lab = take(4ll * B * Lp);
    const long long o_prog = take(4ll * B * NG);
    const long long o_err = take(16);
    const long long o_dbg = take(128ll * B);
    if (total) *total = off;
    CtcWs w;
    char* p = static_cast<char*>(base);
    w.lse = reinterpret_cast<float*>(p + o_lse);
    w.ref = reinterpret_cast<float*>(p + o_ref);
    w.lpg = reinterpret_cast<float*>(p + o_lpg);
    w.pg = reinterpret_cast<double*>(p + o_pg);
    w.pgr = reinterpret_cast<double*>(p + o_pgr);
    w.alpha = reinterpret_cast<double*>(p + o_alpha);
    w.beta = reinterpret_cast<double*>(p + o_beta);
    w.ea = reinterpret_cast<int*>(p + o_ea);
    w.eb = reinterpret_cast<int*>(p + o_eb);
    w.ll = reinterpret_cast<double*>(p + o_ll);
    w.pfin = reinterpret_cast<double*>(p + o_pfin);
    w.efin = reinterpret_cast<int*>(p + o_efin);
    w.flag = reinterpret_cast<int*>(p + o_flag);
    w.canon = reinterpret_cast<int*>(p + o_canon);
    w.nxt = reinterpret_cast<int*>(p + o_nxt);
    w.toff = reinterpret_cast<int*>(p + o_toff);
    w.len = reinterpret_cast<int*>(p + o_len);
    w.lab = reinterpret_cast<int*>(p + o_lab);
    w.prog = reinterpret_cast<int*>(p + o_prog);
    w.err = reinterpret_cast<int*>(p + o_err);
    w.dbg = reinterpret_cast<unsigned long long*>(p + o_dbg);
    w.Lp = Lp; w.NG = NG;
    return w;
}

template <typename T> struct Ld;
template <> struct Ld<float> {
    static constexpr int N = 4;
    static __device__ __forceinline__ void vec(const float* p, float (&o)[4]) {
        const uint4 q = ld_nc_v4(p);
        o[0] = __uint_as_float(q.x); o[1] = __uint_as_float(q.y); o[2] = __uint_as_float(q.z); o[3] = __uint_as_float(q.w);
    }
    static __device__ __forceinline__ float one(const float* p) { return __ldg(p); }
    static __device__ __forceinline__ void st_vec(float* p, const float (&v)[4]) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
    static __device__ __forceinline__ void st_one(float* p, float v) { *p = v; }
};
template <> struct Ld<__nv_bfloat16> {
    static constexpr int N = 8;
    static __device__ __forceinline__ void vec(const __nv_bfloat16* p, float (&o)[8]) {
        const uint4 q = ld_nc_v4(p);
        o[0] = bf16_lo(q.x); o[1] = bf16_hi(q.x); o[2] = bf16_lo(q.y); o[3] = bf16_hi(q.y);
        o[4] = bf16_lo(q.z); o[5] = bf16_hi(q.z); o[6] = bf16_lo(q.w); o[7] = bf16_hi(q.w);
    }
    static __device__ __forceinline__ float one(const __nv_bfloat16* p) {
        return __uint_as_float(static_cast<uint32_t>(*reinterpret_cast<const unsigned short*>(p)) << 16);
    }
    static __device__ __forceinline__ void st_vec(__nv_bfloat16* p, const float (&v)[8]) {
        *reinterpret_cast<uint4*>(p) = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]),
                                                  pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
    }
    static __device__ __forceinline__ void st_one(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
};

// ---------------------------------------------------------------- prep: sanitised labels, canonical states, counters
// Lengths and labels live in device memory and are NOT trusted: every later kernel reads the clamped copies made here
// (length in [0, max_target_len], label in [0, C-1]), so a wrong length can no longer index outside the workspace or
// the logits row. Out-of-range input sets err bit 0 and the loss becomes NaN.
__global__ void ctc_prep_kernel(const int32_t* __restrict__ targets, const int32_t* __restrict__ tlen, int B, int C, int Sp,
                                int max_target_len, int force_log, CtcWs w) {
    __shared__ int s_off, s_len, s_bad;
    const int b = blockIdx.x;
    if (threadIdx.x == 0) {
        int off = 0, bad = 0;
        for (int i = 0; i < b; ++i) {
            const int li = tlen[i];
            bad |= (li < 0 || li > max_target_len);
            off += min(max(li, 0), max_target_len);
        }
        const int lb = tlen[b];
        bad |= (lb < 0 || lb > max_target_len);
        s_off = off; s_len = min(max(lb, 0), max_target_len); s_bad = bad;
        w.toff[b] = off; w.len[b] = s_len;
        w.flag[b] = force_log;              // 1: more states than the one-warp scan holds
    }
    __syncthreads();
    const int L = s_len, S = 2 * L + 1;
    const int32_t* tg = targets + s_off;
    int* lab = w.lab + (long long)b * w.Lp;
    int bad = 0;
    for (int i = threadIdx.x; i < L; i += blockDim.x) {
        const int c = tg[i];
        bad |= (c < 0 || c >= C);
        lab[i] = min(max(c, 0), C - 1);
    }
    if (bad || (threadIdx.x == 0 && s_bad)) atomicOr(&w.err[0], 1);
    for (int g = threadIdx.x; g < w.NG; g += blockDim.x) w.prog[(long long)b * w.NG + g] = 0;
    __syncthreads();
    for (int s = threadIdx.x; s < S; s += blockDim.x) {
        const int c = (s & 1) ? lab[s >> 1] : 0;
        int first = s;
        for (int q = (s & 1) ? 1 : 0; q < s; q += ((c == 0) ? 2 : 1)) {
            const int cq = (q & 1) ? lab[q >> 1] : 0;
            if (cq == c) { first = q; break; }
        }
        w.canon[(long long)b * Sp + s] = first;
        int next = -1;
        for (int q = s + ((c == 0) ? 2 : 1); q < S; q += ((c == 0) ? 2 : 1)) {
            const int cq = (q & 1) ? lab[q >> 1] : 0;
            if (cq == c) { next = q; break; }
        }
        w.nxt[(long long)b * Sp + s] = next;
    }
}

// label probabilities below this take the sequence to the log-space recursion (keeps p representable in fp32 and bounds
// the per-step shrink of the scaled recursion by 2^-116)
constexpr float kMinLinearLogProb = -80.f;

// ---------------------------------------------------------------- pass A: row log-sum-exp + label gather
constexpr int kLseWarps = 8;

template <typename T>
__global__ void __launch_bounds__(kLseWarps * 32)
ctc_lse_gather_kernel(const T* __restrict__ logits, int Tn, int Bn, int C, long long stride_t, long long stride_b,
                      const int32_t* __restrict__ targets, const int32_t* __restrict__ tlen,
                      const int32_t* __restrict__ ilen, int Sp, const float* __restrict__ lse_in, CtcWs w) {
    constexpr int V = Ld<T>::N;
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * kLseWarps + (threadIdx.x >> 5);     // row = b*T + t
    if (row >= (long long)Tn * Bn) return;
    const int b = (int)(row / Tn), t = (int)(row - (long long)b * Tn);
    if (t >= ilen[b]) return;
    const T* p = logits + (long long)t * stride_t + (long long)b * stride_b;
    float lse;
    if (lse_in != nullptr) {
        lse = lse_in[row];            // log-sum-exp already produced by the classifier epilogue: only gather the labels
    } else {
        float m = -FLT_MAX, s = 0.f;
        auto upd = [&](float x) {
            if (x > m) { s = s * __expf(m - x) + 1.f; m = x; } else { s += __expf(x - m); }
        };
        const uintptr_t addr = reinterpret_cast<uintptr_t>(p);
        int head = (int)(((16 - (addr & 15)) & 15) / sizeof(T));
        if (head > C) head = C;
        if (lane < head) upd(Ld<T>::one(p + lane));
        const int nvec = (C - head) / V;
        const T* pv = p + head;
        int vi = lane;
        for (; vi + 96 < nvec; vi += 128) {
            float x[4][V];
    #pragma unroll
            for (int u = 0; u < 4; ++u) Ld<T>::vec(pv + (long long)(vi + 32 * u) * V, x[u]);
    #pragma unroll
            for (int u = 0; u < 4; ++u) {
                float vm = x[u][0];
    #pragma unroll
                for (int j = 1; j < V; ++j) vm = fmaxf(vm, x[u][j]);
                if (vm > m) { s *= __expf(m - vm); m = vm; }
    #pragma unroll
                for (int j = 0; j < V; ++j) s += __expf(x[u][j] - m);
            }
        }
        for (; vi < nvec; vi += 32) {
            float x[V];
            Ld<T>::vec(pv + (long long)vi * V, x);
    #pragma unroll
            for (int j = 0; j < V; ++j) upd(x[j]);
        }
        const int tail0 = head + nvec * V;
        if (tail0 + lane < C) upd(Ld<T>::one(p + tail0 + lane));
        float mm = m;
    #pragma unroll
        for (int o = 16; o > 0; o >>= 1) mm = fmaxf(mm, __shfl_xor_sync(0xffffffffu, mm, o));
        float ss = s * __expf(m - mm);
    #pragma unroll
        for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        lse = mm + logf(ss);
    }
    if (lane == 0) w.lse[row] = lse;
    const int L = w.len[b], S = 2 * L + 1;
    const int* tg = w.lab + (long long)b * w.Lp;
    float* dst = w.lpg + row * Sp;
    double* dpr = w.pg + row * Sp;
    double* dpm = w.pgr + row * Sp;
    bool small = false;
    for (int q = lane; q < Sp; q += 32) {
        float lp = 0.f;
        double pr = 0.0;
        if (q < S) {
            const int c = (q & 1) ? tg[q >> 1] : 0;
            lp = Ld<T>::one(p + c) - lse;
            pr = (double)expf(lp);
            small |= (lp < kMinLinearLogProb) && (lp > -INFINITY);       // exp(-inf) = 0 is exact in linear space
            small |= !(lp == lp);                                        // NaN: let the log-space path propagate it
            dpm[S - 1 - q] = pr;
        } else {
            dpm[q] = 0.0;
        }
        dst[q] = lp;
        dpr[q] = pr;
    }
    if (__any_sync(0xffffffffu, small) && lane == 0) w.flag[b] = 1;
}

// ---------------------------------------------------------------- split schedule: pass A' (log-sum-exp + gather, one read)
// Round 2, second half. The schedule "A' -> [scans || softmax*scale] -> fix": the scans (2*B one-warp CTAs, 0.24 ms of pure
// latency at T = 2048 whatever B is) need nothing but the log-sum-exp and the label probabilities of every row, and the dense
// part of the gradient needs nothing from the scans - so the dense part runs on a helper stream UNDERNEATH them, by plain
// stream fork / join: no progress counters, no polling, correct under tools that serialise kernels. It costs one more read
// of the logits than the one-pass rows kernel ((2*s_in + s_out) * T*B*C bytes, SURVEY 8d's figure), and both of its passes
// are simpler kernels that run closer to the HBM peak than the rows kernel's 65 %.
//   A'  one warp per row, the row read ONCE in register chunks of kLseChunkVec 16-byte vectors per lane (the top-k kernel's
//       scheme, ctc_beam.cu): per chunk the lane's packed maximum, a lane-private online rescale of its sum when the maximum
//       grows (predicated, no cross-lane traffic), FFMA + ex2 + FADD per element; one warp reduction at the end; then the
//       label gather of ctc_lse_gather_kernel (the labels' logits were read microseconds ago: L2 hits).
constexpr int kLseChunkVec = 8;

template <typename T, bool GATHER = true>
__global__ void __launch_bounds__(kLseWarps * 32, 4)
ctc_lse_chunk_kernel(const T* __restrict__ logits, int Tn, int Bn, int C, long long stride_t, long long stride_b,
                     const int32_t* __restrict__ ilen, int Sp, const float* __restrict__ lse_in, CtcWs w) {
    constexpr int V = Ld<T>::N;
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * kLseWarps + (threadIdx.x >> 5);     // row = b*T + t
    if (row >= (long long)Tn * Bn) return;
    const int b = (int)((unsigned)row / (unsigned)Tn), t = (int)((unsigned)row - (unsigned)b * (unsigned)Tn);
    if (t >= ilen[b]) return;
    const T* p = logits + (long long)t * stride_t + (long long)b * stride_b;
    float lse;
    if (lse_in != nullptr) {
        lse = lse_in[row];            // log-sum-exp already produced by the classifier epilogue: only gather the labels
    } else {
        const uintptr_t addr = reinterpret_cast<uintptr_t>(p);
        int head = (int)(((16 - (addr & 15)) & 15) / sizeof(T));
        if (head > C) head = C;
        const int nvec = (C - head) / V;
        const int tail0 = head + nvec * V;
        const T* pv = p + head;
        float x_h = -INFINITY, x_t = -INFINITY;
        if (lane < head) x_h = Ld<T>::one(p + lane);
        if (tail0 + lane < C) x_t = Ld<T>::one(p + tail0 + lane);
        const uint32_t ninf_w = sizeof(T) == 2 ? 0xFF80FF80u : 0xFF800000u;
        float m = -INFINITY;                                                 // lane-private running maximum
        float s4[4] = {0.f, 0.f, 0.f, 0.f};                                  // lane-private sums of exp(x - m)
        for (int base = 0; base < nvec; base += 32 * kLseChunkVec) {
            uint4 q[kLseChunkVec];
            if (base + 32 * kLseChunkVec <= nvec) {
#pragma unroll
                for (int u = 0; u < kLseChunkVec; ++u) q[u] = ld_nc_v4(pv + (long long)(base + 32 * u + lane) * V);
            } else {
#pragma unroll
                for (int u = 0; u < kLseChunkVec; ++u) {
                    const int vi = base + 32 * u + lane;
                    q[u] = make_uint4(ninf_w, ninf_w, ninf_w, ninf_w);
                    if (vi < nvec) q[u] = ld_nc_v4(pv + (long long)vi * V);
                }
            }
            float cm;
            if (sizeof(T) == 2) {
                __nv_bfloat162 a[kLseChunkVec];
#pragma unroll
                for (int u = 0; u < kLseChunkVec; ++u)
                    a[u] = __hmax2(__hmax2(*reinterpret_cast<const __nv_bfloat162*>(&q[u].x), *reinterpret_cast<const __nv_bfloat162*>(&q[u].y)),
                                   __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&q[u].z), *reinterpret_cast<const __nv_bfloat162*>(&q[u].w)));
#pragma unroll
                for (int h = kLseChunkVec / 2; h > 0; h >>= 1) {
#pragma unroll
                    for (int u = 0; u < h; ++u) a[u] = __hmax2(a[u], a[u + h]);
                }
                cm = fmaxf(__low2float(a[0]), __high2float(a[0]));
            } else {
                float a[kLseChunkVec];
#pragma unroll
                for (int u = 0; u < kLseChunkVec; ++u)
                    a[u] = fmaxf(fmaxf(__uint_as_float(q[u].x), __uint_as_float(q[u].y)), fmaxf(__uint_as_float(q[u].z), __uint_as_float(q[u].w)));
#pragma unroll
                for (int h = kLseChunkVec / 2; h > 0; h >>= 1) {
#pragma unroll
                    for (int u = 0; u < h; ++u) a[u] = fmaxf(a[u], a[u + h]);
                }
                cm = a[0];
            }
            if (cm > m) {                                                    // (m = -inf: the sums are still zero)
                const float sc = ex2_fast((m - cm) * kLog2e);
                s4[0] *= sc; s4[1] *= sc; s4[2] *= sc; s4[3] *= sc;
                m = cm;
            }
            const float nml = m > -INFINITY ? -m * kLog2e : 0.f;            // nothing but -inf so far: every term is 2^-inf = 0
            const int rem = nvec - base;                                     // warp-uniform: vector step u has data iff 32 u < rem
#pragma unroll
            for (int u = 0; u < kLseChunkVec; ++u) {
                if (32 * u >= rem) break;
                float x[V];
                if (sizeof(T) == 2) {
                    x[0] = bf16_lo(q[u].x); x[1] = bf16_hi(q[u].x); x[2] = bf16_lo(q[u].y); x[3] = bf16_hi(q[u].y);
                    x[V - 4] = bf16_lo(q[u].z); x[V - 3] = bf16_hi(q[u].z); x[V - 2] = bf16_lo(q[u].w); x[V - 1] = bf16_hi(q[u].w);
                } else {
                    x[0] = __uint_as_float(q[u].x); x[1] = __uint_as_float(q[u].y); x[2] = __uint_as_float(q[u].z); x[3] = __uint_as_float(q[u].w);
                }
                float acc = ex2_fast(fmaf(x[0], kLog2e, nml)) + ex2_fast(fmaf(x[1], kLog2e, nml));
#pragma unroll
                for (int j = 2; j < V; j += 2) acc += ex2_fast(fmaf(x[j], kLog2e, nml)) + ex2_fast(fmaf(x[j + 1], kLog2e, nml));
                s4[u & 3] += acc;
            }
        }
        // ---- the scalars in front of / behind the interior, then across the warp
        const float xs = fmaxf(x_h, x_t);
        float ssum = (s4[0] + s4[1]) + (s4[2] + s4[3]);
        if (xs > m) { ssum *= ex2_fast((m - xs) * kLog2e); m = xs; }
        if (m > -INFINITY) ssum += ex2_fast((x_h - m) * kLog2e) + ex2_fast((x_t - m) * kLog2e);
        const float mm = warp_max(m);
        ssum = warp_sum(m > -INFINITY ? ssum * ex2_fast((m - mm) * kLog2e) : 0.f);
        lse = mm + logf(ssum);
    }
    if (lane == 0) w.lse[row] = lse;
    if (!GATHER) return;                              // relative schedule: the labels were gathered before the scans started
    const int L = w.len[b], S = 2 * L + 1;
    const int* tg = w.lab + (long long)b * w.Lp;
    float* dst = w.lpg + row * Sp;
    double* dpr = w.pg + row * Sp;
    double* dpm = w.pgr + row * Sp;
    bool small = false;
    for (int q = lane; q < Sp; q += 32) {
        float lp = 0.f;
        double pr = 0.0;
        if (q < S) {
            const int c = (q & 1) ? tg[q >> 1] : 0;
            lp = Ld<T>::one(p + c) - lse;
            pr = (double)expf(lp);
            small |= (lp < kMinLinearLogProb) && (lp > -INFINITY);       // exp(-inf) = 0 is exact in linear space
            small |= !(lp == lp);                                        // NaN: let the log-space path propagate it
            dpm[S - 1 - q] = pr;
        } else {
            dpm[q] = 0.0;
        }
        dst[q] = lp;
        dpr[q] = pr;
    }
    if (__any_sync(0xffffffffu, small) && lane == 0) w.flag[b] = 1;
}

// ---------------------------------------------------------------- relative schedule: label gather without the log-sum-exp
// Every alignment path takes exactly one emission per time step, so the recursions do not need NORMALISED label probabilities:
// with p'_t(s) = exp(x_t(l'_s) - r_t) for ANY per-row reference r_t, alpha', beta' and the likelihood L' all carry the same
// factor prod_t exp(lse_t - r_t), the state occupancy alpha' beta' / (p' L') is unchanged, and
//     log L = log L' + sum_t (r_t - lse_t).
// With r_t = the largest logit of the row's label set (blank and labels) the gather needs a few dozen scattered reads per row
// instead of the whole row, p' <= 1 with equality somewhere (the block-floating-point scan's assumptions hold as before), and
// the scans can start at once - the full-row log-sum-exp pass and the dense gradient pass both run underneath them on the
// helper stream. One warp per row; lpg / pg / pgr hold the relative values, w.ref the reference.
template <typename T>
__global__ void __launch_bounds__(kLseWarps * 32)
ctc_gather_rel_kernel(const T* __restrict__ logits, int Tn, int Bn, long long stride_t, long long stride_b,
                      const int32_t* __restrict__ ilen, int Sp, CtcWs w) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * kLseWarps + (threadIdx.x >> 5);     // row = b*T + t
    if (row >= (long long)Tn * Bn) return;
    const int b = (int)((unsigned)row / (unsigned)Tn), t = (int)((unsigned)row - (unsigned)b * (unsigned)Tn);
    if (t >= ilen[b]) return;
    const T* p = logits + (long long)t * stride_t + (long long)b * stride_b;
    const int L = w.len[b], S = 2 * L + 1;
    const int* tg = w.lab + (long long)b * w.Lp;
    float m = -INFINITY;
    bool nan = false;
    for (int q = lane; q < S; q += 32) {
        const float x = Ld<T>::one(p + ((q & 1) ? tg[q >> 1] : 0));
        nan |= !(x == x);
        m = fmaxf(m, x);
    }
    m = warp_max(m);
    if (lane == 0) w.ref[row] = m;
    float* dst = w.lpg + row * Sp;
    double* dpr = w.pg + row * Sp;
    double* dpm = w.pgr + row * Sp;
    bool small = nan;                                                    // NaN: let the log-space path propagate it
    for (int q = lane; q < Sp; q += 32) {
        float lp = 0.f;
        double pr = 0.0;
        if (q < S) {
            const float x = Ld<T>::one(p + ((q & 1) ? tg[q >> 1] : 0));
            lp = (m == -INFINITY) ? -INFINITY : x - m;                   // every label impossible: all zeros, L' = 0
            pr = (double)expf(lp);
            small |= (lp < kMinLinearLogProb) && (lp > -INFINITY);       // exp(-inf) = 0 is exact in linear space
            small |= !(lp == lp);
            dpm[S - 1 - q] = pr;
        } else {
            dpm[q] = 0.0;
        }
        dst[q] = lp;
        dpr[q] = pr;
    }
    if (__any_sync(0xffffffffu, small) && lane == 0) w.flag[b] = 1;
}

// nll_b = -(log L'_b + sum_t (ref_t - lse_t)) for the relative schedule; runs after the scans / the log-space fallback wrote
// nll_b = -log L'_b and w.ll[b] = log L'_b (which stays relative: the gradient kernels use it with the relative tables).
__global__ void __launch_bounds__(256)
ctc_nll_rel_kernel(const int32_t* __restrict__ ilen, int Tn, float* __restrict__ nll_out, CtcWs w) {
    __shared__ double part[256];
    const int b = blockIdx.x;
    const int Tb = min(ilen[b], Tn);
    const float* lse = w.lse + (long long)b * Tn;
    const float* ref = w.ref + (long long)b * Tn;
    double acc = 0.0;
    for (int t = threadIdx.x; t < Tb; t += 256) acc += (double)lse[t] - (double)ref[t];
    part[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const double ll = w.ll[b];
        if (ll > -INFINITY && Tb > 0) {                                  // (-inf: nll stays 0, zero_infinity; NaN propagates)
            float n = (float)(-ll + part[0]);
            if (!(n < INFINITY) && n == n) n = 0.f;
            nll_out[b] = n;
        }
    }
}

// ---------------------------------------------------------------- split schedule: the dense part of the gradient
// grad = softmax * scale = 2^(x log2 e - lse log2 e) * scale for every class (the label classes are corrected by
// ctc_fix_kernel once the scans are done), zeros for the rows beyond a sequence's input length. A pure stream: one warp per
// row, four 16-byte loads in flight per lane, FFMA + ex2 + FMUL per element. Runs on the helper stream beside the scans.
constexpr int kDenseWarps = 8;

template <typename T>
__global__ void __launch_bounds__(kDenseWarps * 32)
ctc_dense_grad_kernel(const T* __restrict__ logits, T* __restrict__ grad, int Tn, int Bn, int C, long long stride_t,
                      long long stride_b, const int32_t* __restrict__ ilen, float grad_scale, CtcWs w) {
    constexpr int V = Ld<T>::N;
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * kDenseWarps + (threadIdx.x >> 5);   // row = b*T + t
    if (row >= (long long)Tn * Bn) return;
    const int b = (int)((unsigned)row / (unsigned)Tn), t = (int)((unsigned)row - (unsigned)b * (unsigned)Tn);
    const T* p = logits + (long long)t * stride_t + (long long)b * stride_b;
    T* g = grad + (long long)t * stride_t + (long long)b * stride_b;
    const bool live = t < ilen[b];
    const float scale = grad_scale / ((float)max(w.len[b], 1) * (float)Bn);
    const float nl = live ? -w.lse[row] * kLog2e : 0.f;
    const float mul = live ? scale : 0.f;
    auto val = [&](float x) { return live ? ex2_fast(fmaf(x, kLog2e, nl)) * mul : 0.f; };
    const uintptr_t addr = reinterpret_cast<uintptr_t>(p);
    if ((reinterpret_cast<uintptr_t>(g) & 15) != (addr & 15)) {                        // never with equal strides and aligned bases
        for (int c = lane; c < C; c += 32) Ld<T>::st_one(g + c, val(live ? Ld<T>::one(p + c) : 0.f));
        return;
    }
    int head = (int)(((16 - (addr & 15)) & 15) / sizeof(T));
    if (head > C) head = C;
    const int nvec = (C - head) / V;
    const int tail0 = head + nvec * V;
    if (lane < head) Ld<T>::st_one(g + lane, val(live ? Ld<T>::one(p + lane) : 0.f));
    if (tail0 + lane < C) Ld<T>::st_one(g + tail0 + lane, val(live ? Ld<T>::one(p + tail0 + lane) : 0.f));
    const T* pv = p + head;
    T* gv = g + head;
    if (!live) {
        for (int vi = lane; vi < nvec; vi += 32) *reinterpret_cast<uint4*>(gv + (long long)vi * V) = make_uint4(0u, 0u, 0u, 0u);
        return;
    }
    int vi = lane;
    for (; vi + 96 < nvec; vi += 128) {
        float x[4][V];
#pragma unroll
        for (int u = 0; u < 4; ++u) Ld<T>::vec(pv + (long long)(vi + 32 * u) * V, x[u]);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int j = 0; j < V; ++j) x[u][j] = ex2_fast(fmaf(x[u][j], kLog2e, nl)) * mul;
            Ld<T>::st_vec(gv + (long long)(vi + 32 * u) * V, x[u]);
        }
    }
    for (; vi < nvec; vi += 32) {
        float x[V];
        Ld<T>::vec(pv + (long long)vi * V, x);
#pragma unroll
        for (int j = 0; j < V; ++j) x[j] = ex2_fast(fmaf(x[j], kLog2e, nl)) * mul;
        Ld<T>::st_vec(gv + (long long)vi * V, x);
    }
}

// ---------------------------------------------------------------- pass B (fast): scaled linear recursion, one warp
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

constexpr int kTinyExp = -760;                       // a state this far below its lane's scale is about to be lost: it
                                                     // shrinks by at most 2^-240 per pair of steps, so it is still exact
                                                     // when the exponent step sees it
constexpr int kTinyHi = (1023 + kTinyExp) << 20;     // high word of 2^kTinyExp
constexpr int kRebaseDiff = 900;                     // adopt the left neighbour's exponent when it is this far above ours
constexpr int kScanChunkStates = 32;                 // steps per staged chunk x states per lane
constexpr int kScanChunkBytes = kScanChunkStates * 32 * 8;       // one staged chunk = 8 KB; NST chunks in flight

// Progress counters are polled with RELAXED gpu-scope loads (served by the L2, the point of coherence): an acquire load
// keeps every later memory operation of the warp - the cp.async of the next chunk, the alpha stores - behind its own L2
// round trip, once per granule, on the critical path of a latency-bound recursion (measured: the overlapped call 1.38 ms
// against 0.57 ms for the same kernels back to back). Ordering comes from the producer (stores, __threadfence, atomicAdd:
// the row is performed at the L2 before the counter moves) and, on this side, from the dependency chain counter value ->
// warp-uniform branch -> cp.async.cg / ld.global.cg of the row, all of which read the L2 and none of which can be issued
// before the counter value has arrived.
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
constexpr unsigned long long kScanWaitLimitNs = 1000000000ull;    // a protocol bug must end the kernel, not hang the GPU

__device__ __forceinline__ double pow2_clamped(int x) {          // 2^x, exact, x clamped to the normal range
    x = max(-1022, min(1023, x));
    return __hiloint2double((1023 + x) << 20, 0);
}

// grid (B, 2), 32 threads: blockIdx.y == 0 runs alpha (and the log-likelihood), 1 runs beta as the alpha recursion of
// the reversed problem (state s' = S-1-s, time i = Tb-1-t; stored mirrored). Lane l owns states l*K .. l*K+K-1 and carries
// its own binary exponent e (true value = v * 2^e, "block floating point"): neighbouring states stay within a bounded
// factor of each other, states far apart do not, so the dynamic range across lanes is unlimited like in log space.
// Steps come in pairs: the first of a pair handles exponents (adopt / compare with the left neighbour, rescale by a
// power of two, look for dropped states), the second is the bare recursion.
// POLL (overlapped path): the label probabilities are being produced by ctc_rows_kernel while this kernel runs; before
// the rows of a granule are staged, lane 0 acquires the granule's progress counter (prefetched one granule ahead so the
// L2 round trip is off the recursion's critical path) and the warp synchronises on the answer.
// NST = staged chunks in flight (dynamic shared memory, NST * 8 KB): 3 (24 steps ahead at K = 4, enough on a quiet device);
// kScanDeepStages when the dense gradient pass streams beside the scans and a DRAM round trip takes several microseconds.
template <int K, bool POLL, int NST>
__global__ void __launch_bounds__(32)
ctc_scan_kernel(const int32_t* __restrict__ ilen, int Tn, int Sp, float* __restrict__ nll_out, CtcWs w) {
    constexpr int CH = kScanChunkStates / K;          // steps per staged chunk (even)
    constexpr int SA = 32 * K;                        // alpha/beta row pitch: every lane's states exist in memory
    constexpr int KP = K / 2;                         // 16-byte pairs of states per lane
    static_assert(CH % 2 == 0 && K % 4 == 0, "pairs of steps, 32-byte aligned runs of states");
    extern __shared__ __align__(16) double stage[];                   // [NST][kScanChunkStates * 32]
    __shared__ double fin_v[2];
    __shared__ int fin_e[2];
    const int b = blockIdx.x, lane = threadIdx.x;
    const bool rev = blockIdx.y == 1;
    if (!POLL && w.flag[b]) return;                   // already sent to the log-space recursion
    const int L = w.len[b], S = 2 * L + 1;
    const int Tb = min(ilen[b], Tn);
    if (Tb <= 0) {
        if (!rev && lane == 0) {
            const double ll = (L == 0) ? 0.0 : -INFINITY;
            w.ll[b] = ll; w.pfin[b] = 1.0; w.efin[b] = 0;
            nll_out[b] = 0.f;                         // -0 or inf -> 0 (zero_infinity)
        }
        return;
    }
    const int* tg = w.lab + (long long)b * w.Lp;
    const int s0 = lane * K;
    // transition s'-2 -> s' is allowed iff l''_{s'} is a label different from l''_{s'-2} (same test in either direction);
    // kept as 0/1 multipliers so that the recursion is fused multiply-adds without selects
    double skd[K];
#pragma unroll
    for (int j = 0; j < K; ++j) {
        const int sp = s0 + j;
        skd[j] = 0.0;
        if (sp >= 2 && sp < S) {
            const int s = rev ? S - 1 - sp : sp;
            const int s2 = rev ? s + 2 : s - 2;
            if ((s & 1) && tg[s >> 1] != tg[s2 >> 1]) skd[j] = 1.0;
        }
    }
    // alpha reads pg, beta reads the mirrored copy pgr: either way state s' of this scan is element s' of the row
    const double* P = (rev ? w.pgr : w.pg) + (long long)b * Tn * Sp;
    const long long tstep = rev ? -1 : 1;             // rows advance forwards for alpha, backwards for beta
    const int t_first = rev ? Tb - 1 : 0;
    double* orow = (rev ? w.beta : w.alpha) + ((long long)b * Tn + t_first) * SA + s0;
    int* erow = (rev ? w.eb : w.ea) + ((long long)b * Tn + t_first) * 32 + lane;

    const int nsteps = Tb - 1;                        // steps i = 1 .. Tb-1; step i reads row t_first + i*tstep
    const int nchunks = (nsteps + CH - 1) / CH;

    // ---- progress protocol (POLL): steps i < ready have their rows in place
    const int* prog = w.prog + (long long)b * w.NG;
    int ready = POLL ? 0 : 0x7fffffff;
    int gnext = rev ? (Tb - 1) / kGranule : 0;        // next granule to acquire: upwards for alpha, downwards for beta
    int pend = 0;                                     // lane 0: prefetched counter of granule gnext
    bool timed_out = false;
    unsigned long long dbg_t0 = 0, dbg_wait = 0, dbg_polls = 0;
    if (POLL && lane == 0) dbg_t0 = global_timer_ns();
    if (POLL && lane == 0) pend = ld_acquire_gpu(prog + gnext);
    auto ensure = [&](int i_hi) {                     // warp-uniform: rows of steps 0..i_hi are complete on return
        while (ready <= i_hi && !timed_out) {
            int ok = 1;
            if (lane == 0) {
                const int target = min(kGranule, Tn - gnext * kGranule);
                int cnt = pend;
                if (cnt < target) {
                    const unsigned long long t0 = global_timer_ns();
                    do {
                        __nanosleep(100);
                        cnt = ld_acquire_gpu(prog + gnext);
                        ++dbg_polls;
                        if (cnt < target && global_timer_ns() - t0 > kScanWaitLimitNs) { ok = 0; break; }
                    } while (cnt < target);
                    dbg_wait += global_timer_ns() - t0;
                }
            }
            ok = __shfl_sync(0xffffffffu, ok, 0);     // the other lanes' reads are ordered after lane 0's acquire
            if (!ok) { timed_out = true; break; }
            ready = rev ? Tb - gnext * kGranule : (gnext + 1) * kGranule;
            gnext += rev ? -1 : 1;
            if (lane == 0 && gnext >= 0 && gnext < w.NG) pend = ld_acquire_gpu(prog + gnext);
        }
    };

    // States beyond Sp are zero-filled (0 source bytes; the address still lies inside the workspace); states in [S, Sp) are
    // zeros in memory. Rows past the last step are clamped to the last row (copied, unused).
    const double* isrc = P + (long long)t_first * Sp + s0;                // row of the next step to be staged
    const long long istride = tstep * Sp;
    int ileft = nsteps;                                                    // steps not yet staged
    int istaged = 0;                                                       // last step whose row has been requested
    int nbytes[KP];
#pragma unroll
    for (int jp = 0; jp < KP; ++jp) nbytes[jp] = (s0 + 2 * jp < Sp) ? 16 : 0;
    const uint32_t stage_u32 = smem_u32(stage) + lane * 16;
    auto issue = [&](int slot) {
        if (POLL) {
            const int hi = min(istaged + CH, nsteps);
            ensure(hi);
            istaged = hi;
        }
        const uint32_t dst = stage_u32 + slot * kScanChunkBytes;
#pragma unroll
        for (int u = 0; u < CH; ++u) {
            if (ileft > 0) { isrc += istride; --ileft; }
#pragma unroll
            for (int jp = 0; jp < KP; ++jp)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;"
                             :: "r"(dst + (u * KP + jp) * 512), "l"(isrc + 2 * jp), "r"(nbytes[jp]) : "memory");
        }
        cp_async_commit();
    };
    if (POLL) ensure(0);
#pragma unroll 1
    for (int i = 0; i < NST - 1; ++i) issue(i);

    // ---- i = 0
    double v[K];
    bool empty = true;                                // all of this lane's states are exactly zero
    {
        const double* src = P + (long long)t_first * Sp + s0;
#pragma unroll
        for (int j = 0; j < K; ++j) {
            const int sp = s0 + j;
            v[j] = (sp < 2 && sp < S) ? __ldcg(src + j) : 0.0;
            empty &= (v[j] == 0.0);
        }
#pragma unroll
        for (int j = 0; j < K; j += 2) *reinterpret_cast<double2*>(orow + j) = make_double2(v[j], v[j + 1]);
        *erow = 0;
    }
    int e = 0;                                        // true value = v * 2^e
    int d_next = 0, d_after = 0;                      // rescale exponents for the next two exponent steps
    int slot = 0;
    for (int c = 0; c < nchunks; ++c) {
        issue(slot == 0 ? NST - 1 : slot - 1);
        if (POLL && timed_out) break;                 // warp-uniform
        if (POLL && lane == 0 && (c == nchunks / 4 || c == nchunks / 2 || c == 3 * nchunks / 4))
            w.dbg[((long long)b * 2 + (rev ? 1 : 0)) * 8 + 4 + (c == nchunks / 4 ? 0 : c == nchunks / 2 ? 1 : 2)] = global_timer_ns();
        cp_async_wait<NST - 1>();                     // chunk c has landed (each lane reads only what it copied itself)
        const double2* buf = reinterpret_cast<const double2*>(stage) + slot * (kScanChunkStates * 16) + lane;
        slot = slot == NST - 1 ? 0 : slot + 1;
#pragma unroll
        for (int u = 0; u < CH; ++u) {
            const int i = 1 + c * CH + u;
            if (i > nsteps) break;                    // warp-uniform
            orow += tstep * SA;
            erow += tstep * 32;
            // ---- the two states to the left live in lane-1, in its units
            const double n1 = __shfl_up_sync(0xffffffffu, v[K - 1], 1);
            const double n2 = __shfl_up_sync(0xffffffffu, v[K - 2], 1);
            const int pk = __shfl_up_sync(0xffffffffu, (e << 1) | (empty ? 1 : 0), 1);
            const bool nempty = (pk & 1) != 0 || lane == 0;
            const int ne = pk >> 1;
            double nv[K];
            if ((u & 1) == 0) {
                // ---- exponent step
                bool drop = false;
                if (empty) { e = ne; d_next = 0; d_after = 0; }           // nothing here yet: take the neighbour's units
                int diff = ne - e;
                if (!nempty && diff > kRebaseDiff) {
                    // what arrives is > 2^400 above anything this lane holds (mantissas span 2^-464 .. 2^4): re-express the
                    // lane in the neighbour's units. What underflows here is old mass of states that receive the arriving
                    // mass within two steps - same states, same beta - so it is a relative loss below 2^-280 of those
                    // states' own paths, whatever beta is: no verification needed.
                    const double rr = pow2_clamped(-diff);
#pragma unroll
                    for (int j = 0; j < K; ++j) v[j] = (diff > 1022) ? 0.0 : v[j] * rr;
                    e = ne; d_next = 0; d_after = 0; diff = 0;
                }
                const double r = nempty ? 0.0 : pow2_clamped(diff);
                const int d = d_next;
                const double sc = pow2_clamped(-d);
                e += d;
                int mh = 0;
                double2 pp[KP];
#pragma unroll
                for (int jp = 0; jp < KP; ++jp) pp[jp] = buf[(u * KP + jp) * 32];
#pragma unroll
                for (int j = 0; j < K; ++j) {
                    const double ps = ((j & 1) ? pp[j >> 1].y : pp[j >> 1].x) * sc;
                    double a;
                    if (j == 0)      a = fma(fma(n2, skd[0], n1), r, v[0]);
                    else if (j == 1) a = fma(n1 * skd[1], r, v[1] + v[0]);
                    else             a = fma(v[j - 2], skd[j], v[j] + v[j - 1]);
                    nv[j] = a * ps;
                    const int hi = __double2hiint(nv[j]);
                    mh = max(mh, hi);
                    drop |= static_cast<unsigned>(hi - 1) < static_cast<unsigned>(kTinyHi - 1);   // 0 < value < 2^kTinyExp
                }
                empty = (mh == 0);
                // dead-beat rescale with a lag of two exponent steps: the maximum seen now, minus the correction
                // that is already scheduled
                const int mexp = mh ? ((mh >> 20) & 0x7ff) - 1023 : 0;
                d_next = d_after;
                d_after = max(-1000, min(1000, mexp - d_next));
                *erow = (e << 1) | (drop ? 1 : 0);
            } else {
                // ---- bare step: this lane keeps its units
                const double r = nempty ? 0.0 : pow2_clamped(ne - e);
                double2 pp[KP];
#pragma unroll
                for (int jp = 0; jp < KP; ++jp) pp[jp] = buf[(u * KP + jp) * 32];
#pragma unroll
                for (int j = 0; j < K; ++j) {
                    const double ps = (j & 1) ? pp[j >> 1].y : pp[j >> 1].x;
                    double a;
                    if (j == 0)      a = fma(fma(n2, skd[0], n1), r, v[0]);
                    else if (j == 1) a = fma(n1 * skd[1], r, v[1] + v[0]);
                    else             a = fma(v[j - 2], skd[j], v[j] + v[j - 1]);
                    nv[j] = a * ps;
                }
                empty = empty && (nv[0] == 0.0) && (nv[1] == 0.0);      // states >= 2 cannot fill before 0 and 1
                *erow = e << 1;
            }
#pragma unroll
            for (int j = 0; j < K; ++j) v[j] = nv[j];
#pragma unroll
            for (int j = 0; j < K; j += 2) *reinterpret_cast<double2*>(orow + j) = make_double2(nv[j], nv[j + 1]);
        }
    }
    cp_async_wait<0>();
    if (POLL && lane == 0) {
        unsigned long long* d = w.dbg + ((long long)b * 2 + (rev ? 1 : 0)) * 8;
        d[0] = dbg_t0; d[1] = global_timer_ns(); d[2] = dbg_wait; d[3] = dbg_polls;
        unsigned smid; asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
        d[7] = smid;
    }
    if (POLL && timed_out) {
        // the rows never arrived (protocol bug, or the two kernels could not run concurrently): report, do not hang
        if (lane == 0) {
            atomicOr(&w.err[0], 2);
            if (!rev) { w.ll[b] = NAN; w.pfin[b] = 1.0; w.efin[b] = 0; nll_out[b] = NAN; }
        }
        return;
    }
    if (!rev) {
        // likelihood = alpha(S-1) + alpha(S-2) at the last step, each in its lane's units
#pragma unroll
        for (int j = 0; j < K; ++j) {
            const int sp = s0 + j;
            if (sp == S - 1) { fin_v[0] = v[j]; fin_e[0] = e; }
            if (sp == S - 2) { fin_v[1] = v[j]; fin_e[1] = e; }
        }
        __syncwarp();
        if (lane == 0) {
            const bool two = S > 1;
            const int em = two ? max(fin_e[0], fin_e[1]) : fin_e[0];
            const double pf = fin_v[0] * pow2_clamped(fin_e[0] - em) + (two ? fin_v[1] * pow2_clamped(fin_e[1] - em) : 0.0);
            if (!(pf > 0.0)) {
                // zero (infeasible alignment, or everything dropped) or NaN: the log-space recursion decides
                w.flag[b] = 1;
            } else {
                const double ll = log(pf) + (double)em * 0.693147180559945309417;
                w.ll[b] = ll; w.pfin[b] = pf; w.efin[b] = em;
                float n = (float)(-ll);
                if (!(n < INFINITY)) n = 0.f;
                nll_out[b] = n;
            }
        }
    }
}

// binary exponent of a positive normal double (others: a very negative number)
__device__ __forceinline__ int exp2_of(double v) {
    const int hi = __double2hiint(v);
    const int ex = (hi >> 20) & 0x7ff;
    return (hi > 0 && ex != 0) ? ex - 1023 : -(1 << 28);
}

// One warp per (b,t) row in which some lane saw a state below 2^kTinyExp of its scale (still exact at that moment, gone a
// few steps later). The paths that will be lost through such a state s weigh w(s) = alpha(s) beta(s) / p(s) - that is the
// total weight of all paths through (t, s) - and the likelihood is at least max_s w(s). If every such state has
// w(s) <= 2^-40 max_s w(s) the loss is invisible in fp32 results; otherwise the sequence goes to the log-space recursion.
// Exponent arithmetic only. Without a gradient (no beta) any such row sends the sequence there.
template <int K>
__global__ void __launch_bounds__(256)
ctc_scan_verify_kernel(const int32_t* __restrict__ tlen, const int32_t* __restrict__ ilen, int Tn, int Bn, int Sp,
                       int have_beta, CtcWs w) {
    constexpr int SA = 32 * K;
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= (long long)Tn * Bn) return;
    const int b = (int)(row / Tn), t = (int)(row - (long long)b * Tn);
    if (t >= ilen[b] || w.flag[b]) return;
    const int ea_l = w.ea[row * 32 + lane];
    const int eb_l = have_beta ? w.eb[row * 32 + lane] : 0;
    if (!__any_sync(0xffffffffu, ((ea_l | eb_l) & 1) != 0)) return;
    if (!have_beta) { if (lane == 0) w.flag[b] = 1; return; }
    const int S = 2 * w.len[b] + 1;
    const double* al = w.alpha + row * SA;
    const double* be = w.beta + row * SA;
    const double* pr = w.pg + row * Sp;
    const int NEG = -(1 << 27);
    int m = NEG, cand = NEG;
    for (int s = lane; s < S; s += 32) {
        const int sm = S - 1 - s;                                      // where the beta scan keeps state s
        const int ma = exp2_of(al[s]), mb = exp2_of(be[sm]);           // relative to the lane scale
        if (ma <= NEG || mb <= NEG) continue;                          // no path through s survives on both sides
        const int x = ma + (w.ea[row * 32 + s / K] >> 1) + mb + (w.eb[row * 32 + sm / K] >> 1) - exp2_of(pr[s]);
        m = max(m, x);
        if (ma < kTinyExp || mb < kTinyExp) cand = max(cand, x);
    }
    m = __reduce_max_sync(0xffffffffu, m);
    cand = __reduce_max_sync(0xffffffffu, cand);
    if (lane == 0 && (m <= NEG || cand + 40 > m)) w.flag[b] = 1;
}

// ---------------------------------------------------------------- pass B (fallback): log-space recursion
__device__ __forceinline__ double lse3(double a, double b, double c) {
    const double mx = fmax(a, fmax(b, c));
    if (mx == -INFINITY) return -INFINITY;
    // the differences are <= 0 and O(1): fp32 transcendentals lose nothing that matters, fp64 keeps the large offset
    const float s = __expf((float)(a - mx)) + __expf((float)(b - mx)) + __expf((float)(c - mx));   // ex2.approx: 2 ulp
    return mx + (double)logf(s);
}

constexpr int kLogPrefetch = 8;

// grid (B, 2): blockIdx.y == 0 runs alpha (and the log-likelihood), blockIdx.y == 1 runs beta. Only flagged sequences.
__global__ void __launch_bounds__(1024)
ctc_alpha_beta_log_kernel(const int32_t* __restrict__ targets, const int32_t* __restrict__ tlen,
                          const int32_t* __restrict__ ilen, int Tn, int Sp, int Sa, float* __restrict__ nll_out, CtcWs w) {
    extern __shared__ double smd[];               // 2 x (Sp + 4) ping-pong state with -inf guards
    const int b = blockIdx.x, s = threadIdx.x;
    if (!w.flag[b]) return;
    const bool is_beta = blockIdx.y == 1;
    const int L = w.len[b], S = 2 * L + 1, Tb = min(ilen[b], Tn);
    const int* tg = w.lab + (long long)b * w.Lp;
    const int W = Sp + 4;
    double* bufA = smd;
    double* bufB = smd + W;
    const bool act = s < S;
    const int cls = act ? ((s & 1) ? tg[s >> 1] : 0) : 0;
    const bool skip_in = act && s > 1 && cls != 0 && cls != ((s & 1) ? tg[(s >> 1) - 1] : 0);      // s-2 -> s allowed
    const bool skip_out = act && (s + 2 < S) && (((s + 2) & 1) ? tg[(s + 2) >> 1] : 0) != 0 &&
                          ((((s + 2) & 1) ? tg[(s + 2) >> 1] : 0) != cls);                              // s -> s+2 allowed
    const float* lp = w.lpg + (long long)b * Tn * Sp;

    for (int i = threadIdx.x; i < 2 * W; i += blockDim.x) smd[i] = -INFINITY;
    __syncthreads();
    if (!is_beta) {
        // ---- alpha: state s lives at index s+2 (two -inf guard cells on the left)
        double* al = w.alpha + (long long)b * Tn * Sa;
        double ll = -INFINITY;
        if (Tb > 0) {
            double a = -INFINITY;
            if (act && s < 2) a = (double)lp[s];
            if (act) { bufA[s + 2] = a; al[s] = a; }
            __syncthreads();
            double* cur = bufA; double* nxt = bufB;
            // the label log-probs are fetched kLogPrefetch steps ahead: a step is shorter than a trip to L2
            float ring[kLogPrefetch];
#pragma unroll
            for (int k = 0; k < kLogPrefetch; ++k) ring[k] = (act && 1 + k < Tb) ? lp[(long long)(1 + k) * Sp + s] : 0.f;
            for (int t0 = 1; t0 < Tb; t0 += kLogPrefetch) {
#pragma unroll
                for (int k = 0; k < kLogPrefetch; ++k) {
                    const int t = t0 + k;
                    if (t >= Tb) break;
                    const double lpt = (double)ring[k];
                    if (act && t + kLogPrefetch < Tb) ring[k] = lp[(long long)(t + kLogPrefetch) * Sp + s];
                    if (act) {
                        const double v = lse3(cur[s + 2], cur[s + 1], skip_in ? cur[s] : -INFINITY);
                        a = (v == -INFINITY) ? -INFINITY : v + lpt;
                        nxt[s + 2] = a;
                        al[(long long)t * Sa + s] = a;
                    }
                    __syncthreads();
                    double* tmp = cur; cur = nxt; nxt = tmp;
                }
            }
            if (threadIdx.x == 0) {
                const double l1 = cur[S - 1 + 2], l2 = (S > 1) ? cur[S - 2 + 2] : -INFINITY;
                const double mx = fmax(l1, l2);
                ll = (mx == -INFINITY) ? -INFINITY : mx + log(exp(l1 - mx) + exp(l2 - mx));
            }
        } else if (L == 0) {
            ll = 0.0;
        }
        if (threadIdx.x == 0) {
            w.ll[b] = ll;
            float n = (float)(-ll);
            if (!(n < INFINITY)) n = 0.f;           // zero_infinity=True (main.py:205)
            nll_out[b] = n;
        }
    } else {
        // ---- beta: state s lives at index s (guard cells on the right)
        if (Tb <= 0) return;
        double* be = w.beta + (long long)b * Tn * Sa;
        double* cur = bufA; double* nxt = bufB;
        {
            double bt = -INFINITY;
            if (act && s >= S - 2) bt = (double)lp[(long long)(Tb - 1) * Sp + s];
            if (act) { cur[s] = bt; be[(long long)(Tb - 1) * Sa + s] = bt; }
        }
        __syncthreads();
        float ring[kLogPrefetch];
#pragma unroll
        for (int k = 0; k < kLogPrefetch; ++k) ring[k] = (act && Tb - 2 - k >= 0) ? lp[(long long)(Tb - 2 - k) * Sp + s] : 0.f;
        for (int t0 = Tb - 2; t0 >= 0; t0 -= kLogPrefetch) {
#pragma unroll
            for (int k = 0; k < kLogPrefetch; ++k) {
                const int t = t0 - k;
                if (t < 0) break;
                const double lpt = (double)ring[k];
                if (act && t - kLogPrefetch >= 0) ring[k] = lp[(long long)(t - kLogPrefetch) * Sp + s];
                if (act) {
                    const double v = lse3(cur[s], cur[s + 1], skip_out ? cur[s + 2] : -INFINITY);
                    const double bt = (v == -INFINITY) ? -INFINITY : v + lpt;
                    nxt[s] = bt;
                    be[(long long)t * Sa + s] = bt;
                }
                __syncthreads();
                double* tmp = cur; cur = nxt; nxt = tmp;
            }
        }
    }
}

// ---------------------------------------------------------------- loss = mean_b(nll_b / max(L_b, 1))
__global__ void ctc_mean_loss_kernel(const float* __restrict__ nll, const int* __restrict__ len, const int* __restrict__ err,
                                     int B, float* __restrict__ loss) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        float acc = 0.f;
        for (int b = 0; b < B; ++b) acc += nll[b] / (float)max(len[b], 1);
        // out-of-range lengths / labels, or a scan that never received its rows: the result is not a loss
        loss[0] = err[0] ? NAN : acc / (float)B;
    }
}

// ---------------------------------------------------------------- pass C: gradient wrt logits
constexpr int kGradThreads = 256;

// value = m * 2^x with m in [1,2); zero / subnormal / non-finite inputs report ok = false
__device__ __forceinline__ bool split_pow2(double v, double& m, int& x) {
    const int hi = __double2hiint(v);
    const int ex = (hi >> 20) & 0x7ff;
    x = ex - 1023;
    m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(v));
    return hi > 0 && ex != 0 && ex != 0x7ff;
}

// state occupancy alpha_t(s) beta_t(s) / (p_t(l'_s) * likelihood) of one row into occ[0..S) (shared memory), by `nthr`
// cooperating threads. Linear mode: mantissa / exponent arithmetic on the block-floating-point scan results.
__device__ __forceinline__ void ctc_occupancy(long long row, int b, int S, int Sp, int Sa, int kscan, double ll, const CtcWs& w,
                                              float* occ, int tid, int nthr) {
    const double* al = w.alpha + row * Sa;
    const double* be = w.beta + row * Sa;
    if (w.flag[b]) {
        const float* lpr = w.lpg + row * Sp;
        for (int s = tid; s < S; s += nthr) {
            const double e = al[s] + be[s] - (double)lpr[s] - ll;        // -inf if either side is unreachable
            occ[s] = (e == e) ? expf((float)e) : 0.f;
        }
    } else {
        const double* pr = w.pg + row * Sp;
        const int* ea = w.ea + row * 32;
        const int* eb = w.eb + row * 32;
        const int efin = w.efin[b];
        double mf; int xf;
        split_pow2(w.pfin[b], mf, xf);
        for (int s = tid; s < S; s += nthr) {
            double ma, mb, mp; int xa, xb, xp;
            const bool ok = split_pow2(al[s], ma, xa) & split_pow2(be[S - 1 - s], mb, xb) & split_pow2(pr[s], mp, xp);
            float o = 0.f;
            if (ok) {
                const int x = xa + xb - xp - xf + (ea[s / kscan] >> 1) + (eb[(S - 1 - s) / kscan] >> 1) - efin;
                // mantissas are in [1,2): single precision is plenty for a value compared at 1e-5
                o = (x < -140) ? 0.f : ldexpf(__fdividef((float)ma * (float)mb, (float)mp * (float)mf), min(x, 8));
            }
            occ[s] = o;
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(kGradThreads)
ctc_grad_kernel(const T* __restrict__ logits, T* __restrict__ grad, int Tn, int Bn, int C, long long stride_t,
                long long stride_b, const int32_t* __restrict__ targets, const int32_t* __restrict__ tlen,
                const int32_t* __restrict__ ilen, int Sp, int Sa, int kscan, float grad_scale, CtcWs w) {
    constexpr int V = Ld<T>::N;
    extern __shared__ float occ[];                                   // [Sp] state occupancy alpha*beta/(p*likelihood)
    const long long row = blockIdx.x;                                // row = b*T + t
    const int b = (int)(row / Tn), t = (int)(row - (long long)b * Tn);
    const T* p = logits + (long long)t * stride_t + (long long)b * stride_b;
    T* g = grad + (long long)t * stride_t + (long long)b * stride_b;
    const int L = w.len[b], S = 2 * L + 1;
    const double ll = w.ll[b];
    const bool dead = (t >= ilen[b]) || (ll == -INFINITY);          // beyond the input length / zero_infinity
    const float scale = dead ? 0.f : grad_scale / ((float)max(L, 1) * (float)Bn);
    const float lse = dead ? 0.f : w.lse[row];

    if (!dead) ctc_occupancy(row, b, S, Sp, Sa, kscan, ll, w, occ, threadIdx.x, blockDim.x);
    // dense part: softmax * scale
    const uintptr_t addr = reinterpret_cast<uintptr_t>(p);
    int head = (int)(((16 - (addr & 15)) & 15) / sizeof(T));
    if (head > C) head = C;
    const bool same_align = ((reinterpret_cast<uintptr_t>(g) & 15) == (addr & 15));
    if (same_align) {
        if ((int)threadIdx.x < head) Ld<T>::st_one(g + threadIdx.x, dead ? 0.f : __expf(Ld<T>::one(p + threadIdx.x) - lse) * scale);
        const int nvec = (C - head) / V;
        for (int vi = threadIdx.x; vi < nvec; vi += blockDim.x) {
            float x[V];
            if (!dead) {
                Ld<T>::vec(p + head + (long long)vi * V, x);
#pragma unroll
                for (int j = 0; j < V; ++j) x[j] = __expf(x[j] - lse) * scale;
            } else {
#pragma unroll
                for (int j = 0; j < V; ++j) x[j] = 0.f;
            }
            Ld<T>::st_vec(g + head + (long long)vi * V, x);
        }
        const int tail0 = head + nvec * V;
        if (tail0 + (int)threadIdx.x < C)
            Ld<T>::st_one(g + tail0 + threadIdx.x, dead ? 0.f : __expf(Ld<T>::one(p + tail0 + threadIdx.x) - lse) * scale);
    } else {
        for (int c = threadIdx.x; c < C; c += blockDim.x)
            Ld<T>::st_one(g + c, dead ? 0.f : __expf(Ld<T>::one(p + c) - lse) * scale);
    }
    if (dead) return;
    __syncthreads();
    // label classes: subtract the occupancy, summed over the states that share a class in a fixed order
    const int* canon = w.canon + (long long)b * Sp;
    const int* tg = w.lab + (long long)b * w.Lp;
    for (int s = threadIdx.x; s < S; s += blockDim.x) {
        if (canon[s] != s) continue;
        float acc = 0.f;
        for (int q = s; q < S; q += ((s & 1) ? 1 : 2))
            if (canon[q] == s) acc += occ[q];
        const int c = (s & 1) ? tg[s >> 1] : 0;
        Ld<T>::st_one(g + c, (__expf(Ld<T>::one(p + c) - lse) - acc) * scale);
    }
}


// ================================================================ overlapped path: rows / fix kernels
constexpr int kRowThreads = 128;

template <typename T> struct RowVec;                                   // a 16-byte vector of logits kept in registers
template <> struct RowVec<float> {
    static constexpr int N = 4;
    static __device__ __forceinline__ void unpack(const uint4& q, float (&o)[4]) {
        o[0] = __uint_as_float(q.x); o[1] = __uint_as_float(q.y); o[2] = __uint_as_float(q.z); o[3] = __uint_as_float(q.w);
    }
};
template <> struct RowVec<__nv_bfloat16> {
    static constexpr int N = 8;
    static __device__ __forceinline__ void unpack(const uint4& q, float (&o)[8]) {
        o[0] = bf16_lo(q.x); o[1] = bf16_hi(q.x); o[2] = bf16_lo(q.y); o[3] = bf16_hi(q.y);
        o[4] = bf16_lo(q.z); o[5] = bf16_hi(q.z); o[6] = bf16_lo(q.w); o[7] = bf16_hi(q.w);
    }
};

// Four warps per logits row, the row in REGISTERS (NV 16-byte vectors per thread, all loads issued at once: a row's whole
// 15-30 KB is in flight, several rows per SM, no staging buffer):
//   max, exp and sum from the registers (two 128-thread barriers) -> log-sum-exp; the label gather for the recursion (plain,
//   mirrored, log; scattered 2/4-byte reads that hit L2) and the row's progress counter (release); then the dense part of
//   the gradient softmax * scale from the same registers. DRAM traffic: logits once, gradient once.
// Rows are dealt in the order (t = 0, all b), (t = T-1, all b), (t = 1, all b), ... so that both recursions are fed from
// their starts; blocks are dispatched in index order.
// Measured and dropped: the row staged in a warp-private shared-memory buffer by bulk copy, one warp per row (6 warps per SM
// cannot hide their own latencies: the whole call 1.53 ms at B = 16 against 0.74 ms for the sequential passes); one warp per
// row with two passes from global memory, the second an L2 hit (0.32 ms for this kernel at B = 16, 3.0 TB/s: DRAM 55 %
// busy, no eligible warp 68 % of the cycles - a warp waits a full DRAM round trip per batch of four loads).
template <typename T, int NV, bool GRAD>
__global__ void __launch_bounds__(kRowThreads)
ctc_rows_kernel(const T* __restrict__ logits, T* __restrict__ grad, int Tn, int Bn, int C, long long stride_t,
                long long stride_b, const int32_t* __restrict__ ilen, int Sp, const float* __restrict__ lse_in,
                float grad_scale, CtcWs w) {
    constexpr int V = RowVec<T>::N;
    __shared__ float w_m[4], w_s[4];
    __shared__ int s_small;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long v = blockIdx.x;
    const int k = (int)(v / (2 * Bn));
    const int r = (int)(v - (long long)k * 2 * Bn);
    int t, b;
    if (r < Bn) { t = k; b = r; }
    else { t = Tn - 1 - k; b = r - Bn; if (t == k) return; }           // odd T: the middle row belongs to the first half
    const long long row = (long long)b * Tn + t;
    const T* p = logits + (long long)t * stride_t + (long long)b * stride_b;
    T* g = GRAD ? grad + (long long)t * stride_t + (long long)b * stride_b : nullptr;
    int* prog = &w.prog[(long long)b * w.NG + t / kGranule];
    const uintptr_t addr = reinterpret_cast<uintptr_t>(p);
    int head = (int)(((16 - (addr & 15)) & 15) / sizeof(T));
    if (head > C) head = C;
    const int nvec = (C - head) / V;
    const int tail0 = head + nvec * V;
    const T* pv = p + head;

    if (t >= ilen[b]) {
        // beyond the input length of this sequence: zero gradient, nothing for the scans
        if (GRAD) {
            const uintptr_t ga = reinterpret_cast<uintptr_t>(g);
            int gh = (int)(((16 - (ga & 15)) & 15) / sizeof(T));
            if (gh > C) gh = C;
            const int gn = (C - gh) / V;
            if (tid < gh) Ld<T>::st_one(g + tid, 0.f);
            for (int vi = tid; vi < gn; vi += kRowThreads) *reinterpret_cast<uint4*>(g + gh + (long long)vi * V) = make_uint4(0u, 0u, 0u, 0u);
            if (gh + gn * V + tid < C) Ld<T>::st_one(g + gh + gn * V + tid, 0.f);
        }
        if (tid == 0) atomicAdd(prog, 1);
        return;
    }
    // ---- the whole row into registers
    uint4 xq[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int vi = tid + i * kRowThreads;
        xq[i] = make_uint4(0u, 0u, 0u, 0u);
        if (vi < nvec) xq[i] = ld_nc_v4(pv + (long long)vi * V);
    }
    const int sc = tid < head ? tid : tail0 + (tid - head);            // scalar element in front of / behind the interior
    const bool has_sc = tid < 2 * V && sc < C;
    float x_sc = 0.f;
    if (has_sc) x_sc = Ld<T>::one(p + sc);
    if (tid == 0) s_small = 0;

    const int L = w.len[b], S = 2 * L + 1;
    // the raw logit of this thread's first label state, fetched while the row is in flight (its two dependent round trips -
    // label, then logit - would otherwise sit behind the log-sum-exp on the CTA's critical path)
    float xg0 = 0.f;
    if (tid < S) {
        const int c = (tid & 1) ? w.lab[(long long)b * w.Lp + (tid >> 1)] : 0;
        xg0 = Ld<T>::one(p + c);
    }
    const float scale = grad_scale / ((float)max(L, 1) * (float)Bn);
    float lse, mul, inv_sum = 0.f, m_row = 0.f;
    bool have_e = false;               // x_sc (and, for fp32 rows, the registers) hold exp(x - m_row) instead of x
    if (lse_in != nullptr) {
        lse = lse_in[row];            // log-sum-exp already produced by the classifier epilogue
        mul = scale;
        __syncthreads();              // s_small initialised
    } else {
        // ---- maximum
        float m = has_sc ? x_sc : -INFINITY;
        if (V == 8) {
            // bf16 rows stay packed: 4 HMNMX2 per 8 elements, two independent chains
            __nv_bfloat162 a0 = __float2bfloat162_rn(-INFINITY), a1 = a0;
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                if (tid + i * kRowThreads < nvec) {
                    a0 = __hmax2(a0, __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&xq[i].x), *reinterpret_cast<const __nv_bfloat162*>(&xq[i].y)));
                    a1 = __hmax2(a1, __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&xq[i].z), *reinterpret_cast<const __nv_bfloat162*>(&xq[i].w)));
                }
            }
            a0 = __hmax2(a0, a1);
            m = fmaxf(m, fmaxf(__low2float(a0), __high2float(a0)));
        } else {
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                if (tid + i * kRowThreads < nvec) {
                    float x[V];
                    RowVec<T>::unpack(xq[i], x);
                    float a = fmaxf(x[0], x[1]);
#pragma unroll
                    for (int j = 2; j < V; j += 2) a = fmaxf(a, fmaxf(x[j], x[j + 1]));
                    m = fmaxf(m, a);
                }
            }
        }
        m = warp_max(m);
        if (lane == 0) w_m[warp] = m;
        __syncthreads();
        m = fmaxf(fmaxf(w_m[0], w_m[1]), fmaxf(w_m[2], w_m[3]));
        // ---- sum of exp(x - m)
        float s0 = 0.f, s1 = 0.f;
        if (has_sc) { x_sc = ex2_fast((x_sc - m) * kLog2e); s0 = x_sc; }      // from here on x_sc = exp(x - m)
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            if (tid + i * kRowThreads < nvec) {
                float x[V];
                RowVec<T>::unpack(xq[i], x);
                float acc = 0.f;
#pragma unroll
                for (int j = 0; j < V; ++j) { x[j] = ex2_fast((x[j] - m) * kLog2e); acc += x[j]; }
                if (i & 1) s1 += acc; else s0 += acc;
                if (V == 4)          // fp32 rows: exp(x - m) replaces x in the registers, the gradient pass only scales it
                    xq[i] = make_uint4(__float_as_uint(x[0]), __float_as_uint(x[1]), __float_as_uint(x[2]), __float_as_uint(x[3 % V]));
            }
        }
        const float ws = warp_sum(s0 + s1);
        if (lane == 0) w_s[warp] = ws;
        __syncthreads();
        const float sum = (w_s[0] + w_s[1]) + (w_s[2] + w_s[3]);
        lse = m + logf(sum);
        mul = scale;
        have_e = true;
        inv_sum = 1.f / sum;
        m_row = m;
    }
    if (tid == 0) w.lse[row] = lse;
    // ---- label probabilities for the recursion (plain and mirrored), log-probs for the log-space fallback
    {
        const int* tg = w.lab + (long long)b * w.Lp;
        float* dst = w.lpg + row * Sp;
        double* dpr = w.pg + row * Sp;
        double* dpm = w.pgr + row * Sp;
        bool small = false;
        for (int q = tid; q < Sp; q += kRowThreads) {
            float lp = 0.f;
            double pr = 0.0;
            if (q < S) {
                float xc = xg0;
                if (q != tid) {
                    const int c = (q & 1) ? tg[q >> 1] : 0;
                    xc = Ld<T>::one(p + c);
                }
                lp = xc - lse;
                pr = (double)expf(lp);
                small |= (lp < kMinLinearLogProb) && (lp > -INFINITY);
                small |= !(lp == lp);
                dpm[S - 1 - q] = pr;
            } else {
                dpm[q] = 0.0;
            }
            dst[q] = lp;
            dpr[q] = pr;
        }
        if (small) s_small = 1;
    }
    // ---- publish (release) for the scans before the gradient pass, whose stores the fence would otherwise wait for
    __syncthreads();
    if (tid == 0) {
        if (s_small) w.flag[b] = 1;
        __threadfence();
        atomicAdd(prog, 1);
    }
    if (!GRAD) return;
    // ---- dense part of the gradient from the registers, softmax * scale (label classes: ctc_fix_kernel)
    const float nl = -lse * kLog2e;
    const bool same_align = ((reinterpret_cast<uintptr_t>(g) & 15) == (addr & 15));
    const bool scaled = have_e && V == 4;          // registers hold exp(x - m): softmax = e / sum
    const float emul = mul * inv_sum;
    (void)m_row;
    if (has_sc) Ld<T>::st_one(g + sc, have_e ? x_sc * emul : ex2_fast(fmaf(x_sc, kLog2e, nl)) * mul);
    if (same_align) {
        T* gv = g + head;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int vi = tid + i * kRowThreads;
            if (vi < nvec) {
                float x[V];
                RowVec<T>::unpack(xq[i], x);
#pragma unroll
                for (int j = 0; j < V; ++j) x[j] = scaled ? x[j] * emul : ex2_fast(fmaf(x[j], kLog2e, nl)) * mul;
                Ld<T>::st_vec(gv + (long long)vi * V, x);
            }
        }
    } else {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int vi = tid + i * kRowThreads;
            if (vi < nvec) {
                float x[V];
                RowVec<T>::unpack(xq[i], x);
#pragma unroll
                for (int j = 0; j < V; ++j)
                    Ld<T>::st_one(g + head + (long long)vi * V + j, scaled ? x[j] * emul : ex2_fast(fmaf(x[j], kLog2e, nl)) * mul);
            }
        }
    }
}

// After the scans: one warp per row subtracts the state occupancy at the label classes,
//   grad[c] = (softmax_c - sum_{s: l'_s = c} occ(s)) * scale
// duplicates of a class summed in a fixed order (deterministic), and zeroes the rows of sequences with an infinite loss
// (zero_infinity). Same arithmetic as the tail of ctc_grad_kernel.
constexpr int kFixWarps = 8;
template <typename T>
__global__ void __launch_bounds__(kFixWarps * 32)
ctc_fix_kernel(T* __restrict__ grad, int Tn, int Bn, int C, long long stride_t, long long stride_b,
               const int32_t* __restrict__ ilen, int Sp, int Sa, int kscan, float grad_scale, int relative, CtcWs w) {
    extern __shared__ float fix_occ[];                                  // [kFixWarps][Sp]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long row = (long long)blockIdx.x * kFixWarps + warp;     // row = b*T + t
    if (row >= (long long)Tn * Bn) return;
    const int b = (int)(row / Tn), t = (int)(row - (long long)b * Tn);
    if (t >= ilen[b]) return;                                            // zeroed by the rows kernel
    T* g = grad + (long long)t * stride_t + (long long)b * stride_b;
    const double ll = w.ll[b];
    if (ll == -INFINITY) {
        // zero_infinity: the whole row (the rows kernel wrote softmax * scale before the likelihood was known)
        const RowGeom gg = row_geom(g, C * (int)sizeof(T));
        unsigned char* gb = reinterpret_cast<unsigned char*>(g);
        const int nvec = gg.body_bytes >> 4;
        for (int vi = lane; vi < nvec; vi += 32) *reinterpret_cast<uint4*>(gb + gg.head_bytes + (vi << 4)) = make_uint4(0u, 0u, 0u, 0u);
        const int nhead = gg.head_bytes / (int)sizeof(T);
        const int tail0 = nhead + nvec * (16 / (int)sizeof(T));
        const int sc = lane < nhead ? lane : tail0 + (lane - nhead);
        if (lane < 16 && sc < C) Ld<T>::st_one(g + sc, 0.f);
        return;
    }
    const int L = w.len[b], S = 2 * L + 1;
    const float scale = grad_scale / ((float)max(L, 1) * (float)Bn);
    float* occ = fix_occ + (size_t)warp * Sp;
    ctc_occupancy(row, b, S, Sp, Sa, kscan, ll, w, occ, lane, 32);
    __syncwarp();
    const int* canon = w.canon + (long long)b * Sp;
    const int* nxt = w.nxt + (long long)b * Sp;
    const int* tg = w.lab + (long long)b * w.Lp;
    const float* lpr = w.lpg + row * Sp;                                 // log-softmax at the label classes ...
    const float corr = relative ? w.ref[row] - w.lse[row] : 0.f;         // ... relative to the row's reference in the relative schedule
    // blank (every even state): a fixed tree over the warp
    {
        float acc = 0.f;
        for (int s = 2 * lane; s < S; s += 64) acc += occ[s];
        acc = warp_sum(acc);
        if (lane == 0) Ld<T>::st_one(g, (__expf(lpr[0] + corr) - acc) * scale);
    }
    // labels (odd states): the first state of a class walks the chain of its repeats, in state order
    for (int s = 2 * lane + 1; s < S; s += 64) {
        if (canon[s] != s) continue;
        float acc = occ[s];
        for (int q = nxt[s]; q >= 0; q = nxt[q]) acc += occ[q];
        Ld<T>::st_one(g + tg[s >> 1], (__expf(lpr[s] + corr) - acc) * scale);
    }
}


// ---------------------------------------------------------------- host side of the overlapped path
// helper stream + fork/join events, one set per device, created on first use (the only persistent state of this file)
struct CtcSide {
    std::mutex mu;
    cudaStream_t helper = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
};
static int ctc_side(CtcSide** out) {
    static CtcSide sides[PerDeviceOnce::kMaxDev];
    static std::mutex create_mu;
    int dev = 0;
    HCTR_CUDA(cudaGetDevice(&dev));
    HCTR_CHECK(dev >= 0 && dev < PerDeviceOnce::kMaxDev, HCTR_ERR_INVALID, "ctc_loss: device index %d out of range", dev);
    CtcSide& sd = sides[dev];
    {
        std::lock_guard<std::mutex> guard(create_mu);
        if (sd.helper == nullptr) {
            // highest priority: the 2*B latency-bound scan warps must get their SM slots ahead of the rows kernel's blocks
            int prio_lo = 0, prio_hi = 0;
            HCTR_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
            HCTR_CUDA(cudaStreamCreateWithPriority(&sd.helper, cudaStreamNonBlocking, prio_hi));
            HCTR_CUDA(cudaEventCreateWithFlags(&sd.fork, cudaEventDisableTiming));
            HCTR_CUDA(cudaEventCreateWithFlags(&sd.join, cudaEventDisableTiming));
        }
    }
    *out = &sd;
    return HCTR_OK;
}

// One launch site for the scans: K states per lane, NST staged chunks (NST * 8 KB of dynamic shared memory, or more when the
// caller asks for a whole SM's worth to keep the one-warp CTAs one to an SM).
constexpr int kScanStages = 3;
constexpr int kScanDeepStages = 12;
constexpr size_t kScanExclusiveSmem = 226 * 1024;      // + static + 1 KB reserved = the SM's 228 KB
constexpr int kSplitMaxScanCtas = 48;                     // SMs the split schedule may take away from the dense pass
template <int K, bool POLL, int NST>
static int scan_launch_k(dim3 grid, size_t min_smem, cudaStream_t st, const int32_t* ilen, int T, int Sp, float* nll, const CtcWs& w) {
    auto kern = ctc_scan_kernel<K, POLL, NST>;
    static PerDeviceOnce once;
    int dev;
    if (once.need(dev)) {
        HCTR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kScanExclusiveSmem));
        once.mark(dev);
    }
    size_t smem = (size_t)NST * kScanChunkBytes;
    if (smem < min_smem) smem = min_smem;
    kern<<<grid, 32, smem, st>>>(ilen, T, Sp, nll, w);
    HCTR_CUDA(cudaGetLastError());
    return HCTR_OK;
}
template <bool POLL, int NST>
static int scan_launch(int kscan, dim3 grid, size_t min_smem, cudaStream_t st, const int32_t* ilen, int T, int Sp, float* nll,
                       const CtcWs& w) {
    switch (kscan) {
        case 4:  return scan_launch_k<4, POLL, NST>(grid, min_smem, st, ilen, T, Sp, nll, w);
        case 8:  return scan_launch_k<8, POLL, NST>(grid, min_smem, st, ilen, T, Sp, nll, w);
        case 16: return scan_launch_k<16, POLL, NST>(grid, min_smem, st, ilen, T, Sp, nll, w);
        default: return HCTR_OK;                      // too many states: the log-space recursion only
    }
}

// The split schedule is the default when a gradient is wanted, the scans' 2*B one-warp CTAs can each have an SM to themselves
// without starving the dense pass (B <= 24), and the logits are large enough for the dense pass to be worth hiding.
static bool split_by_default(int T, int B, int C, int esz, bool want_grad) {
    return want_grad && 2 * B <= kSplitMaxScanCtas && (long long)T * B * C * esz >= (32ll << 20);
}

// vectors per thread the row needs; 0 = the row does not fit the registers of four warps (sequential passes instead)
static int rows_nv(int C, int esz) {
    const int V = 16 / esz;
    const int need = (C / V + kRowThreads - 1) / kRowThreads;
    return need <= 4 ? 4 : need <= 8 ? 8 : need <= 16 ? 16 : 0;
}

// Loads the kernel before the scan is in flight: CUDA loads functions lazily, and loading may wait for the device to drain -
// which never happens while a scan kernel spins on rows this very kernel is supposed to produce.
template <typename T, int NV, bool GRAD>
static int rows_call(bool launch, const void* logits, void* grad, int T_, int B, int C, long long stride_t, long long stride_b,
                     const int32_t* ilen, int Sp, const float* row_lse, float grad_scale, const CtcWs* w, cudaStream_t s) {
    auto kern = ctc_rows_kernel<T, NV, GRAD>;
    if (!launch) {
        static PerDeviceOnce once;
        int dev;
        if (once.need(dev)) {
            cudaFuncAttributes fa;
            HCTR_CUDA(cudaFuncGetAttributes(&fa, kern));
            once.mark(dev);
        }
        return HCTR_OK;
    }
    const long long total = (long long)((T_ + 1) / 2) * 2 * B;
    HCTR_CHECK(total < (1ll << 31), HCTR_ERR_INVALID, "ctc_loss: too many rows");
    kern<<<(unsigned)total, kRowThreads, 0, s>>>(static_cast<const T*>(logits), static_cast<T*>(grad), T_, B, C, stride_t, stride_b,
                                                 ilen, Sp, row_lse, grad_scale, *w);
    HCTR_CUDA(cudaGetLastError());
    return HCTR_OK;
}
template <typename T, bool GRAD>
static int rows_call_nv(int nv, bool launch, const void* logits, void* grad, int T_, int B, int C, long long stride_t,
                        long long stride_b, const int32_t* ilen, int Sp, const float* row_lse, float grad_scale, const CtcWs* w,
                        cudaStream_t s) {
    switch (nv) {
        case 4:  return rows_call<T, 4, GRAD>(launch, logits, grad, T_, B, C, stride_t, stride_b, ilen, Sp, row_lse, grad_scale, w, s);
        case 8:  return rows_call<T, 8, GRAD>(launch, logits, grad, T_, B, C, stride_t, stride_b, ilen, Sp, row_lse, grad_scale, w, s);
        default: return rows_call<T, 16, GRAD>(launch, logits, grad, T_, B, C, stride_t, stride_b, ilen, Sp, row_lse, grad_scale, w, s);
    }
}
static int rows_dispatch(bool launch, int nv, const void* logits, void* grad, int dtype, bool want_grad, int T_, int B, int C,
                         long long stride_t, long long stride_b, const int32_t* ilen, int Sp, const float* row_lse,
                         float grad_scale, const CtcWs* w, cudaStream_t s) {
    if (dtype == HCTR_F32)
        return want_grad ? rows_call_nv<float, true>(nv, launch, logits, grad, T_, B, C, stride_t, stride_b, ilen, Sp, row_lse, grad_scale, w, s)
                         : rows_call_nv<float, false>(nv, launch, logits, grad, T_, B, C, stride_t, stride_b, ilen, Sp, row_lse, grad_scale, w, s);
    return want_grad ? rows_call_nv<__nv_bfloat16, true>(nv, launch, logits, grad, T_, B, C, stride_t, stride_b, ilen, Sp, row_lse, grad_scale, w, s)
                     : rows_call_nv<__nv_bfloat16, false>(nv, launch, logits, grad, T_, B, C, stride_t, stride_b, ilen, Sp, row_lse, grad_scale, w, s);
}

}  // namespace hctr

using namespace hctr;

extern "C" {

long long hctr_ctc_loss_flag_offset(int T, int B, int max_target_len) {
    if (T <= 0 || B <= 0 || max_target_len < 0) return -1;
    const int Sp = state_pitch(max_target_len);
    CtcWs w = carve(nullptr, T, B, Sp, alpha_pitch(Sp), nullptr);
    return static_cast<long long>(reinterpret_cast<intptr_t>(w.flag));
}

long long hctr_ctc_loss_workspace_bytes(int T, int B, int max_target_len) {
    if (T <= 0 || B <= 0 || max_target_len < 0) return 0;
    long long total = 0;
    const int Sp = state_pitch(max_target_len);
    carve(nullptr, T, B, Sp, alpha_pitch(Sp), &total);
    return total;
}

int hctr_ctc_loss_fwd_bwd(const void* logits, int dtype, int T, int B, int C, long long stride_t, long long stride_b,
                          const int32_t* targets, const int32_t* target_lengths, const int32_t* input_lengths,
                          int max_target_len, const float* row_lse, float* nll, float* loss, void* grad, float grad_scale,
                          void* workspace, long long workspace_bytes, void* stream) {
    HCTR_CHECK(logits && target_lengths && input_lengths && nll && loss, HCTR_ERR_INVALID, "ctc_loss: null pointer");
    HCTR_CHECK(T > 0 && B > 0 && C > 1, HCTR_ERR_INVALID, "ctc_loss: bad shape T=%d B=%d C=%d", T, B, C);
    HCTR_CHECK(dtype == HCTR_F32 || dtype == HCTR_BF16, HCTR_ERR_INVALID, "ctc_loss: bad dtype");
    HCTR_CHECK(max_target_len >= 0 && 2 * max_target_len + 1 <= 1024, HCTR_ERR_INVALID,
               "ctc_loss: target length %d exceeds the 511-label limit of the one-CTA-per-sequence recursion", max_target_len);
    HCTR_CHECK(targets != nullptr || max_target_len == 0, HCTR_ERR_INVALID, "ctc_loss: null targets");
    const int Sp = state_pitch(max_target_len);
    const int kscan = scan_lane_states(Sp);           // states per lane of the one-warp scan; 0 = log-space only
    const int Sa = alpha_pitch(Sp);
    long long need = 0;
    CtcWs w = carve(workspace, T, B, Sp, Sa, &need);
    HCTR_CHECK(workspace && workspace_bytes >= need, HCTR_ERR_INVALID, "ctc_loss: workspace too small (%lld < %lld)", workspace_bytes, need);
    HCTR_CHECK((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, HCTR_ERR_INVALID, "ctc_loss: workspace must be 256-byte aligned");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const long long rows = (long long)T * B;
    HCTR_CHECK(rows < (1ll << 31), HCTR_ERR_INVALID, "ctc_loss: too many rows");

    static const bool debug_force_log = getenv("HCTR_CTC_DEBUG_FORCE_LOG") != nullptr;           // diagnostics only
    static const bool debug_no_fallback = getenv("HCTR_CTC_DEBUG_NO_FALLBACK") != nullptr;     // diagnostics only
    const char* ov = getenv("HCTR_CTC_OVERLAP");                      // read per call: tests and profiling runs flip it
    const bool overlap_off = ov && ov[0] == '0';
    HCTR_CUDA(cudaMemsetAsync(w.err, 0, 16, s));
    ctc_prep_kernel<<<B, 128, 0, s>>>(targets, target_lengths, B, C, Sp, max_target_len, (kscan == 0 || debug_force_log) ? 1 : 0, w);
    HCTR_CUDA(cudaGetLastError());

    const int have_beta = grad != nullptr;
    const dim3 gridB(B, have_beta ? 2 : 1);
    const int blocksV = (int)((rows + 7) / 8);
    const int esz = dtype == HCTR_F32 ? 4 : 2;

    // ---- the one-pass rows kernel needs the row in the registers of four warps; the overlapped schedule needs the scans'
    //      2*B one-warp CTAs all resident beside it ("2": the same kernels back to back on one stream, e.g. under ncu)
    const int nv = rows_nv(C, esz);
    const bool rows_path = !overlap_off && kscan != 0 && !debug_force_log && nv != 0;
    // Schedule. Default: rows -> scans -> fix back to back on the caller's stream. HCTR_CTC_OVERLAP=1 opts into the overlapped
    // schedule: the scans run on a helper stream underneath the rows kernel, fed through the progress counters; when the rows
    // kernel ends each scan still has the second half of its steps to go (the rows were dealt from both ends), so the gain is
    // at most half a scan (0.12 ms) minus the cost of sharing the SMs. Measured (T=2048, C=7375, bf16; serial / overlapped):
    // B=2 0.30 / 0.34 ms, B=16 0.56 / 0.49, B=64 1.42 / 1.47. It stays opt-in because it needs the two kernels to really run
    // at the same time: under a tool that serialises kernels (ncu, compute-sanitizer) the scans wait for rows that cannot
    // start, give up after kScanWaitLimitNs and the loss comes out NaN.
    const bool want_overlap = ov && ov[0] == '1';
    // The split schedule (see ctc_lse_chunk_kernel): A' -> [scans || dense gradient on the helper stream] -> fix. "4" forces it,
    // "2" forces the one-pass rows kernel back to back with the scans.
    const bool split = kscan != 0 && !debug_force_log && !overlap_off && !want_overlap && !(ov && (ov[0] == '2' || ov[0] == '3')) &&
                       ((ov && (ov[0] == '4' || ov[0] == '5')) || split_by_default(T, B, C, esz, grad != nullptr));
    // The relative form of the split schedule (ctc_gather_rel_kernel): the scans start right after a sparse label gather and the
    // full-row log-sum-exp pass joins the dense gradient pass on the helper stream underneath them. Used when the caller has
    // no row log-sum-exp to offer (with one, the gather-only form of pass A' is just as short); "5" forces it, "4" the other.
    const bool relative = split && grad != nullptr && ((ov && ov[0] == '5') || (row_lse == nullptr && !(ov && ov[0] == '4')));
    const bool serial = !split && rows_path && (!want_overlap || 2ll * B > 2048);
    const bool overlap = rows_path || split;                          // the gradient's label classes are ctc_fix_kernel's
    CtcSide* split_side = nullptr;
    if (split) {
        const long long blocksA = (rows + kLseWarps - 1) / kLseWarps;
        if (relative) {
            if (dtype == HCTR_F32)
                ctc_gather_rel_kernel<float><<<(int)blocksA, kLseWarps * 32, 0, s>>>(
                    static_cast<const float*>(logits), T, B, stride_t, stride_b, input_lengths, Sp, w);
            else
                ctc_gather_rel_kernel<__nv_bfloat16><<<(int)blocksA, kLseWarps * 32, 0, s>>>(
                    static_cast<const __nv_bfloat16*>(logits), T, B, stride_t, stride_b, input_lengths, Sp, w);
        } else if (dtype == HCTR_F32)
            ctc_lse_chunk_kernel<float><<<(int)blocksA, kLseWarps * 32, 0, s>>>(
                static_cast<const float*>(logits), T, B, C, stride_t, stride_b, input_lengths, Sp, row_lse, w);
        else
            ctc_lse_chunk_kernel<__nv_bfloat16><<<(int)blocksA, kLseWarps * 32, 0, s>>>(
                static_cast<const __nv_bfloat16*>(logits), T, B, C, stride_t, stride_b, input_lengths, Sp, row_lse, w);
        HCTR_CUDA(cudaGetLastError());
        if (grad != nullptr) {
            int rc = ctc_side(&split_side);
            if (rc) return rc;
            split_side->mu.lock();                                    // one enqueue at a time per device (shared helper stream)
            cudaError_t e = cudaEventRecord(split_side->fork, s);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(split_side->helper, split_side->fork, 0);
            if (e != cudaSuccess) { split_side->mu.unlock(); HCTR_CUDA(e); }
        }
        // the scans first: their 2*B one-warp CTAs take their SM slots on an idle device, the dense pass fills in around them
        const bool timing = grad != nullptr && getenv("HCTR_CTC_TIMING") != nullptr;     // diagnostics: scan / dense pass times
        cudaEvent_t tev[4] = {nullptr, nullptr, nullptr, nullptr};
        if (timing) { for (int i = 0; i < 4; ++i) cudaEventCreate(&tev[i]); cudaEventRecord(tev[0], s); }
        {
            // Beside the dense pass each scan CTA asks for (nearly) a whole SM's shared memory, and the dense kernel for a token
            // 4 KB, so that no dense CTA can share an SM with a scan: co-resident, the dense warps keep the SM's load/store
            // queue full and every shared-memory read, cp.async and store of the scan waits in it - measured, the scans then
            // advance at 15 % of their speed until the dense pass is over (0.38 ms instead of 0.24; 0.28 with the SM to
            // themselves and kScanDeepStages chunks staged ahead, which covers the longer DRAM round trips of a busy device).
            const int rc = grad != nullptr ? scan_launch<false, kScanDeepStages>(kscan, gridB, kScanExclusiveSmem, s, input_lengths, T, Sp, nll, w)
                                           : scan_launch<false, kScanStages>(kscan, gridB, 0, s, input_lengths, T, Sp, nll, w);
            if (rc) { if (split_side) split_side->mu.unlock(); return rc; }
        }
        if (timing) { cudaEventRecord(tev[1], s); cudaEventRecord(tev[2], split_side->helper); }
        if (grad != nullptr) {
            if (relative) {
                // the rows' log-sum-exp (given, or one read of every row), on the helper stream beside the scans; like the dense
                // kernel it asks for a token 4 KB of shared memory so that none of its CTAs can share an SM with a scan
                if (dtype == HCTR_F32)
                    ctc_lse_chunk_kernel<float, false><<<(int)blocksA, kLseWarps * 32, 4096, split_side->helper>>>(
                        static_cast<const float*>(logits), T, B, C, stride_t, stride_b, input_lengths, Sp, row_lse, w);
                else
                    ctc_lse_chunk_kernel<__nv_bfloat16, false><<<(int)blocksA, kLseWarps * 32, 4096, split_side->helper>>>(
                        static_cast<const __nv_bfloat16*>(logits), T, B, C, stride_t, stride_b, input_lengths, Sp, row_lse, w);
            }
            const long long blocksD = (rows + kDenseWarps - 1) / kDenseWarps;
            if (dtype == HCTR_F32)
                ctc_dense_grad_kernel<float><<<(int)blocksD, kDenseWarps * 32, 4096, split_side->helper>>>(
                    static_cast<const float*>(logits), static_cast<float*>(grad), T, B, C, stride_t, stride_b, input_lengths, grad_scale, w);
            else
                ctc_dense_grad_kernel<__nv_bfloat16><<<(int)blocksD, kDenseWarps * 32, 4096, split_side->helper>>>(
                    static_cast<const __nv_bfloat16*>(logits), static_cast<__nv_bfloat16*>(grad), T, B, C, stride_t, stride_b,
                    input_lengths, grad_scale, w);
            cudaError_t e = cudaGetLastError();
            if (timing) {
                cudaEventRecord(tev[3], split_side->helper);
                cudaEventSynchronize(tev[1]); cudaEventSynchronize(tev[3]);
                float scan_ms = 0.f, dense_ms = 0.f;
                cudaEventElapsedTime(&scan_ms, tev[0], tev[1]); cudaEventElapsedTime(&dense_ms, tev[2], tev[3]);
                fprintf(stderr, "hctr ctc timing (split): scans %.3f ms, dense gradient pass %.3f ms\n", scan_ms, dense_ms);
                for (int i = 0; i < 4; ++i) cudaEventDestroy(tev[i]);
            }
            if (e == cudaSuccess) e = cudaEventRecord(split_side->join, split_side->helper);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(s, split_side->join, 0);    // join: everything after this on s follows the dense pass
            split_side->mu.unlock();
            HCTR_CUDA(e);
        }
        HCTR_CUDA(cudaGetLastError());
    } else if (serial) {
        int rc = rows_dispatch(true, nv, logits, grad, dtype, grad != nullptr, T, B, C, stride_t, stride_b, input_lengths, Sp, row_lse,
                               grad_scale, &w, s);
        if (rc) return rc;
        rc = (ov && ov[0] == '3') ? scan_launch<true, kScanStages>(kscan, gridB, 0, s, input_lengths, T, Sp, nll, w)
                                  : scan_launch<false, kScanStages>(kscan, gridB, 0, s, input_lengths, T, Sp, nll, w);
        if (rc) return rc;
    } else if (overlap) {
        CtcSide* side = nullptr;
        int rc = ctc_side(&side);
        if (rc) return rc;
        rc = rows_dispatch(false, nv, logits, grad, dtype, grad != nullptr, T, B, C, stride_t, stride_b, input_lengths, Sp, row_lse,
                           grad_scale, &w, s);                       // load the function before the scan is in flight
        if (rc) return rc;
        std::lock_guard<std::mutex> guard(side->mu);                 // one enqueue at a time per device (shared helper stream)
        // fork: the scans only need the prep kernel
        HCTR_CUDA(cudaEventRecord(side->fork, s));
        HCTR_CUDA(cudaStreamWaitEvent(side->helper, side->fork, 0));
        const char* od = getenv("HCTR_CTC_ROWS_FIRST");
        const bool rows_first = od && od[0] == '1';
        const bool timing = getenv("HCTR_CTC_TIMING") != nullptr;    // diagnostics: per-kernel times of the overlapped schedule
        cudaEvent_t tev[4] = {nullptr, nullptr, nullptr, nullptr};
        if (timing) { for (int i = 0; i < 4; ++i) cudaEventCreate(&tev[i]); cudaEventRecord(tev[2], side->helper); }
        if (timing && rows_first) cudaEventRecord(tev[0], s);
        if (rows_first) {
            rc = rows_dispatch(true, nv, logits, grad, dtype, grad != nullptr, T, B, C, stride_t, stride_b, input_lengths, Sp, row_lse,
                               grad_scale, &w, s);
            if (rc) return rc;
        }
        // One scan CTA per SM, enforced by an unused dynamic shared-memory request: launched beside a kernel that fills the
        // GPU, the one-warp CTAs are otherwise packed eight to an SM - on the few SMs that free up first - and, one warp per
        // CTA, onto the same scheduler of it: the recursion then runs 4x slower (measured: 32 scan CTAs on SMs 0, 6, 32, 64).
        const size_t spread = 112 * 1024;
        rc = scan_launch<true, kScanStages>(kscan, gridB, spread, side->helper, input_lengths, T, Sp, nll, w);
        if (rc) return rc;
        if (timing) cudaEventRecord(tev[3], side->helper);
        HCTR_CUDA(cudaEventRecord(side->join, side->helper));
        if (!rows_first) {
            if (timing) cudaEventRecord(tev[0], s);
            rc = rows_dispatch(true, nv, logits, grad, dtype, grad != nullptr, T, B, C, stride_t, stride_b, input_lengths, Sp, row_lse,
                               grad_scale, &w, s);
            if (rc) return rc;
        }
        if (timing) {
            cudaEventRecord(tev[1], s);
            cudaEventSynchronize(tev[1]); cudaEventSynchronize(tev[3]);
            float rows_ms = 0.f, scan_ms = 0.f, lead_ms = 0.f, tail_ms = 0.f;
            cudaEventElapsedTime(&rows_ms, tev[0], tev[1]); cudaEventElapsedTime(&scan_ms, tev[2], tev[3]);
            cudaEventElapsedTime(&lead_ms, tev[2], tev[0]); cudaEventElapsedTime(&tail_ms, tev[1], tev[3]);
            fprintf(stderr, "hctr ctc timing: rows %.3f ms, scan %.3f ms, scan start -> rows start %.3f ms, rows end -> scan end %.3f ms\n",
                    rows_ms, scan_ms, lead_ms, tail_ms);
            for (int i = 0; i < 4; ++i) cudaEventDestroy(tev[i]);
            {
                unsigned long long all[16 * 2 * 8];
                const int nb = B < 16 ? B : 16;
                cudaMemcpy(all, w.dbg, sizeof(unsigned long long) * nb * 16, cudaMemcpyDeviceToHost);
                fprintf(stderr, "   scan CTA -> SM:");
                for (int q = 0; q < nb * 2; ++q) fprintf(stderr, " %llu", all[q * 8 + 7]);
                fprintf(stderr, "\n");
            }
            unsigned long long hd[32];
            cudaMemcpy(hd, w.dbg, sizeof(hd), cudaMemcpyDeviceToHost);           // sequences 0 and 1, both directions
            for (int q = 0; q < 4; ++q)
                fprintf(stderr, "   scan b=%d %s: lifetime %.3f ms (quarters at %.3f %.3f %.3f), waiting %.3f ms in %llu polls\n", q / 2,
                        (q & 1) ? "beta " : "alpha", (hd[q * 8 + 1] - hd[q * 8]) * 1e-6, (hd[q * 8 + 4] - hd[q * 8]) * 1e-6,
                        (hd[q * 8 + 5] - hd[q * 8]) * 1e-6, (hd[q * 8 + 6] - hd[q * 8]) * 1e-6, hd[q * 8 + 2] * 1e-6, hd[q * 8 + 3]);
        }
        HCTR_CUDA(cudaStreamWaitEvent(s, side->join, 0));           // join
    } else {
        const long long blocksA = (rows + kLseWarps - 1) / kLseWarps;
        if (dtype == HCTR_F32)
            ctc_lse_gather_kernel<float><<<(int)blocksA, kLseWarps * 32, 0, s>>>(
                static_cast<const float*>(logits), T, B, C, stride_t, stride_b, targets, target_lengths, input_lengths, Sp, row_lse, w);
        else
            ctc_lse_gather_kernel<__nv_bfloat16><<<(int)blocksA, kLseWarps * 32, 0, s>>>(
                static_cast<const __nv_bfloat16*>(logits), T, B, C, stride_t, stride_b, targets, target_lengths, input_lengths, Sp, row_lse, w);
        HCTR_CUDA(cudaGetLastError());
        // one warp per (sequence, direction); the beta recursion is only needed for the gradient
        {
            const int rc = scan_launch<false, kScanStages>(kscan, gridB, 0, s, input_lengths, T, Sp, nll, w);
            if (rc) return rc;
        }
    }
    switch (kscan) {
        case 4:  ctc_scan_verify_kernel<4><<<blocksV, 256, 0, s>>>(target_lengths, input_lengths, T, B, Sp, have_beta, w); break;
        case 8:  ctc_scan_verify_kernel<8><<<blocksV, 256, 0, s>>>(target_lengths, input_lengths, T, B, Sp, have_beta, w); break;
        case 16: ctc_scan_verify_kernel<16><<<blocksV, 256, 0, s>>>(target_lengths, input_lengths, T, B, Sp, have_beta, w); break;
        default: break;
    }
    HCTR_CUDA(cudaGetLastError());
    // log-space recursion for the flagged sequences (exits at once for the others)
    const int threads = (2 * max_target_len + 1 + 31) / 32 * 32;
    const size_t smB = (size_t)(2 * (Sp + 4) + 2) * sizeof(double);
    if (!debug_no_fallback) {
        ctc_alpha_beta_log_kernel<<<gridB, threads, smB, s>>>(targets, target_lengths, input_lengths, T, Sp, Sa, nll, w);
        HCTR_CUDA(cudaGetLastError());
    }
    if (relative) {
        ctc_nll_rel_kernel<<<B, 256, 0, s>>>(input_lengths, T, nll, w);
        HCTR_CUDA(cudaGetLastError());
    }
    ctc_mean_loss_kernel<<<1, 32, 0, s>>>(nll, w.len, w.err, B, loss);
    HCTR_CUDA(cudaGetLastError());
    if (grad != nullptr) {
        if (overlap) {
            const size_t smF = (size_t)kFixWarps * Sp * sizeof(float);
            const int blocksF = (int)((rows + kFixWarps - 1) / kFixWarps);
            if (dtype == HCTR_F32)
                ctc_fix_kernel<float><<<blocksF, kFixWarps * 32, smF, s>>>(static_cast<float*>(grad), T, B, C, stride_t, stride_b,
                                                                         input_lengths, Sp, Sa, kscan, grad_scale, relative ? 1 : 0, w);
            else
                ctc_fix_kernel<__nv_bfloat16><<<blocksF, kFixWarps * 32, smF, s>>>(static_cast<__nv_bfloat16*>(grad), T, B, C, stride_t,
                                                                                 stride_b, input_lengths, Sp, Sa, kscan, grad_scale,
                                                                                 relative ? 1 : 0, w);
        } else {
            const size_t smC = (size_t)Sp * sizeof(float);
            if (dtype == HCTR_F32)
                ctc_grad_kernel<float><<<(int)rows, kGradThreads, smC, s>>>(
                    static_cast<const float*>(logits), static_cast<float*>(grad), T, B, C, stride_t, stride_b, targets,
                    target_lengths, input_lengths, Sp, Sa, kscan, grad_scale, w);
            else
                ctc_grad_kernel<__nv_bfloat16><<<(int)rows, kGradThreads, smC, s>>>(
                    static_cast<const __nv_bfloat16*>(logits), static_cast<__nv_bfloat16*>(grad), T, B, C, stride_t, stride_b,
                    targets, target_lengths, input_lengths, Sp, Sa, kscan, grad_scale, w);
        }
        HCTR_CUDA(cudaGetLastError());
    }
    return HCTR_OK;
}

}  // extern "C"
