// CTC loss forward/backward fused with log-softmax (reference call: main.py:205,406-409,
//   criterion = CTCLoss(zero_infinity=True); loss = criterion(preds.log_softmax(2), targets, T*B, lengths)).
// Three passes:
//   A (HBM-bound): one warp per (t,b) row -> log-sum-exp, and the gathered log-probs of the blank-interleaved
//                  label sequence l' (S = 2L+1) for the recursion.
//   B (latency-bound): two CTAs per sequence run the alpha and the beta recursion CONCURRENTLY, state in shared memory
//                  (ping-pong, one barrier per step), 3-way log-sum-exp as ATen's ctc_loss. The state is fp64 -
//                  |log alpha| reaches ~2e4 at T=2048, C=7375, where fp32 log-space resolves only ~2e-3 - but the
//                  exp/log act on O(1) differences and run in fp32 (fast SFU path), which keeps a step at ~0.2 us.
//   C (HBM-bound): one CTA per row: grad = (softmax - occupancy) * scale written in one pass, where
//                  occupancy_c = sum_{s: l'_s = c} exp(alpha_t(s) + beta_t(s) - ll - lp[t, l'_s]).
// Logits are read twice and the gradient written once: (2*s_in + s_out) * T*B*C bytes.
#include <cfloat>

#include "common.cuh"
#include "../../include/hctr_b200.h"

namespace hctr {

struct CtcWs {
    float* lse;        // [B][T]
    float* lpg;        // [B][T][Smax]   log-prob of l'_s at (t,b)
    double* alpha;     // [B][T][Smax]   alpha (fp64 state)
    double* beta;      // [B][T][Smax]   beta (fp64 state); occupancy = exp(alpha + beta - lp - ll)
    double* ll;        // [B]            log-likelihood (may be -inf), fp64
    int* canon;        // [B][Smax]      first s' with the same class as s
    int* toff;         // [B]            offset of sequence b in the concatenated targets
};

__host__ __device__ inline long long align_up(long long v, long long a) { return (v + a - 1) / a * a; }

static CtcWs carve(void* base, int T, int B, int Smax, long long* total) {
    long long off = 0;
    auto take = [&](long long bytes) { long long o = off; off = align_up(off + bytes, 256); return o; };
    const long long o_lse = take(4ll * B * T);
    const long long o_lpg = take(4ll * B * T * Smax);
    const long long o_alpha = take(8ll * B * T * Smax);
    const long long o_beta = take(8ll * B * T * Smax);
    const long long o_ll = take(8ll * B);
    const long long o_canon = take(4ll * B * Smax);
    const long long o_toff = take(4ll * B);
    if (total) *total = off;
    CtcWs w;
    char* p = static_cast<char*>(base);
    w.lse = reinterpret_cast<float*>(p + o_lse);
    w.lpg = reinterpret_cast<float*>(p + o_lpg);
    w.alpha = reinterpret_cast<double*>(p + o_alpha);
    w.beta = reinterpret_cast<double*>(p + o_beta);
    w.ll = reinterpret_cast<double*>(p + o_ll);
    w.canon = reinterpret_cast<int*>(p + o_canon);
    w.toff = reinterpret_cast<int*>(p + o_toff);
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

// ---------------------------------------------------------------- prep: target offsets + canonical states
__global__ void ctc_prep_kernel(const int32_t* __restrict__ targets, const int32_t* __restrict__ tlen, int B, int Smax,
                                CtcWs w) {
    __shared__ int s_off;
    const int b = blockIdx.x;
    if (threadIdx.x == 0) {
        int off = 0;
        for (int i = 0; i < b; ++i) off += tlen[i];
        s_off = off; w.toff[b] = off;
    }
    __syncthreads();
    const int L = tlen[b], S = 2 * L + 1;
    const int32_t* tg = targets + s_off;
    for (int s = threadIdx.x; s < S; s += blockDim.x) {
        const int c = (s & 1) ? tg[s >> 1] : 0;
        int first = s;
        for (int q = (s & 1) ? 1 : 0; q < s; q += ((c == 0) ? 2 : 1)) {
            const int cq = (q & 1) ? tg[q >> 1] : 0;
            if (cq == c) { first = q; break; }
        }
        w.canon[(long long)b * Smax + s] = first;
    }
}

// ---------------------------------------------------------------- pass A: row log-sum-exp + label gather
constexpr int kLseWarps = 8;

template <typename T>
__global__ void __launch_bounds__(kLseWarps * 32)
ctc_lse_gather_kernel(const T* __restrict__ logits, int Tn, int Bn, int C, long long stride_t, long long stride_b,
                      const int32_t* __restrict__ targets, const int32_t* __restrict__ tlen,
                      const int32_t* __restrict__ ilen, int Smax, const float* __restrict__ lse_in, CtcWs w) {
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
    const int L = tlen[b], S = 2 * L + 1;
    const int32_t* tg = targets + w.toff[b];
    float* dst = w.lpg + row * Smax;
    for (int q = lane; q < S; q += 32) {
        const int c = (q & 1) ? tg[q >> 1] : 0;
        dst[q] = Ld<T>::one(p + c) - lse;
    }
}

// ---------------------------------------------------------------- pass B: alpha / beta recursion
__device__ __forceinline__ double lse3(double a, double b, double c) {
    const double mx = fmax(a, fmax(b, c));
    if (mx == -INFINITY) return -INFINITY;
    // the differences are <= 0 and O(1): fp32 transcendentals lose nothing that matters, fp64 keeps the large offset
    const float s = __expf((float)(a - mx)) + __expf((float)(b - mx)) + __expf((float)(c - mx));   // ex2.approx: 2 ulp
    return mx + (double)logf(s);
}

// grid (B, 2): blockIdx.y == 0 runs alpha (and the log-likelihood), blockIdx.y == 1 runs beta.
__global__ void __launch_bounds__(1024)
ctc_alpha_beta_kernel(const int32_t* __restrict__ targets, const int32_t* __restrict__ tlen,
                      const int32_t* __restrict__ ilen, int Tn, int Smax, float* __restrict__ nll_out, CtcWs w) {
    extern __shared__ double smd[];               // 2 x (Smax + 4) ping-pong state with -inf guards, + 1 scratch
    const int b = blockIdx.x, s = threadIdx.x;
    const bool is_beta = blockIdx.y == 1;
    const int L = tlen[b], S = 2 * L + 1, Tb = ilen[b];
    const int32_t* tg = targets + w.toff[b];
    const int W = Smax + 4;
    double* bufA = smd;
    double* bufB = smd + W;
    const bool act = s < S;
    const int cls = act ? ((s & 1) ? tg[s >> 1] : 0) : 0;
    const bool skip_in = act && s > 1 && cls != 0 && cls != ((s & 1) ? tg[(s >> 1) - 1] : 0);      // s-2 -> s allowed
    const bool skip_out = act && (s + 2 < S) && (((s + 2) & 1) ? tg[(s + 2) >> 1] : 0) != 0 &&
                          ((((s + 2) & 1) ? tg[(s + 2) >> 1] : 0) != cls);                              // s -> s+2 allowed
    const float* lp = w.lpg + (long long)b * Tn * Smax;

    for (int i = threadIdx.x; i < 2 * W; i += blockDim.x) smd[i] = -INFINITY;
    __syncthreads();
    if (!is_beta) {
        // ---- alpha: state s lives at index s+2 (two -inf guard cells on the left)
        double* al = w.alpha + (long long)b * Tn * Smax;
        double ll = -INFINITY;
        if (Tb > 0) {
            double a = -INFINITY;
            if (act && s < 2) a = (double)lp[s];
            if (act) { bufA[s + 2] = a; al[s] = a; }
            __syncthreads();
            double* cur = bufA; double* nxt = bufB;
            float lp_next = (act && Tb > 1) ? lp[(long long)Smax + s] : 0.f;
            for (int t = 1; t < Tb; ++t) {
                const double lpt = (double)lp_next;
                if (act && t + 1 < Tb) lp_next = lp[(long long)(t + 1) * Smax + s];
                if (act) {
                    const double v = lse3(cur[s + 2], cur[s + 1], skip_in ? cur[s] : -INFINITY);
                    a = (v == -INFINITY) ? -INFINITY : v + lpt;
                    nxt[s + 2] = a;
                    al[(long long)t * Smax + s] = a;
                }
                __syncthreads();
                double* tmp = cur; cur = nxt; nxt = tmp;
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
        double* be = w.beta + (long long)b * Tn * Smax;
        double* cur = bufA; double* nxt = bufB;
        {
            const long long o = (long long)(Tb - 1) * Smax + s;
            double bt = -INFINITY;
            if (act && s >= S - 2) bt = (double)lp[o];
            if (act) { cur[s] = bt; be[o] = bt; }
        }
        __syncthreads();
        float lp_next = (act && Tb > 1) ? lp[(long long)(Tb - 2) * Smax + s] : 0.f;
        for (int t = Tb - 2; t >= 0; --t) {
            const double lpt = (double)lp_next;
            if (act && t > 0) lp_next = lp[(long long)(t - 1) * Smax + s];
            if (act) {
                const double v = lse3(cur[s], cur[s + 1], skip_out ? cur[s + 2] : -INFINITY);
                const double bt = (v == -INFINITY) ? -INFINITY : v + lpt;
                nxt[s] = bt;
                be[(long long)t * Smax + s] = bt;
            }
            __syncthreads();
            double* tmp = cur; cur = nxt; nxt = tmp;
        }
    }
}

// ---------------------------------------------------------------- loss = mean_b(nll_b / max(L_b, 1))
__global__ void ctc_mean_loss_kernel(const float* __restrict__ nll, const int32_t* __restrict__ tlen, int B,
                                     float* __restrict__ loss) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        float acc = 0.f;
        for (int b = 0; b < B; ++b) acc += nll[b] / (float)max(tlen[b], 1);
        loss[0] = acc / (float)B;
    }
}

// ---------------------------------------------------------------- pass C: gradient wrt logits
constexpr int kGradThreads = 256;

template <typename T>
__global__ void __launch_bounds__(kGradThreads)
ctc_grad_kernel(const T* __restrict__ logits, T* __restrict__ grad, int Tn, int Bn, int C, long long stride_t,
                long long stride_b, const int32_t* __restrict__ targets, const int32_t* __restrict__ tlen,
                const int32_t* __restrict__ ilen, int Smax, float grad_scale, CtcWs w) {
    constexpr int V = Ld<T>::N;
    extern __shared__ float occ[];                                   // [Smax] exp(alpha+beta-lp-ll) per state
    const long long row = blockIdx.x;                                // row = b*T + t
    const int b = (int)(row / Tn), t = (int)(row - (long long)b * Tn);
    const T* p = logits + (long long)t * stride_t + (long long)b * stride_b;
    T* g = grad + (long long)t * stride_t + (long long)b * stride_b;
    const int L = tlen[b], S = 2 * L + 1;
    const double ll = w.ll[b];
    const bool dead = (t >= ilen[b]) || (ll == -INFINITY);          // beyond the input length / zero_infinity
    const float scale = dead ? 0.f : grad_scale / ((float)max(L, 1) * (float)Bn);
    const float lse = dead ? 0.f : w.lse[row];

    if (!dead) {
        const double* al = w.alpha + row * Smax;
        const double* be = w.beta + row * Smax;
        const float* lpr = w.lpg + row * Smax;
        for (int s = threadIdx.x; s < S; s += blockDim.x) {
            const double e = al[s] + be[s] - (double)lpr[s] - ll;        // -inf if either side is unreachable
            occ[s] = (e == e) ? expf((float)e) : 0.f;
        }
    }
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
    const int* canon = w.canon + (long long)b * Smax;
    const int32_t* tg = targets + w.toff[b];
    for (int s = threadIdx.x; s < S; s += blockDim.x) {
        if (canon[s] != s) continue;
        float acc = 0.f;
        for (int q = s; q < S; q += ((s & 1) ? 1 : 2))
            if (canon[q] == s) acc += occ[q];
        const int c = (s & 1) ? tg[s >> 1] : 0;
        Ld<T>::st_one(g + c, (__expf(Ld<T>::one(p + c) - lse) - acc) * scale);
    }
}

}  // namespace hctr

using namespace hctr;

extern "C" {

long long hctr_ctc_loss_workspace_bytes(int T, int B, int max_target_len) {
    if (T <= 0 || B <= 0 || max_target_len < 0) return 0;
    long long total = 0;
    carve(nullptr, T, B, 2 * max_target_len + 1, &total);
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
    const int Smax = 2 * max_target_len + 1;
    long long need = 0;
    CtcWs w = carve(workspace, T, B, Smax, &need);
    HCTR_CHECK(workspace && workspace_bytes >= need, HCTR_ERR_INVALID, "ctc_loss: workspace too small (%lld < %lld)", workspace_bytes, need);
    HCTR_CHECK((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, HCTR_ERR_INVALID, "ctc_loss: workspace must be 256-byte aligned");
    cudaStream_t s = static_cast<cudaStream_t>(stream);

    ctc_prep_kernel<<<B, 128, 0, s>>>(targets, target_lengths, B, Smax, w);
    HCTR_CUDA(cudaGetLastError());
    const long long rows = (long long)T * B;
    const long long blocksA = (rows + kLseWarps - 1) / kLseWarps;
    HCTR_CHECK(rows < (1ll << 31), HCTR_ERR_INVALID, "ctc_loss: too many rows");
    if (dtype == HCTR_F32)
        ctc_lse_gather_kernel<float><<<(int)blocksA, kLseWarps * 32, 0, s>>>(
            static_cast<const float*>(logits), T, B, C, stride_t, stride_b, targets, target_lengths, input_lengths, Smax, row_lse, w);
    else
        ctc_lse_gather_kernel<__nv_bfloat16><<<(int)blocksA, kLseWarps * 32, 0, s>>>(
            static_cast<const __nv_bfloat16*>(logits), T, B, C, stride_t, stride_b, targets, target_lengths, input_lengths, Smax, row_lse, w);
    HCTR_CUDA(cudaGetLastError());
    int threads = (Smax + 31) / 32 * 32;
    const size_t smB = (size_t)(2 * (Smax + 4) + 2) * sizeof(double);
    ctc_alpha_beta_kernel<<<dim3(B, grad != nullptr ? 2 : 1), threads, smB, s>>>(targets, target_lengths, input_lengths, T, Smax, nll, w);
    HCTR_CUDA(cudaGetLastError());
    ctc_mean_loss_kernel<<<1, 32, 0, s>>>(nll, target_lengths, B, loss);
    HCTR_CUDA(cudaGetLastError());
    if (grad != nullptr) {
        const size_t smC = (size_t)Smax * sizeof(float);
        if (dtype == HCTR_F32)
            ctc_grad_kernel<float><<<(int)rows, kGradThreads, smC, s>>>(
                static_cast<const float*>(logits), static_cast<float*>(grad), T, B, C, stride_t, stride_b, targets,
                target_lengths, input_lengths, Smax, grad_scale, w);
        else
            ctc_grad_kernel<__nv_bfloat16><<<(int)rows, kGradThreads, smC, s>>>(
                static_cast<const __nv_bfloat16*>(logits), static_cast<__nv_bfloat16*>(grad), T, B, C, stride_t, stride_b,
                targets, target_lengths, input_lengths, Smax, grad_scale, w);
        HCTR_CUDA(cudaGetLastError());
    }
    return HCTR_OK;
}

}  // extern "C"
