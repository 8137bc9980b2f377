// CTC greedy decode on device (reference: ctc_codec.__greedy_search__, utils/ctc_codec.py:70-99).
//   pass 1 (HBM-bound): one warp per (t,b) row, 16-byte vector loads, warp-shuffle arg-max with
//                       numpy.argmax semantics (first maximum; a NaN beats everything, first NaN wins).
//   pass 2: one block per sequence, blank / unknown / repeat collapse with a block-wide scan.
// The logits never leave the device; only the compact [B][<=T] index array is copied to host.
#include "common.cuh"
#include "../../include/hctr_b200.h"

namespace hctr {

struct Best {
    float v;     // value with NaN mapped to +inf rank via `nan`
    int i;
    int nan;
};

__device__ __forceinline__ void consider(Best& b, float v, int i) {
    // elements are visited in increasing index order inside a thread: strict '>' keeps the first max
    const int vn = (v != v);
    if (b.i < 0) { b.v = v; b.i = i; b.nan = vn; return; }
    if (b.nan) return;                       // an earlier NaN already won
    if (vn || v > b.v) { b.v = v; b.i = i; b.nan = vn; }
}

__device__ __forceinline__ bool better(const Best& a, const Best& b) {
    // true if a should replace b when merging lanes (order-independent total order)
    if (a.i < 0) return false;
    if (b.i < 0) return true;
    if (a.nan != b.nan) return a.nan > b.nan;
    if (a.nan) return a.i < b.i;
    if (a.v != b.v) return a.v > b.v;
    return a.i < b.i;
}

template <typename T> struct Vec;
template <> struct Vec<float> {
    static constexpr int N = 4;
    static __device__ __forceinline__ void load(const float* p, float (&o)[4]) {
        const uint4 q = ld_nc_v4(p);
        o[0] = __uint_as_float(q.x); o[1] = __uint_as_float(q.y); o[2] = __uint_as_float(q.z); o[3] = __uint_as_float(q.w);
    }
    static __device__ __forceinline__ float one(const float* p) { return __ldg(p); }
};
template <> struct Vec<__nv_bfloat16> {
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

constexpr int kArgmaxWarps = 8;
constexpr int kUnroll = 8;

// ---- exact (slow) row scan: numpy.argmax semantics incl. NaN; used for the unaligned head/tail and for
//      rows in which the fast path saw a NaN.
template <typename T>
__device__ __noinline__ Best row_argmax_exact(const T* __restrict__ p, int C, int lane) {
    Best best; best.v = 0.f; best.i = -1; best.nan = 0;
    for (int c = lane; c < C; c += 32) consider(best, Vec<T>::one(p + c), c);
    return best;
}

// ---- fast per-thread scan of the 16-byte-aligned body: only a running maximum and the index of the VECTOR
//      that produced it (strict '>' keeps the earliest); the element is located afterwards. ~1 instruction per
//      element, which is what lets the kernel run at HBM speed (the budget is ~8 issue slots per bf16 element).
__device__ __forceinline__ float vec_max_nanprop(const uint4& q, __nv_bfloat16*) {
    // packed bf16x2 NaN-propagating max tree
    const __nv_bfloat162 a = __hmax2_nan(*reinterpret_cast<const __nv_bfloat162*>(&q.x), *reinterpret_cast<const __nv_bfloat162*>(&q.y));
    const __nv_bfloat162 b = __hmax2_nan(*reinterpret_cast<const __nv_bfloat162*>(&q.z), *reinterpret_cast<const __nv_bfloat162*>(&q.w));
    const __nv_bfloat162 c = __hmax2_nan(a, b);
    const uint32_t u = *reinterpret_cast<const uint32_t*>(&c);
    const float lo = bf16_lo(u), hi = bf16_hi(u);
    return (lo != lo || hi != hi) ? __int_as_float(0x7fc00000) : fmaxf(lo, hi);
}
__device__ __forceinline__ float vec_max_nanprop(const uint4& q, float*) {
    const float x0 = __uint_as_float(q.x), x1 = __uint_as_float(q.y), x2 = __uint_as_float(q.z), x3 = __uint_as_float(q.w);
    const float m = fmaxf(fmaxf(x0, x1), fmaxf(x2, x3));
    return (x0 != x0 || x1 != x1 || x2 != x2 || x3 != x3) ? __int_as_float(0x7fc00000) : m;
}

template <typename T>
__global__ void __launch_bounds__(kArgmaxWarps * 32)
ctc_argmax_kernel(const T* __restrict__ logits, int Tn, int Bn, int C, long long stride_t, long long stride_b,
                  int32_t* __restrict__ argmax_bt) {
    constexpr int V = Vec<T>::N;
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * kArgmaxWarps + (threadIdx.x >> 5);   // row = b*T + t
    if (row >= (long long)Tn * Bn) return;
    const int b = (int)(row / Tn), t = (int)(row - (long long)b * Tn);
    const T* p = logits + (long long)t * stride_t + (long long)b * stride_b;

    const uintptr_t addr = reinterpret_cast<uintptr_t>(p);
    int head = (int)(((16 - (addr & 15)) & 15) / sizeof(T));
    if (head > C) head = C;
    const int nvec = (C - head) / V;
    const T* pv = p + head;
    const int tail0 = head + nvec * V;

    float bm = -INFINITY;       // running max over this thread's vectors
    int bvi = -1;               // vector index that produced it
    bool saw_nan = false;
    int vi = lane;
    for (; vi + (kUnroll - 1) * 32 < nvec; vi += kUnroll * 32) {
        uint4 q[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) q[u] = ld_nc_v4(pv + (long long)(vi + u * 32) * V);
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const float m = vec_max_nanprop(q[u], static_cast<T*>(nullptr));
            if (!(m <= bm)) {                       // greater, or NaN
                if (m != m) saw_nan = true; else { bm = m; bvi = vi + u * 32; }
            }
        }
    }
    for (; vi < nvec; vi += 32) {
        const uint4 q = ld_nc_v4(pv + (long long)vi * V);
        const float m = vec_max_nanprop(q, static_cast<T*>(nullptr));
        if (!(m <= bm)) {
            if (m != m) saw_nan = true; else { bm = m; bvi = vi; }
        }
    }
    Best best; best.v = 0.f; best.i = -1; best.nan = 0;
    if (__any_sync(0xffffffffu, saw_nan)) {
        best = row_argmax_exact<T>(p, C, lane);      // rare: exact NaN-aware scan of the whole row
    } else {
        // head elements (lowest indices) first, then the located body element, then the tail: increasing index order
        if (lane < head) consider(best, Vec<T>::one(p + lane), lane);
        if (bvi >= 0) {
            float x[V];
            Vec<T>::load(pv + (long long)bvi * V, x);
#pragma unroll
            for (int j = 0; j < V; ++j) consider(best, x[j], head + bvi * V + j);
        } else if (nvec > 0 && lane < nvec) {
            // every vector of this thread was all -inf: the first of them holds the thread's first maximum
            consider(best, -INFINITY, head + lane * V);
        }
        if (tail0 + lane < C) consider(best, Vec<T>::one(p + tail0 + lane), tail0 + lane);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        Best other;
        other.v = __shfl_xor_sync(0xffffffffu, best.v, o);
        other.i = __shfl_xor_sync(0xffffffffu, best.i, o);
        other.nan = __shfl_xor_sync(0xffffffffu, best.nan, o);
        if (better(other, best)) best = other;
    }
    if (lane == 0) argmax_bt[row] = best.i;
}

// One block per sequence: keep t iff idx[t] != blank && idx[t] != unknown && !(t>0 && idx[t-1]==idx[t]).
__global__ void __launch_bounds__(256)
ctc_collapse_kernel(const int32_t* __restrict__ argmax_bt, int Tn, int unknown, int32_t* __restrict__ out_idx,
                    int32_t* __restrict__ out_len) {
    __shared__ int warp_sums[8];
    __shared__ int carry;
    const int b = blockIdx.x;
    const int32_t* a = argmax_bt + (long long)b * Tn;
    int32_t* o = out_idx + (long long)b * Tn;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < Tn; base += blockDim.x) {
        const int t = base + threadIdx.x;
        int cur = 0, keep = 0;
        if (t < Tn) {
            cur = a[t];
            keep = (cur != 0) && (cur != unknown) && !(t > 0 && a[t - 1] == cur);
        }
        const unsigned ballot = __ballot_sync(0xffffffffu, keep);
        const int prefix = __popc(ballot & ((1u << lane) - 1));
        if (lane == 0) warp_sums[warp] = __popc(ballot);
        __syncthreads();
        int off = carry;
        for (int w2 = 0; w2 < warp; ++w2) off += warp_sums[w2];
        if (keep) o[off + prefix] = cur;
        __syncthreads();
        if (threadIdx.x == 0) {
            int tot = 0;
            for (int w2 = 0; w2 < (int)(blockDim.x >> 5); ++w2) tot += warp_sums[w2];
            carry += tot;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) out_len[b] = carry;
}

}  // namespace hctr

using namespace hctr;

extern "C" int hctr_ctc_greedy_decode(const void* logits, int dtype, int T, int B, int C, long long stride_t,
                                      long long stride_b, int32_t* argmax_out, int32_t* out_idx, int32_t* out_len,
                                      void* stream) {
    HCTR_CHECK(T >= 0 && B >= 0 && C > 0, HCTR_ERR_INVALID, "greedy: bad shape T=%d B=%d C=%d", T, B, C);
    HCTR_CHECK(dtype == HCTR_F32 || dtype == HCTR_BF16, HCTR_ERR_INVALID, "greedy: bad dtype %d", dtype);
    if (B == 0) return HCTR_OK;
    HCTR_CHECK(out_len != nullptr, HCTR_ERR_INVALID, "greedy: null output");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (T == 0) {   // reference: a zero-length sample yields no text (utils/ctc_codec.py:85-86)
        HCTR_CUDA(cudaMemsetAsync(out_len, 0, sizeof(int32_t) * B, s));
        return HCTR_OK;
    }
    HCTR_CHECK(logits != nullptr && out_idx != nullptr, HCTR_ERR_INVALID, "greedy: null logits / output");
    HCTR_CHECK(argmax_out != nullptr, HCTR_ERR_INVALID, "greedy: argmax workspace [B][T] is required");
    const long long rows = (long long)T * B;
    const long long blocks = (rows + kArgmaxWarps - 1) / kArgmaxWarps;
    HCTR_CHECK(blocks < (1ll << 31), HCTR_ERR_INVALID, "greedy: too many rows");
    if (dtype == HCTR_F32)
        ctc_argmax_kernel<float><<<(int)blocks, kArgmaxWarps * 32, 0, s>>>(static_cast<const float*>(logits), T, B, C,
                                                                          stride_t, stride_b, argmax_out);
    else
        ctc_argmax_kernel<__nv_bfloat16><<<(int)blocks, kArgmaxWarps * 32, 0, s>>>(
            static_cast<const __nv_bfloat16*>(logits), T, B, C, stride_t, stride_b, argmax_out);
    HCTR_CUDA(cudaGetLastError());
    ctc_collapse_kernel<<<B, 256, 0, s>>>(argmax_out, T, C - 1, out_idx, out_len);
    HCTR_CUDA(cudaGetLastError());
    return HCTR_OK;
}

extern "C" int hctr_ctc_collapse(const int32_t* argmax_bt, int T, int B, int C, int32_t* out_idx, int32_t* out_len, void* stream) {
    HCTR_CHECK(T >= 0 && B >= 0 && C > 0, HCTR_ERR_INVALID, "collapse: bad shape T=%d B=%d C=%d", T, B, C);
    if (B == 0) return HCTR_OK;
    HCTR_CHECK(out_len != nullptr, HCTR_ERR_INVALID, "collapse: null output");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (T == 0) {
        HCTR_CUDA(cudaMemsetAsync(out_len, 0, sizeof(int32_t) * B, s));
        return HCTR_OK;
    }
    HCTR_CHECK(argmax_bt != nullptr && out_idx != nullptr, HCTR_ERR_INVALID, "collapse: null pointer");
    ctc_collapse_kernel<<<B, 256, 0, s>>>(argmax_bt, T, C - 1, out_idx, out_len);
    HCTR_CUDA(cudaGetLastError());
    return HCTR_OK;
}
