// CTC prefix beam search on device (reference: ctc_codec.decode beam branch, utils/ctc_codec.py:63-67,
// __cbs_full__ :183-210, __context_beam_search__ :212-285, Beam :288-307).
//   kernel 1 (HBM-bound): per (t,b) row, one warp: online log-sum-exp + per-lane sorted top-k in registers,
//                         merged across the warp; emits candidates in descending log-prob order.
//   kernel 2 (latency-bound): one 128-thread CTA per sequence walks the time steps, every step fully
//                         parallel over (beam, candidate) pairs; float64 accumulators, np.logaddexp
//                         semantics, insertion-ordered stable ranking exactly as the reference's dict +
//                         sorted(reverse=True).
#include <cfloat>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "row_stage.cuh"
#include "ngram_lm.cuh"
#include "../../include/hctr_b200.h"

namespace hctr {

constexpr int kMaxK = 16;        // search_depth limit
constexpr int kMaxBeam = 16;     // beam_size limit
constexpr int kMaxGen = kMaxBeam * (kMaxK + 1);

// ---------------------------------------------------------------------------------------------- top-k
__device__ __forceinline__ bool cand_better(float v, int i, float w, int j) {
    return v > w || (v == w && i < j);
}

// exp(x) for x <= 0 as one multiply and one MUFU: ex2.approx.ftz flushes results below 2^-126 to zero, which is what a
// sum of exponentials wants (__expf adds a range check and two scalings per element for the denormal range)
__device__ __forceinline__ float exp_neg_fast(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x * 1.4426950408889634f));
    return y;
}

template <typename T> struct LoadVec;
template <> struct LoadVec<float> {
    static constexpr int N = 4;
    static __device__ __forceinline__ void load(const float* p, float (&o)[4]) {
        const uint4 q = ld_nc_v4(p);
        o[0] = __uint_as_float(q.x); o[1] = __uint_as_float(q.y); o[2] = __uint_as_float(q.z); o[3] = __uint_as_float(q.w);
    }
    static __device__ __forceinline__ float one(const float* p) { return __ldg(p); }
    static __device__ __forceinline__ float one_smem(const float* p) { return *p; }
    static __device__ __forceinline__ void load_smem(const unsigned char* p, float (&o)[4]) {
        const float4 q = *reinterpret_cast<const float4*>(p);
        o[0] = q.x; o[1] = q.y; o[2] = q.z; o[3] = q.w;
    }
};
template <> struct LoadVec<__nv_bfloat16> {
    static constexpr int N = 8;
    static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&o)[8]) {
        const uint4 q = ld_nc_v4(p);
        o[0] = bf16_lo(q.x); o[1] = bf16_hi(q.x); o[2] = bf16_lo(q.y); o[3] = bf16_hi(q.y);
        o[4] = bf16_lo(q.z); o[5] = bf16_hi(q.z); o[6] = bf16_lo(q.w); o[7] = bf16_hi(q.w);
    }
    static __device__ __forceinline__ float one(const __nv_bfloat16* p) {
        return __uint_as_float(static_cast<uint32_t>(*reinterpret_cast<const unsigned short*>(p)) << 16);
    }
    static __device__ __forceinline__ float one_smem(const __nv_bfloat16* p) { return one(p); }
    static __device__ __forceinline__ void load_smem(const unsigned char* p, float (&o)[8]) {
        const uint4 q = *reinterpret_cast<const uint4*>(p);
        o[0] = bf16_lo(q.x); o[1] = bf16_hi(q.x); o[2] = bf16_lo(q.y); o[3] = bf16_hi(q.y);
        o[4] = bf16_lo(q.z); o[5] = bf16_hi(q.z); o[6] = bf16_lo(q.w); o[7] = bf16_hi(q.w);
    }
};

// Persistent CTAs of 256 threads; a CTA walks rows (t,b) = blockIdx.x, +gridDim.x, ... and keeps TWO row buffers in
// shared memory: while row i is reduced, the bulk-copy engine (cp.async.bulk global -> shared, mbarrier complete_tx)
// already streams row i+1, so HBM latency is never on a row's critical path and the SM needs no spare occupancy to hide it.
// A row is staged in its own dtype (bf16 rows cost half the shared memory and bandwidth). Rows need not be 16-byte
// aligned (C = 7375 floats): the bulk copy moves the 16-byte-aligned interior [lo, hi) of the row to offset 16 of the
// buffer, the < 16-byte head and tail land next to it through registers, fetched one row ahead as well.
// The kernel is bound by instruction issue, not by latency (first persistent version: 38 instructions per element, 56 %
// issue-active with 24 warps/SM, slower than HBM allows), so the two passes work on 16-byte shared-memory vectors and
// everything per row is kept short:
//   pass 1  per-thread maximum (ld.shared.v4 + a max tree); maxima of 16 thread groups (half-warps at 256 threads per row,
//           quarter-warps at 128) -> shared memory
//   bound   the k-th largest of the 16 group maxima is a valid lower bound tau for the k-th largest element of the
//           row (k distinct elements are >= it; ~15 elements of a random row reach it); each warp ranks two of the 16 values
//           (one compare per lane + a ballot); the row maximum falls out of the same ranking
//   pass 2  sum exp(x - max) and collect the few elements >= tau (one compare per vector on its maximum)
//   rank    candidates ranked by counting (value desc, index asc) -> the top k in order; exact for any input.
// If more than kMaxCand elements reach tau (e.g. constant rows) an exact k-round arg-max fallback runs instead.
// Four block barriers per row, no single-thread sections, no division per row.
// Measured and dropped: a register-resident variant (two warps per row, 29 independent 16-byte loads per lane, both passes
// on registers, no staging) executes a third fewer instructions but needs 221 registers - 8 warps per SM - and a fully
// unrolled 2000-instruction body per row: 27 % issue-active (fixed-latency and instruction-fetch stalls), 1.9 ms against
// 1.2 ms for this kernel on the config-5 tensor.
constexpr int kMaxCand = 512;
constexpr int kTopkEdge = 16;          // head + tail elements of a row that do not belong to the aligned interior (< 16 B each)

// geometry of one row in global memory / in its shared-memory buffer (element 0 lives at byte 16 - head_bytes)
struct TopkRow {
    const unsigned char* a0;           // first byte of the row
    int head_bytes;                    // [a0, lo): bytes in front of the 16-byte-aligned interior
    int body_bytes;                    // [lo, hi): multiple of 16 (0 for tiny rows: everything travels through registers)
};
template <typename T>
__device__ __forceinline__ TopkRow topk_row(const T* logits, unsigned t, unsigned b, long long stride_t, long long stride_b,
                                            int row_bytes) {
    TopkRow r;
    r.a0 = reinterpret_cast<const unsigned char*>(logits + (long long)t * stride_t + (long long)b * stride_b);
    const uintptr_t a = reinterpret_cast<uintptr_t>(r.a0);
    const uintptr_t lo = (a + 15) & ~uintptr_t(15), hi = (a + row_bytes) & ~uintptr_t(15);
    if (hi > lo) { r.head_bytes = (int)(lo - a); r.body_bytes = (int)(hi - lo); }
    else { r.head_bytes = 0; r.body_bytes = 0; }
    return r;
}

// THREADS = 256 (round 1: two row buffers per CTA, three CTAs per SM) or 128 (round 2 default: ONE row buffer per CTA, seven
// CTAs per SM at C = 7375 fp32 - the other CTAs' bulk copies are in flight while one reduces). Half the threads per row double
// every thread's share of the row (58 instead of 29 elements), so the per-row bookkeeping - barriers, ranking, staging
// protocol - weighs half as much per element. Config-5 tensor, fp32: 1.00 ms (0.59 of the HBM peak) -> 0.79 ms (0.75);
// 64 threads per row: 1.00 ms; 128 threads with two buffers (three CTAs per SM): 1.19 ms.
template <typename T, int THREADS>
__global__ void __launch_bounds__(THREADS)
ctc_topk_logsoftmax_kernel(const T* __restrict__ logits, long long rows, int Bn, int C, long long stride_t, long long stride_b,
                           int k, int nbuf, int buf_bytes, int32_t* __restrict__ topk_idx, float* __restrict__ topk_logp,
                           float* __restrict__ lse_out) {
    constexpr int V = LoadVec<T>::N;
    extern __shared__ __align__(128) unsigned char rowbufs[];      // [nbuf][buf_bytes]
    __shared__ __align__(8) uint64_t full_bar[2];
    __shared__ float gmax[16], red_sum[THREADS / 32];
    constexpr int GRP = THREADS / 16;                 // lanes per group: 16 group maxima bound the k-th largest element
    __shared__ float s_tau, s_m;
    __shared__ int s_ncand[2];
    __shared__ float cand_v[kMaxCand];
    __shared__ int cand_i[kMaxCand];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int row_bytes = C * (int)sizeof(T);
    const int kk = k < 16 ? k : 16;

    // thread 0: bulk copy of the interior of row r into buffer s
    auto issue = [&](const TopkRow& r, int s) {
        if (r.body_bytes > 0) {
            mbar_arrive_expect_tx(&full_bar[s], (uint32_t)r.body_bytes);
            bulk_g2s(rowbufs + (size_t)s * buf_bytes + 16, r.a0 + r.head_bytes, (uint32_t)r.body_bytes, &full_bar[s]);
        } else {
            mbar_arrive(&full_bar[s]);
        }
    };
    // threads < kTopkEdge: byte offset inside the row of the head/tail element this thread carries (-1: none)
    auto edge_offset = [&](const TopkRow& r) -> int {
        const int nhead = r.head_bytes / (int)sizeof(T);
        const int off = tid < nhead ? tid * (int)sizeof(T) : r.head_bytes + r.body_bytes + (tid - nhead) * (int)sizeof(T);
        return (tid < kTopkEdge && off < row_bytes) ? off : -1;
    };

    if (tid == 0) {
        mbar_init(&full_bar[0], 1); mbar_init(&full_bar[1], 1);
        fence_barrier_init();
        s_ncand[0] = 0; s_ncand[1] = 0;
    }
    __syncthreads();
    long long row = blockIdx.x;
    // (t, b) of the rows this CTA walks, advanced without a division: row += gridDim.x
    const unsigned step_t = gridDim.x / (unsigned)Bn, step_b = gridDim.x % (unsigned)Bn;
    unsigned row_t = blockIdx.x / (unsigned)Bn, row_b = blockIdx.x % (unsigned)Bn;
    TopkRow cur = topk_row(logits, row < rows ? row_t : 0u, row < rows ? row_b : 0u, stride_t, stride_b, row_bytes);
    if (row < rows) {
        if (tid == 0) issue(cur, 0);
        const int off = edge_offset(cur);
        if (off >= 0) *reinterpret_cast<T*>(rowbufs + 16 - cur.head_bytes + off) = *reinterpret_cast<const T*>(cur.a0 + off);
    }
    __syncthreads();

    for (int it = 0; row < rows; row += gridDim.x, ++it) {
        const int s = nbuf == 2 ? (it & 1) : 0;
        const long long next = row + gridDim.x;
        // ---- row i+1: bulk copy into the other buffer (its last reader finished before the barrier that ended row i-1)
        TopkRow nxt = cur;
        T edge_val = T();
        int edge_off = -1;
        if (next < rows) {
            row_t += step_t; row_b += step_b;
            if (row_b >= (unsigned)Bn) { row_b -= (unsigned)Bn; ++row_t; }
            nxt = topk_row(logits, row_t, row_b, stride_t, stride_b, row_bytes);
            if (nbuf == 2) {
                if (tid == 0) issue(nxt, s ^ 1);
                edge_off = edge_offset(nxt);
                if (edge_off >= 0) edge_val = *reinterpret_cast<const T*>(nxt.a0 + edge_off);
            }
        }
        unsigned char* buf = rowbufs + (size_t)s * buf_bytes;
        const T* rb = reinterpret_cast<const T*>(buf + 16 - cur.head_bytes);          // element 0 of the row
        const int nhead = cur.head_bytes / (int)sizeof(T);
        const int nvec = cur.body_bytes >> 4;
        const int tail0 = nhead + nvec * V;
        // the scalar (head/tail) element of this thread, if any
        const int sc = tid < nhead ? tid : tail0 + (tid - nhead);
        const bool has_sc = tid < kTopkEdge && sc < C;
        mbar_wait(&full_bar[s], (nbuf == 2 ? (it >> 1) : it) & 1);

        // ---- pass 1: per-thread max
        float tmax = has_sc ? LoadVec<T>::one_smem(rb + sc) : -INFINITY;
        if (sizeof(T) == 2) {
            // bf16 rows stay packed: 4 HMNMX2 per 8 elements, two independent chains
            __nv_bfloat162 a0 = __float2bfloat162_rn(-INFINITY), a1 = a0;
#pragma unroll 2
            for (int vi = tid; vi < nvec; vi += THREADS) {
                const uint4 q = *reinterpret_cast<const uint4*>(buf + 16 + (vi << 4));
                a0 = __hmax2(a0, __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&q.x), *reinterpret_cast<const __nv_bfloat162*>(&q.y)));
                a1 = __hmax2(a1, __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&q.z), *reinterpret_cast<const __nv_bfloat162*>(&q.w)));
            }
            a0 = __hmax2(a0, a1);
            tmax = fmaxf(tmax, fmaxf(__low2float(a0), __high2float(a0)));
        } else {
#pragma unroll 2
            for (int vi = tid; vi < nvec; vi += THREADS) {
                float x[V];
                LoadVec<T>::load_smem(buf + 16 + (vi << 4), x);
#pragma unroll
                for (int j = 0; j < V; j += 2) tmax = fmaxf(tmax, fmaxf(x[j], x[j + 1]));
            }
        }
        float hmax = tmax;                                           // maximum of this half-warp
#pragma unroll
        for (int o = GRP / 2; o > 0; o >>= 1) hmax = fmaxf(hmax, __shfl_xor_sync(0xffffffffu, hmax, o));
        if ((tid & (GRP - 1)) == 0) gmax[tid / GRP] = hmax;
        if (tid == 0) s_ncand[(it + 1) & 1] = 0;
        __syncthreads();
        // rank the 16 group maxima (value desc, index asc; a permutation of 0..15): pair pr = (value pr >> 4, other value
        // pr & 15), so a half-warp ranks one value with one compare per lane and a ballot; tau = the kk-th largest, m = the largest
        for (int pr = tid; pr < 256; pr += THREADS) {     // pair (value vi, other value pr & 15); a half-warp ranks one value
            const int vi = pr >> 4;
            const float gv = gmax[vi], o = gmax[pr & 15];
            const bool ahead = o > gv || (o == gv && (pr & 15) < vi);
            const unsigned bal = __ballot_sync(0xffffffffu, ahead);
            const int grank = __popc(lane < 16 ? (bal & 0xffffu) : (bal >> 16));
            if ((lane & 15) == 0) {
                if (grank == kk - 1) s_tau = gv;
                if (grank == 0) s_m = gv;
            }
        }
        __syncthreads();
        const float tau = s_tau, m = s_m;

        // ---- pass 2: sum of exp and candidate collection
        float sum = 0.f;
        if (has_sc) {
            const float x = LoadVec<T>::one_smem(rb + sc);
            sum += exp_neg_fast(x - m);
            if (x >= tau) {
                const int slot = atomicAdd(&s_ncand[it & 1], 1);
                if (slot < kMaxCand) { cand_v[slot] = x; cand_i[slot] = sc; }
            }
        }
#pragma unroll 2
        for (int vi = tid; vi < nvec; vi += THREADS) {
            float x[V];
            LoadVec<T>::load_smem(buf + 16 + (vi << 4), x);
            float vmax = fmaxf(x[0], x[1]);
#pragma unroll
            for (int j = 2; j < V; j += 2) vmax = fmaxf(vmax, fmaxf(x[j], x[j + 1]));
#pragma unroll
            for (int j = 0; j < V; ++j) sum += exp_neg_fast(x[j] - m);
            if (!(vmax < tau)) {
#pragma unroll
                for (int j = 0; j < V; ++j) {
                    if (x[j] >= tau) {
                        const int slot = atomicAdd(&s_ncand[it & 1], 1);
                        if (slot < kMaxCand) { cand_v[slot] = x[j]; cand_i[slot] = nhead + vi * V + j; }
                    }
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if (lane == 0) red_sum[warp] = sum;
        __syncthreads();
        float sacc = 0.f;
#pragma unroll
        for (int i = 0; i < THREADS / 32; ++i) sacc += red_sum[i];          // fixed order
        const float logs = logf(sacc);
        if (tid == 0) lse_out[row] = m + logs;
        const int ncand = s_ncand[it & 1];
        if (ncand <= kMaxCand) {
            // ---- rank by counting: exact order (value desc, index asc), independent of the collection order
            for (int e = tid; e < ncand; e += THREADS) {
                const float v = cand_v[e]; const int ci = cand_i[e];
                int rank = 0;
                for (int f = 0; f < ncand; ++f) rank += cand_better(cand_v[f], cand_i[f], v, ci) ? 1 : 0;
                if (rank < k) {
                    topk_idx[row * k + rank] = ci;
                    topk_logp[row * k + rank] = (v - m) - logs;         // scipy: (x - max) - log(sum(exp(x - max)))
                }
            }
        } else {
            // ---- exact fallback: k rounds of block arg-best with exclusion of what was already emitted
            __shared__ float bv_s[8];
            __shared__ int bi_s[8];
            float lv = INFINITY; int li = -1;
            for (int r = 0; r < k; ++r) {
                float bv = -INFINITY; int bi = 0x7fffffff;
                for (int c = tid; c < C; c += THREADS) {
                    const float x = LoadVec<T>::one_smem(rb + c);
                    // eligible: strictly after (lv, li) in the (value desc, index asc) order
                    const bool elig = (x < lv) || (x == lv && c > li);
                    if (elig && cand_better(x, c, bv, bi)) { bv = x; bi = c; }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
                    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                    if (cand_better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
                }
                __syncthreads();                                    // previous round's readers are done
                if (lane == 0) { bv_s[warp] = bv; bi_s[warp] = bi; }
                __syncthreads();
                float v = bv_s[0]; int i2 = bi_s[0];
                for (int q = 1; q < THREADS / 32; ++q) if (cand_better(bv_s[q], bi_s[q], v, i2)) { v = bv_s[q]; i2 = bi_s[q]; }
                lv = v; li = i2;
                if (tid == 0) {
                    topk_idx[row * k + r] = i2;
                    topk_logp[row * k + r] = (v - m) - logs;
                }
            }
        }
        // ---- end of row: the edge elements of row i+1 go next to its bulk-copied interior
        if (edge_off >= 0) *reinterpret_cast<T*>(rowbufs + (size_t)(s ^ 1) * buf_bytes + 16 - nxt.head_bytes + edge_off) = edge_val;
        __syncthreads();
        if (nbuf == 1 && next < rows) {                              // single-buffer mode (rows too large for two buffers)
            if (tid == 0) issue(nxt, 0);
            const int off = edge_offset(nxt);
            if (off >= 0) *reinterpret_cast<T*>(rowbufs + 16 - nxt.head_bytes + off) = *reinterpret_cast<const T*>(nxt.a0 + off);
            __syncthreads();
        }
        cur = nxt;
    }
}

// ---------------------------------------------------------------------------------------------- top-k, one warp per row
// Round 2. The CTA-per-row kernel above spends 28 warp-instructions per element (ncu: issue-bound at 62 %, DRAM 39 %): a
// 7375-class row gives each of 256 threads 29 elements, and four block barriers, the shared-memory staging protocol and
// per-row pointer arithmetic weigh more than the element loops. Here ONE WARP owns a row - 230 elements per lane, the
// per-row bookkeeping amortised eight times better, no block barrier - and reads it twice straight from global memory:
//   pass 1  per-lane maximum, 16-byte loads, four independent chains (the only pass that goes to DRAM)
//   bound   the k-th largest of the 32 lane maxima is a lower bound tau for the k-th largest element (k distinct elements
//           reach it); found by k rounds of redux.sync max over order-preserving integer keys; round 1 gives the row max
//   pass 2  the same warp re-reads its row (an L2 hit: it touched it microseconds ago): sum exp(x - max); the few elements
//           >= tau go to a per-warp candidate list
//   rank    candidates ranked by counting (value desc, index asc) -> the top k in order; exact for any input; a row with
//           more than kWarpMaxCand candidates (constant rows) takes an exact k-round arg-max instead.
// 32-40 warps per SM hide the latency; every byte crosses the L2 twice, and the L2 (about 8 TB/s) becomes the bound:
// bf16 rows 0.475 ms on the config-5 tensor = 0.62 of the HBM peak (CTA-per-row: 0.85 ms, 0.35); fp32 rows 1.08 ms, no
// better than the CTA-per-row kernel's 1.00 ms - so fp32 rows stay there (unless they do not fit its shared memory).
// Measured and dropped on the way (fp32 / bf16 on that tensor), each meant to read the row only once:
//   the row staged in a warp-private shared-memory buffer by bulk copy: 7 warps per SM cannot hide their own latencies,
//     1.29 / 0.95 ms;
//   one pass with a per-lane top-4 in registers + online log-sum-exp: a lane's rare insertions are not rare per warp
//     (32 lanes): 26 instructions per element, 1.53 / 1.22 ms;
//   four warps per row with the row in registers (all loads in flight at once): 155 / 113 registers, 12-16 warps per SM,
//     20 instructions per element once the candidate pushes and the ranking are counted: 1.46 / 1.23 ms.
constexpr int kWarpMaxCand = 128;
constexpr int kTopkWarps = 8;

__device__ __forceinline__ int float_order_key(float x) {            // monotonic float -> int (NaN sorts above +inf)
    const int i = __float_as_int(x);
    return i ^ ((i >> 31) & 0x7fffffff);
}

template <typename T>
__global__ void __launch_bounds__(kTopkWarps * 32)
ctc_topk_warp_kernel(const T* __restrict__ logits, long long rows, int Bn, int C, long long stride_t, long long stride_b,
                     int k, int32_t* __restrict__ topk_idx, float* __restrict__ topk_logp, float* __restrict__ lse_out) {
    constexpr int V = LoadVec<T>::N;
    __shared__ float cand_v[kTopkWarps][kWarpMaxCand];
    __shared__ int cand_i[kTopkWarps][kWarpMaxCand];
    __shared__ int ncand_s[kTopkWarps];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long row = (long long)blockIdx.x * kTopkWarps + warp;
    if (lane == 0) ncand_s[warp] = 0;
    __syncwarp();
    if (row >= rows) return;
    const long long t = row / Bn, b = row - t * Bn;
    const T* p = logits + t * stride_t + b * stride_b;
    const uintptr_t addr = reinterpret_cast<uintptr_t>(p);
    int head = (int)(((16 - (addr & 15)) & 15) / sizeof(T));
    if (head > C) head = C;
    const int nvec = (C - head) / V;
    const int tail0 = head + nvec * V;
    const T* pv = p + head;
    const int kk = k < 32 ? k : 32;
    // the scalar elements in front of / behind the aligned interior (< V each)
    float x_h = -INFINITY, x_t = -INFINITY;
    if (lane < head) x_h = LoadVec<T>::one(p + lane);
    if (tail0 + lane < C) x_t = LoadVec<T>::one(p + tail0 + lane);

    // ---- pass 1: per-lane maximum
    float lm = fmaxf(x_h, x_t);
    if (sizeof(T) == 2) {
        // bf16 rows stay packed: 4 HMNMX2 per 8 elements, four independent chains (the kernel is issue-bound: 12 instructions
        // per element with the rows unpacked to fp32 in both passes)
        const __nv_bfloat162 ninf = __float2bfloat162_rn(-INFINITY);
        __nv_bfloat162 a[4] = {ninf, ninf, ninf, ninf};
        int vi = lane;
        for (; vi + 96 < nvec; vi += 128) {
            uint4 q[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) q[u] = ld_nc_v4(pv + (long long)(vi + 32 * u) * V);
#pragma unroll
            for (int u = 0; u < 4; ++u)
                a[u] = __hmax2(a[u], __hmax2(__hmax2(*reinterpret_cast<const __nv_bfloat162*>(&q[u].x), *reinterpret_cast<const __nv_bfloat162*>(&q[u].y)),
                                             __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&q[u].z), *reinterpret_cast<const __nv_bfloat162*>(&q[u].w))));
        }
        for (; vi < nvec; vi += 32) {
            const uint4 q = ld_nc_v4(pv + (long long)vi * V);
            a[0] = __hmax2(a[0], __hmax2(__hmax2(*reinterpret_cast<const __nv_bfloat162*>(&q.x), *reinterpret_cast<const __nv_bfloat162*>(&q.y)),
                                         __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&q.z), *reinterpret_cast<const __nv_bfloat162*>(&q.w))));
        }
        const __nv_bfloat162 t2 = __hmax2(__hmax2(a[0], a[1]), __hmax2(a[2], a[3]));
        lm = fmaxf(lm, fmaxf(__low2float(t2), __high2float(t2)));
    } else {
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        int vi = lane;
        for (; vi + 96 < nvec; vi += 128) {
            float x[4][V];
#pragma unroll
            for (int u = 0; u < 4; ++u) LoadVec<T>::load(pv + (long long)(vi + 32 * u) * V, x[u]);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
#pragma unroll
                for (int j = 0; j < V; j += 2) m4[u] = fmaxf(m4[u], fmaxf(x[u][j], x[u][j + 1]));
            }
        }
        for (; vi < nvec; vi += 32) {
            float x[V];
            LoadVec<T>::load(pv + (long long)vi * V, x);
#pragma unroll
            for (int j = 0; j < V; j += 2) m4[0] = fmaxf(m4[0], fmaxf(x[j], x[j + 1]));
        }
        lm = fmaxf(lm, fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])));
    }
    // ---- tau = kk-th largest lane maximum, m = the largest
    float m, tau;
    {
        int key = float_order_key(lm);
        const int dead = (int)0x80000000;
        int mx = __reduce_max_sync(0xffffffffu, key);
        m = __shfl_sync(0xffffffffu, lm, __ffs(__ballot_sync(0xffffffffu, key == mx)) - 1);
        tau = m;
        for (int r = 1; r < kk; ++r) {
            const unsigned hit = __ballot_sync(0xffffffffu, key == mx);
            const int first = __ffs(hit) - 1;
            if (lane == first) key = dead;                              // one lane leaves per round
            mx = __reduce_max_sync(0xffffffffu, key);
        }
        if (kk > 1) tau = __shfl_sync(0xffffffffu, lm, __ffs(__ballot_sync(0xffffffffu, key == mx)) - 1);
    }
    // ---- pass 2: sum of exp and candidate collection
    auto push = [&](float x, int idx) {
        const int slot = atomicAdd(&ncand_s[warp], 1);
        if (slot < kWarpMaxCand) { cand_v[warp][slot] = x; cand_i[warp][slot] = idx; }
    };
    float sum = 0.f;
    const float nml = -m * kLog2e;                      // exp(x - m) = 2^(x*log2e - m*log2e): one FFMA + one MUFU per element
    if (lane < head) { sum += exp_neg_fast(x_h - m); if (x_h >= tau) push(x_h, lane); }
    if (tail0 + lane < C) { sum += exp_neg_fast(x_t - m); if (x_t >= tau) push(x_t, tail0 + lane); }
    {
        float s4[4] = {0.f, 0.f, 0.f, 0.f};
        int vi = lane;
        for (; vi + 96 < nvec; vi += 128) {
            float x[4][V];
#pragma unroll
            for (int u = 0; u < 4; ++u) LoadVec<T>::load(pv + (long long)(vi + 32 * u) * V, x[u]);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                float vmax = fmaxf(x[u][0], x[u][1]);
                float acc = ex2_fast(fmaf(x[u][0], kLog2e, nml)) + ex2_fast(fmaf(x[u][1], kLog2e, nml));
#pragma unroll
                for (int j = 2; j < V; j += 2) {
                    vmax = fmaxf(vmax, fmaxf(x[u][j], x[u][j + 1]));
                    acc += ex2_fast(fmaf(x[u][j], kLog2e, nml)) + ex2_fast(fmaf(x[u][j + 1], kLog2e, nml));
                }
                s4[u] += acc;
                if (!(vmax < tau)) {
#pragma unroll
                    for (int j = 0; j < V; ++j)
                        if (x[u][j] >= tau) push(x[u][j], head + (vi + 32 * u) * V + j);
                }
            }
        }
        for (; vi < nvec; vi += 32) {
            float x[V];
            LoadVec<T>::load(pv + (long long)vi * V, x);
#pragma unroll
            for (int j = 0; j < V; ++j) {
                s4[0] += exp_neg_fast(x[j] - m);
                if (x[j] >= tau) push(x[j], head + vi * V + j);
            }
        }
        sum += (s4[0] + s4[1]) + (s4[2] + s4[3]);
    }
    sum = warp_sum(sum);
    const float logs = logf(sum);
    if (lane == 0) lse_out[row] = m + logs;
    __syncwarp();                                                       // candidate list complete
    const int ncand = ncand_s[warp];
    int32_t* oi = topk_idx + row * k;
    float* op = topk_logp + row * k;
    if (ncand <= kWarpMaxCand) {
        // ---- rank by counting: exact order (value desc, index asc), independent of the collection order
        for (int e = lane; e < ncand; e += 32) {
            const float v = cand_v[warp][e]; const int ci = cand_i[warp][e];
            int rank = 0;
            for (int f = 0; f < ncand; ++f) rank += cand_better(cand_v[warp][f], cand_i[warp][f], v, ci) ? 1 : 0;
            if (rank < k) {
                oi[rank] = ci;
                op[rank] = (v - m) - logs;                               // scipy: (x - max) - log(sum(exp(x - max)))
            }
        }
    } else {
        // ---- exact fallback: k rounds of warp arg-best with exclusion of what was already emitted
        float lv = INFINITY; int li = -1;
        for (int r = 0; r < k; ++r) {
            float bv = -INFINITY; int bi = 0x7fffffff;
            for (int c = lane; c < C; c += 32) {
                const float x = LoadVec<T>::one(p + c);
                const bool elig = (x < lv) || (x == lv && c > li);
                if (elig && cand_better(x, c, bv, bi)) { bv = x; bi = c; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
                const int ov_i = __shfl_xor_sync(0xffffffffu, bi, o);
                if (cand_better(ov, ov_i, bv, bi)) { bv = ov; bi = ov_i; }
            }
            lv = bv; li = bi;
            if (lane == 0) { oi[r] = bi; op[r] = (bv - m) - logs; }
        }
    }
}

// ---------------------------------------------------------------------------------------------- top-k, one warp per row, one read
// Round 2, after the measurement in scripts/topk_ramp.py: the two-pass warp kernel above keeps 4736 rows in flight (32 warps
// on each of 148 SMs) - 70 MB of bf16 rows - and its second pass hits the L2 only on a quiet device: with ANY other touched
// allocation alive (0.25 GB is enough) and always under ncu the re-read misses (L2 hit rate 9 %, DRAM traffic 2x the tensor)
// and the kernel falls from 0.43 to 0.57 ms. This kernel reads every byte once and needs no L2 residency at all: the warp
// walks its row in CHUNKS of kChunkVec 16-byte vectors per lane held in registers, and does both passes on a chunk while it
// is there:
//   max     per-vector and per-lane maxima of the chunk (bf16 rows packed, HMNMX2); the lane's RUNNING maximum over all
//           chunks so far; one redux.sync gives the row maximum so far, m
//   rescale the running sum of exponentials is kept relative to m; when a chunk raises m the four partial sums are
//           multiplied by 2^((m_old - m_new) log2 e) - one warp-uniform branch per chunk (online log-sum-exp at chunk, not
//           element, granularity: the per-element work is the FFMA + MUFU + FADD of the two-pass kernel)
//   bound   tau = the k-th largest RUNNING lane maximum (k distinct elements of the row reach it, so it never exceeds the
//           k-th largest element; it only grows). Refreshed after chunks 0, 1, 3, 7, ... and once more at the end
//   mark    a lane whose chunk maximum reaches tau notes WHICH of its vectors do (vector index -> the warp's list, one shared-
//           memory atomic per lane and chunk). Nothing per element: the first version pushed the elements themselves from inside the loop, and since a warp step
//           covers 256 elements, 4 out of 5 steps had some lane on that 170-instruction path - 20 issue slots per element,
//           0.56 ms for bf16 rows. Every element of the final top k is >= the final tau >= the tau of its own chunk, so
//           its vector is in the list
//   collect after the row: the elements >= the FINAL tau of the ~50 listed vectors become the candidates (~13). The list keeps
//           the vectors' 16 bytes next to their indices: re-reading them from global memory cost fp32 rows 7 % more DRAM
//           traffic (a 32-byte sector per entry, and 140 MB of rows in flight do not stay in the L2)
//   rank    by counting, exactly as in the kernels above.
// An exact k-round arg-max over the row (re-read from global memory) covers lists that overflow (ascending or constant rows).
constexpr int kChunkVec = 8;
constexpr int kChunkMaxList = 192;     // vectors that may hold a candidate, per row (index + 16 bytes of data each)
constexpr int kChunkMaxCand = 128;     // elements >= the final tau, per row

__device__ __forceinline__ float order_key_to_float(int k) { return __int_as_float(k ^ ((k >> 31) & 0x7fffffff)); }

template <typename T>
__global__ void __launch_bounds__(kTopkWarps * 32, 4)
ctc_topk_chunk_kernel(const T* __restrict__ logits, long long rows, int Bn, int C, long long stride_t, long long stride_b,
                      int k, int32_t* __restrict__ topk_idx, float* __restrict__ topk_logp, float* __restrict__ lse_out) {
    constexpr int V = LoadVec<T>::N;
    constexpr unsigned kFull = 0xffffffffu;
    __shared__ uint4 list_q[kTopkWarps][kChunkMaxList];
    __shared__ int list_v[kTopkWarps][kChunkMaxList];
    __shared__ float cand_v[kTopkWarps][kChunkMaxCand];
    __shared__ int cand_i[kTopkWarps][kChunkMaxCand];
    __shared__ int nlist_s[kTopkWarps], ncand_s[kTopkWarps];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long row = (long long)blockIdx.x * kTopkWarps + warp;
    if (lane == 0) { nlist_s[warp] = 0; ncand_s[warp] = 0; }
    __syncwarp();
    if (row >= rows) return;
    const unsigned t = (unsigned)row / (unsigned)Bn, b = (unsigned)row - t * (unsigned)Bn;     // rows < 2^31 (checked by the caller)
    const T* p = logits + (long long)t * stride_t + (long long)b * stride_b;
    const uintptr_t addr = reinterpret_cast<uintptr_t>(p);
    int head = (int)(((16 - (addr & 15)) & 15) / sizeof(T));
    if (head > C) head = C;
    const int nvec = (C - head) / V;
    const int tail0 = head + nvec * V;
    const T* pv = p + head;
    const int kk = k < 32 ? k : 32;
    float x_h = -INFINITY, x_t = -INFINITY;
    if (lane < head) x_h = LoadVec<T>::one(p + lane);
    if (tail0 + lane < C) x_t = LoadVec<T>::one(p + tail0 + lane);

    auto push = [&](float x, int idx) {
        const int slot = atomicAdd(&ncand_s[warp], 1);
        if (slot < kChunkMaxCand) { cand_v[warp][slot] = x; cand_i[warp][slot] = idx; }
    };
    // tau <- max(tau, kk-th largest running lane maximum). The keys are made distinct (low five bits <- 31 - lane, the result
    // floored to a multiple of 32 in key order: at most 32 fp32 ulps low, nothing for bf16 rows), so a round is redux + compare
    auto refresh_tau = [&](float lmax, float tau) {
        int key = (float_order_key(lmax) & ~31) | (31 - lane);
        const int dead = (int)0x80000000;
        int mx = __reduce_max_sync(kFull, key);
        for (int r = 1; r < kk; ++r) {
            if (key == mx) key = dead;                                      // one lane leaves per round
            mx = __reduce_max_sync(kFull, key);
        }
        return fmaxf(tau, order_key_to_float(mx & ~31));
    };
    auto unpack = [](const uint4& q, float (&x)[V]) {
        if (sizeof(T) == 2) {
            x[0] = bf16_lo(q.x); x[1] = bf16_hi(q.x); x[2] = bf16_lo(q.y); x[3] = bf16_hi(q.y);
            x[V - 4] = bf16_lo(q.z); x[V - 3] = bf16_hi(q.z); x[V - 2] = bf16_lo(q.w); x[V - 1] = bf16_hi(q.w);
        } else {
            x[0] = __uint_as_float(q.x); x[1] = __uint_as_float(q.y); x[2] = __uint_as_float(q.z); x[3] = __uint_as_float(q.w);
        }
    };

    const uint32_t ninf_w = sizeof(T) == 2 ? 0xFF80FF80u : 0xFF800000u;
    float lmax = fmaxf(x_h, x_t);                                           // running lane maximum (the lane's own elements)
    float m = -INFINITY, tau = -INFINITY;
    float s4[4] = {0.f, 0.f, 0.f, 0.f};
    int chunk = 0;
    for (int base = 0; base < nvec; base += 32 * kChunkVec, ++chunk) {
        uint4 q[kChunkVec];
        if (base + 32 * kChunkVec <= nvec) {
#pragma unroll
            for (int u = 0; u < kChunkVec; ++u) q[u] = ld_nc_v4(pv + (long long)(base + 32 * u + lane) * V);
        } else {
#pragma unroll
            for (int u = 0; u < kChunkVec; ++u) {
                const int vi = base + 32 * u + lane;
                q[u] = make_uint4(ninf_w, ninf_w, ninf_w, ninf_w);
                if (vi < nvec) q[u] = ld_nc_v4(pv + (long long)vi * V);
            }
        }
        // ---- maxima: per vector (kept for the marking), per lane
        uint32_t vmx[kChunkVec];
        float cm;
        if (sizeof(T) == 2) {
            __nv_bfloat162 a[kChunkVec];
#pragma unroll
            for (int u = 0; u < kChunkVec; ++u) {
                a[u] = __hmax2(__hmax2(*reinterpret_cast<const __nv_bfloat162*>(&q[u].x), *reinterpret_cast<const __nv_bfloat162*>(&q[u].y)),
                               __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&q[u].z), *reinterpret_cast<const __nv_bfloat162*>(&q[u].w)));
                vmx[u] = *reinterpret_cast<const uint32_t*>(&a[u]);
            }
#pragma unroll
            for (int w = kChunkVec / 2; w > 0; w >>= 1) {
#pragma unroll
                for (int u = 0; u < w; ++u) a[u] = __hmax2(a[u], a[u + w]);
            }
            cm = fmaxf(__low2float(a[0]), __high2float(a[0]));
        } else {
            float a[kChunkVec];
#pragma unroll
            for (int u = 0; u < kChunkVec; ++u) {
                a[u] = fmaxf(fmaxf(__uint_as_float(q[u].x), __uint_as_float(q[u].y)), fmaxf(__uint_as_float(q[u].z), __uint_as_float(q[u].w)));
                vmx[u] = __float_as_uint(a[u]);
            }
#pragma unroll
            for (int w = kChunkVec / 2; w > 0; w >>= 1) {
#pragma unroll
                for (int u = 0; u < w; ++u) a[u] = fmaxf(a[u], a[u + w]);
            }
            cm = a[0];
        }
        lmax = fmaxf(lmax, cm);
        // ---- row maximum so far; tau after chunks 0, 1, 3, 7, ...
        const int mx = __reduce_max_sync(kFull, float_order_key(lmax));
        const float m_new = order_key_to_float(mx);
        if ((chunk & (chunk + 1)) == 0) tau = refresh_tau(lmax, tau);
        if (m_new > m) {                                                    // warp-uniform
            const float sc = ex2_fast((m - m_new) * kLog2e);                // m = -inf: 0
            s4[0] *= sc; s4[1] *= sc; s4[2] *= sc; s4[3] *= sc;
            m = m_new;
        }
        // ---- mark the vectors that may hold one of the top k: only the (about three) lanes whose chunk maximum reaches tau
        if (!(cm < tau)) {
            const __nv_bfloat162 tau2 = __float2bfloat162_rn(tau);          // exact: tau is an element of a bf16 row (or -inf)
            const int nvalid = nvec - base - lane;                          // vector u of this lane exists iff 32 u < nvalid
            unsigned msk = 0;
#pragma unroll
            for (int u = 0; u < kChunkVec; ++u) {
                bool hit;
                if (sizeof(T) == 2) hit = !__hblt2(*reinterpret_cast<const __nv_bfloat162*>(&vmx[u]), tau2);   // some half >= tau (or NaN)
                else hit = !(__uint_as_float(vmx[u]) < tau);
                if (hit && 32 * u < nvalid) msk |= 1u << u;
            }
            int slot = atomicAdd(&nlist_s[warp], __popc(msk));
#pragma unroll
            for (int u = 0; u < kChunkVec; ++u) {
                if (msk >> u & 1u) {
                    if (slot < kChunkMaxList) { list_v[warp][slot] = base + 32 * u + lane; list_q[warp][slot] = q[u]; }
                    ++slot;
                }
            }
        }
        // ---- sum of exponentials, from the registers
        const float nml = m > -INFINITY ? -m * kLog2e : 0.f;                // all -inf so far: every term is 2^-inf = 0
        const int rem = nvec - base;                                        // warp-uniform: vector step u has data iff 32 u < rem
#pragma unroll
        for (int u = 0; u < kChunkVec; ++u) {
            if (32 * u >= rem) break;
            float x[V];
            unpack(q[u], x);
            float acc = ex2_fast(fmaf(x[0], kLog2e, nml)) + ex2_fast(fmaf(x[1], kLog2e, nml));
#pragma unroll
            for (int j = 2; j < V; j += 2) acc += ex2_fast(fmaf(x[j], kLog2e, nml)) + ex2_fast(fmaf(x[j + 1], kLog2e, nml));
            s4[u & 3] += acc;
        }
    }
    // ---- final maximum and bound, unless the last chunk has just refreshed them (also covers rows with no aligned interior)
    if (chunk == 0 || ((chunk - 1) & chunk) != 0) {
        const int mx = __reduce_max_sync(kFull, float_order_key(lmax));
        const float m_new = order_key_to_float(mx);
        tau = refresh_tau(lmax, tau);
        if (m_new > m) {
            const float sc = ex2_fast((m - m_new) * kLog2e);
            s4[0] *= sc; s4[1] *= sc; s4[2] *= sc; s4[3] *= sc;
            m = m_new;
        }
    }
    // ---- the scalars in front of / behind the interior; candidates out of the marked vectors
    float sum = (s4[0] + s4[1]) + (s4[2] + s4[3]);
    if (lane < head) { sum += exp_neg_fast(x_h - m); if (x_h >= tau) push(x_h, lane); }
    if (tail0 + lane < C) { sum += exp_neg_fast(x_t - m); if (x_t >= tau) push(x_t, tail0 + lane); }
    sum = warp_sum(sum);                                                    // (its shuffles also order the list writes)
    const float logs = logf(sum);
    if (lane == 0) lse_out[row] = m + logs;
    __syncwarp();
    const int nlist = nlist_s[warp];
    if (nlist <= kChunkMaxList) {
        for (int e = lane; e < nlist; e += 32) {
            const int vi = list_v[warp][e];
            float x[V];
            unpack(list_q[warp][e], x);
#pragma unroll
            for (int j = 0; j < V; ++j)
                if (x[j] >= tau) push(x[j], head + vi * V + j);
        }
    }
    __syncwarp();                                                           // candidate list complete
    const int ncand = ncand_s[warp];
    int32_t* oi = topk_idx + row * k;
    float* op = topk_logp + row * k;
    if (nlist <= kChunkMaxList && ncand <= kChunkMaxCand) {
        // ---- rank by counting: exact order (value desc, index asc), independent of the collection order
        for (int e = lane; e < ncand; e += 32) {
            const float v = cand_v[warp][e]; const int ci = cand_i[warp][e];
            int rank = 0;
            for (int f = 0; f < ncand; ++f) rank += cand_better(cand_v[warp][f], cand_i[warp][f], v, ci) ? 1 : 0;
            if (rank < k) {
                oi[rank] = ci;
                op[rank] = (v - m) - logs;                                   // scipy: (x - max) - log(sum(exp(x - max)))
            }
        }
    } else {
        // ---- exact fallback: k rounds of warp arg-best with exclusion of what was already emitted
        float lv = INFINITY; int li = -1;
        for (int r = 0; r < k; ++r) {
            float bv = -INFINITY; int bi = 0x7fffffff;
            for (int c = lane; c < C; c += 32) {
                const float x = LoadVec<T>::one(p + c);
                const bool elig = (x < lv) || (x == lv && c > li);
                if (elig && cand_better(x, c, bv, bi)) { bv = x; bi = c; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(kFull, bv, o);
                const int ov_i = __shfl_xor_sync(kFull, bi, o);
                if (cand_better(ov, ov_i, bv, bi)) { bv = ov; bi = ov_i; }
            }
            lv = bv; li = bi;
            if (lane == 0) { oi[r] = bi; op[r] = (bv - m) - logs; }
        }
    }
}

// ---------------------------------------------------------------------------------------------- beam search
__device__ __forceinline__ double logaddexp_np(double x, double y) {
    // numpy npy_logaddexp for doubles:  x == y -> x + ln 2;  d = x - y;  d > 0 -> x + log1p(exp(-d));  d <= 0 -> y + log1p(exp(d));
    // NaN -> d.  Written with selects so that all lanes of a warp run ONE exp and ONE log1p (the branchy form serialised
    // the lanes that took different sides); the operations and their operands are the same, so are the bits.
    const double d = __dsub_rn(x, y);
    const bool pos = d > 0;
    const double hi = pos ? x : y;
    const double arg = pos ? -d : d;
    const double r = __dadd_rn(hi, log1p(exp(arg)));
    if (x == y) return __dadd_rn(x, 0.693147180559945309417232121458176568);
    if (d != d) return d;
    return r;
}

constexpr int kBeamThreads = 128;
constexpr int kNoKey = 0x7fffffff;

struct KeptState {
    int node[kMaxBeam];                    // trie node of the prefix
    int par[kMaxBeam];                     // trie node of the prefix minus its last character
    int len[kMaxBeam];
    int last[kMaxBeam];
    unsigned long long hash[kMaxBeam];
    unsigned long long phash[kMaxBeam];    // hash of the prefix minus its last character
    double pb[kMaxBeam], pnb[kMaxBeam], P[kMaxBeam], lmsum[kMaxBeam];     // P = Beam.prob() = logaddexp(pb, pnb)
};

__device__ __forceinline__ unsigned long long mix_hash(unsigned long long h, int c) {
    h ^= (unsigned long long)(c + 1) * 0x9E3779B97F4A7C15ull;
    h *= 0xFF51AFD7ED558CCDull;
    h ^= h >> 29;
    return h;
}

// exact string equality of two trie nodes known to have equal length
__device__ bool same_string(const int* parent, const int* chr, int a, int b) {
    while (a != b) {
        if (chr[a] != chr[b]) return false;
        a = parent[a]; b = parent[b];
    }
    return true;
}

// One CTA of 128 threads per sequence. A time step of the reference (:212-285) is five block-wide phases, none of them
// serial in the number of beams x candidates:
//   A  per kept beam: which kept beam is its prefix minus the last character (hash + exact check); prefetch of the
//      next step's candidates
//   B  per (beam j, candidate q) pair: does prefix_j + c already exist among the kept beams? -> "new entry" bit mask;
//      per kept beam: the position in the reference's double loop at which its dict entry is created (its own first
//      valid candidate, or the earlier pair that extends its parent to it)
//   C  dict insertion order = rank of the creation positions: popcount of the new-entry mask + count of earlier kept keys
//   D  per entry: <= 2 contributions per accumulator, two convergent float64 logaddexp calls, LM score
//   E  stable descending rank by counting, the best beam_size become the next kept beams
__global__ void __launch_bounds__(kBeamThreads)
ctc_prefix_beam_kernel(const int32_t* __restrict__ topk_idx, const float* __restrict__ topk_logp, int Tn, int Bn, int C,
                       int k, int beam_size, double lm_penalty, double len_bonus, const double* __restrict__ lm_table,
                       const hctr_ngram_lm ng_lm, int32_t* __restrict__ out_idx, int32_t* __restrict__ out_len,
                       int32_t* __restrict__ status, unsigned char* __restrict__ workspace, long long ws_per_seq) {
    __shared__ KeptState kept[2];
    __shared__ int parentk[kMaxBeam];
    __shared__ int key[kMaxBeam];                   // creation position of a kept beam's entry (kNoKey: not created)
    __shared__ int cand[2][kMaxK];
    __shared__ double candp[2][kMaxK];
    __shared__ unsigned newmask[kMaxBeam * kMaxK / 32];
    __shared__ int e_kind[kMaxGen];                 // >=0: kept index j''; <0: NEW
    __shared__ int e_src[kMaxGen], e_chr[kMaxGen];
    __shared__ double e_pb[kMaxGen], e_pnb[kMaxGen], e_P[kMaxGen], e_tot[kMaxGen], e_lm[kMaxGen];
    __shared__ int s_ngen, s_ng, s_qblank;

    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int unknown = C - 1;
    const bool use_ngram = ng_lm.entries != nullptr;
    unsigned char* ws = workspace + (long long)b * ws_per_seq;
    int* g_char = reinterpret_cast<int*>(ws);
    int* g_time = g_char + Tn;
    const int cap = Tn * beam_size + 1;
    int* n_parent = g_time + Tn;
    int* n_chr = n_parent + cap;

    // ---- top_line: greedy (char, t) list from the top-1 candidates (:188-195), compaction by warp 0
    if (warp == 0) {
        int ng = 0;
        for (int base = 0; base < Tn; base += 32) {
            const int t = base + lane;
            int cur = 0, keep = 0;
            if (t < Tn) {
                cur = topk_idx[((long long)t * Bn + b) * k];
                const int prev = t > 0 ? topk_idx[((long long)(t - 1) * Bn + b) * k] : -1;
                keep = (cur != 0) && (cur != unknown) && !(t > 0 && prev == cur);
            }
            const unsigned bal = __ballot_sync(0xffffffffu, keep);
            if (keep) {
                const int pos = ng + __popc(bal & ((1u << lane) - 1));
                g_char[pos] = cur; g_time[pos] = t;
            }
            ng += __popc(bal);
        }
        if (lane == 0) {
            s_ng = ng;
            n_parent[0] = 0; n_chr[0] = -1;
            kept[0].node[0] = 0; kept[0].par[0] = 0; kept[0].len[0] = 0; kept[0].last[0] = -1;
            kept[0].hash[0] = 0x243F6A8885A308D3ull; kept[0].phash[0] = 0;
            kept[0].pb[0] = 0.0; kept[0].pnb[0] = -INFINITY; kept[0].P[0] = 0.0; kept[0].lmsum[0] = 0.0;   // Beam() :289-297
        }
    }
    if (tid >= 32 && tid < 32 + k) {                 // candidates of step 0
        const long long o = (long long)b * k + (tid - 32);
        cand[0][tid - 32] = topk_idx[o];
        candp[0][tid - 32] = (double)topk_logp[o];
    }
    __syncthreads();
    const int ng = s_ng;
    if (ng == 0) {                                   // reference: top_line[-1] -> IndexError (:198)
        if (tid == 0) { status[b] = HCTR_ERR_INDEX; out_len[b] = 0; }
        return;
    }
    int end_step = g_time[ng - 1] + 4;               // :198-199
    if (end_step >= Tn) end_step = Tn;

    int nkept = 1, cur_buf = 0, gptr = 0;
    int next_time = g_time[0];

    for (int t = 0; t < end_step; ++t) {
        KeptState& K = kept[cur_buf];
        KeptState& Kn = kept[cur_buf ^ 1];
        const int* cd = cand[t & 1];
        const double* cp = candp[t & 1];               // fp32 log-probs, promoted on addition
        while (gptr < ng && next_time <= t) {          // suffix = next <=4 greedy chars after t (:202-203)
            ++gptr;
            next_time = gptr < ng ? g_time[gptr] : kNoKey;
        }
        int nsuf = ng - gptr; if (nsuf > 4) nsuf = 4;

        // ---- A
        if (tid < nkept) {
            // which kept beam (if any) is this beam's prefix minus its last character?
            int pk = -1;
            if (K.len[tid] > 0) {
                const int par = K.par[tid];
                const unsigned long long ph = K.phash[tid];
                for (int j = 0; j < nkept; ++j) {
                    if (K.len[j] != K.len[tid] - 1) continue;
                    if (K.node[j] == par || (K.hash[j] == ph && same_string(n_parent, n_chr, K.node[j], par))) {
                        pk = j; break;
                    }
                }
            }
            parentk[tid] = pk;
        } else if (tid >= 32 && tid < 32 + k) {
            if (t + 1 < end_step) {
                const long long o = ((long long)(t + 1) * Bn + b) * k + (tid - 32);
                cand[(t + 1) & 1][tid - 32] = topk_idx[o];
                candp[(t + 1) & 1][tid - 32] = (double)topk_logp[o];
            }
        } else if (tid == 64) {
            int qb = -1;
            for (int q = 0; q < k; ++q) if (cd[q] == 0) qb = q;
            s_qblank = qb;
        }
        __syncthreads();

        // ---- B
        const int npairs = nkept * k;
        for (int base = 0; base < npairs; base += kBeamThreads) {
            const int p = base + tid;
            bool isnew = false;
            if (p < npairs) {
                const int j = p / k, q = p - j * k;
                const int idx = cd[q];
                if (idx < unknown && idx != 0) {                                  // :238-239; blank extends nothing
                    isnew = true;                                                 // unless prefix_j + idx is a kept prefix
                    for (int j2 = 0; j2 < nkept; ++j2)
                        if (parentk[j2] == j && K.last[j2] == idx) { isnew = false; break; }
                }
            }
            const unsigned bal = __ballot_sync(0xffffffffu, isnew);
            if (lane == 0) newmask[(base >> 5) + warp] = bal;
        }
        if (tid < nkept) {
            // the dict entry of kept beam `tid` is created by its own first valid candidate (:243-244) or, earlier, by the
            // pair (parent beam, candidate == its last character) that extends the parent to it (:250-252)
            int own = kNoKey, ref = kNoKey;
            for (int q = k - 1; q >= 0; --q) if (cd[q] < unknown) own = 2 * (tid * k + q);
            const int pj = parentk[tid];
            if (pj >= 0)
                for (int q = 0; q < k; ++q) if (cd[q] == K.last[tid]) ref = 2 * (pj * k + q) + 1;
            key[tid] = own < ref ? own : ref;
        }
        __syncthreads();

        // ---- C: entry index = number of entries created at earlier positions
        const int nwords = (npairs + 31) >> 5;
        for (int base = 0; base < npairs; base += kBeamThreads) {
            const int p = base + tid;
            if (p < npairs && ((newmask[p >> 5] >> (p & 31)) & 1u)) {
                int e = __popc(newmask[p >> 5] & ((1u << (p & 31)) - 1u));
                for (int w = 0; w < (p >> 5); ++w) e += __popc(newmask[w]);
                for (int j2 = 0; j2 < nkept; ++j2) e += (key[j2] < 2 * p + 1) ? 1 : 0;
                const int j = p / k;
                e_kind[e] = -1; e_src[e] = j; e_chr[e] = cd[p - j * k];            // :253-255
            }
        }
        if (tid < nkept && key[tid] != kNoKey) {
            const int lim = key[tid] >> 1;                       // new entries of pairs p < lim come first
            int e = 0;
            for (int w = 0; w < nwords; ++w) {
                const int lo = w << 5;
                unsigned m = newmask[w];
                if (lim < lo + 32) m = (lim <= lo) ? 0u : (m & ((1u << (lim - lo)) - 1u));
                e += __popc(m);
            }
            for (int j2 = 0; j2 < nkept; ++j2) e += (key[j2] < key[tid]) ? 1 : 0;
            e_kind[e] = tid;
        }
        if (tid == kBeamThreads - 1) {
            int n = 0;
            for (int w = 0; w < nwords; ++w) n += __popc(newmask[w]);
            for (int j2 = 0; j2 < nkept; ++j2) n += (key[j2] != kNoKey) ? 1 : 0;
            s_ngen = n;
        }
        __syncthreads();
        const int ngen = s_ngen;
        const int qblank = s_qblank;

        // ---- D: scores. Every accumulator receives at most two contributions (the candidates of a step are distinct
        //      classes); np.logaddexp is symmetric and logaddexp(-inf, a) == a + 0, so one unconditional call per
        //      accumulator is bit-identical to the reference's accumulation in loop order.
        for (int e = tid; e < ngen; e += kBeamThreads) {
            double pb = -INFINITY, c1 = -INFINITY, c2 = -INFINITY, lm, plen;
            const int kind = e_kind[e];
            if (kind >= 0) {
                const int j2 = kind;
                if (qblank >= 0) pb = __dadd_rn(__dadd_rn(K.P[j2], cp[qblank]), 0.0);                    // :246-249
                int qrep = -1;
                for (int q = 0; q < k; ++q) if (cd[q] == K.last[j2]) qrep = q;
                if (qrep >= 0) {
                    c1 = __dadd_rn(K.pnb[j2], cp[qrep]);                                                 // :264-265
                    const int j = parentk[j2];
                    if (j >= 0) c2 = (cd[qrep] != K.last[j]) ? __dadd_rn(K.P[j], cp[qrep])               // :256-258
                                                             : __dadd_rn(K.pb[j], cp[qrep]);             // :261-262
                }
                lm = K.lmsum[j2]; plen = (double)K.len[j2];
            } else {
                const int j = e_src[e], idx = e_chr[e];
                double p = 0.0;
                for (int q = 0; q < k; ++q) if (cd[q] == idx) p = cp[q];
                c1 = (idx != K.last[j]) ? __dadd_rn(K.P[j], p) : __dadd_rn(K.pb[j], p);
                if (use_ngram) {
                    // kenlm: float32 running total; p(idx | last order-1 characters of prefix_j, <s> in front)
                    int ctx[kNgramMaxOrder - 1]; int m;
                    trie_context(ng_lm, n_parent, n_chr, K.node[j], -1, ctx, m);
                    lm = (double)__fadd_rn((float)K.lmsum[j], ngram_word_score(ng_lm, ctx, m, __ldg(ng_lm.vocab + idx)));
                } else {
                    lm = lm_table ? __dadd_rn(K.lmsum[j], lm_table[idx]) : 0.0;
                }
                plen = (double)(K.len[j] + 1);
            }
            const double pnb = logaddexp_np(c1, c2);
            e_lm[e] = lm;                                                       // LM sum over the prefix only
            double lmt = lm;
            if (lm_table) for (int c = 0; c < nsuf; ++c) lmt = __dadd_rn(lmt, lm_table[g_char[gptr + c]]);
            if (use_ngram && nsuf > 0) {
                int ctx[kNgramMaxOrder - 1]; int m;
                if (kind >= 0) trie_context(ng_lm, n_parent, n_chr, K.node[kind], -1, ctx, m);
                else trie_context(ng_lm, n_parent, n_chr, K.node[e_src[e]], e_chr[e], ctx, m);
                float tot = (float)lm;
                for (int c = 0; c < nsuf; ++c) {
                    const int w = __ldg(ng_lm.vocab + g_char[gptr + c]);
                    tot = __fadd_rn(tot, ngram_word_score(ng_lm, ctx, m, w));
                    ngram_push(ctx, m, ng_lm.order - 1, w);
                }
                lmt = (double)tot;
            }
            const double pt = __dadd_rn(__dmul_rn(lmt, lm_penalty), __dmul_rn(plen, len_bonus));   // :277-281
            const double P = logaddexp_np(pb, pnb);                              // Beam.prob() :299-300
            e_pb[e] = pb; e_pnb[e] = pnb; e_P[e] = P;
            e_tot[e] = __dadd_rn(P, pt);                                         // Beam.total() :302-303
        }
        __syncthreads();

        // ---- E: sorted(..., key=total, reverse=True)[:beam_size]: stable, ties keep insertion order (:283-285)
        const int keep_n = ngen < beam_size ? ngen : beam_size;
        for (int e = tid; e < ngen; e += kBeamThreads) {
            const double te = e_tot[e];
            int rank = 0;
            for (int f = 0; f < ngen; ++f) rank += (e_tot[f] > te) || (e_tot[f] == te && f < e);
            if (rank < keep_n) {
                Kn.pb[rank] = e_pb[e]; Kn.pnb[rank] = e_pnb[e]; Kn.P[rank] = e_P[e]; Kn.lmsum[rank] = e_lm[e];
                if (e_kind[e] >= 0) {
                    const int j2 = e_kind[e];
                    Kn.node[rank] = K.node[j2]; Kn.par[rank] = K.par[j2]; Kn.len[rank] = K.len[j2]; Kn.last[rank] = K.last[j2];
                    Kn.hash[rank] = K.hash[j2]; Kn.phash[rank] = K.phash[j2];
                } else {
                    const int j = e_src[e], idx = e_chr[e];
                    const int id = 1 + t * beam_size + rank;
                    n_parent[id] = K.node[j]; n_chr[id] = idx;
                    Kn.node[rank] = id; Kn.par[rank] = K.node[j]; Kn.len[rank] = K.len[j] + 1; Kn.last[rank] = idx;
                    Kn.hash[rank] = mix_hash(K.hash[j], idx); Kn.phash[rank] = K.hash[j];
                }
            }
        }
        __syncthreads();                                 // also orders the trie writes in global memory for the block
        nkept = keep_n;
        cur_buf ^= 1;
    }

    // ---- texts.append(kept_beams[0].prefix) (:208)
    if (tid == 0) {
        const KeptState& K = kept[cur_buf];
        const int L = K.len[0];
        int node = K.node[0];
        for (int c = L - 1; c >= 0; --c) { out_idx[(long long)b * Tn + c] = n_chr[node]; node = n_parent[node]; }
        out_len[b] = L;
        status[b] = 0;
    }
}

}  // namespace hctr

using namespace hctr;

extern "C" {

int hctr_ctc_topk_logsoftmax(const void* logits, int dtype, int T, int B, int C, long long stride_t,
                             long long stride_b, int k, int32_t* topk_idx, float* topk_logp, float* lse, void* stream) {
    HCTR_CHECK(topk_idx && topk_logp && lse, HCTR_ERR_INVALID, "topk: null output");
    HCTR_CHECK(T >= 0 && B >= 0 && C > 0, HCTR_ERR_INVALID, "topk: bad shape");
    HCTR_CHECK(k >= 1 && k <= kMaxK && k <= C, HCTR_ERR_INVALID, "topk: search depth must be in [1,%d] and <= C (got %d)", kMaxK, k);
    HCTR_CHECK(dtype == HCTR_F32 || dtype == HCTR_BF16, HCTR_ERR_INVALID, "topk: bad dtype");
    if (T == 0 || B == 0) return HCTR_OK;
    HCTR_CHECK(logits != nullptr, HCTR_ERR_INVALID, "topk: null logits");
    const long long rows = (long long)T * B;
    HCTR_CHECK(rows < (1ll << 31), HCTR_ERR_INVALID, "topk: too many rows");
    const size_t esz = dtype == HCTR_F32 ? 4 : 2;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    static PerDeviceOnce once;                     // value = SM count of the device
    int dev, num_sms;
    if (once.need(dev)) {
        HCTR_CUDA(cudaFuncSetAttribute(ctc_topk_logsoftmax_kernel<float, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        HCTR_CUDA(cudaFuncSetAttribute(ctc_topk_logsoftmax_kernel<__nv_bfloat16, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        HCTR_CUDA(cudaFuncSetAttribute(ctc_topk_logsoftmax_kernel<float, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        HCTR_CUDA(cudaFuncSetAttribute(ctc_topk_logsoftmax_kernel<__nv_bfloat16, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        HCTR_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
        once.mark(dev, num_sms);
    } else {
        num_sms = once.get(dev);
    }
    // ---- one warp per row, the row read once in register chunks
    {
        const char* cv = getenv("HCTR_TOPK_CHUNK");                       // "0": off (A/B against the kernels below)
        const bool use_chunk = cv ? cv[0] != '0' : true;
        if (use_chunk) {
            const long long blocks = (rows + kTopkWarps - 1) / kTopkWarps;
            if (dtype == HCTR_F32)
                ctc_topk_chunk_kernel<float><<<(int)blocks, kTopkWarps * 32, 0, s>>>(
                    static_cast<const float*>(logits), rows, B, C, stride_t, stride_b, k, topk_idx, topk_logp, lse);
            else
                ctc_topk_chunk_kernel<__nv_bfloat16><<<(int)blocks, kTopkWarps * 32, 0, s>>>(
                    static_cast<const __nv_bfloat16*>(logits), rows, B, C, stride_t, stride_b, k, topk_idx, topk_logp, lse);
            HCTR_CUDA(cudaGetLastError());
            return HCTR_OK;
        }
    }
    // ---- one warp per row (two passes, the second from L2): rows too large for the shared-memory kernel
    {
        const char* wv = getenv("HCTR_TOPK_WARP");                        // "0" / "1": force the CTA-per-row / warp-per-row kernel
        const bool fits_smem = (size_t)C * 4 <= 160 * 1024;
        bool use_warp = dtype == HCTR_BF16 || !fits_smem;
        if (wv && wv[0] == '0' && fits_smem) use_warp = false;
        if (wv && wv[0] == '1') use_warp = true;
        if (use_warp) {
            const long long blocks = (rows + kTopkWarps - 1) / kTopkWarps;
            // four CTAs (32 warps) per SM, enforced by an unused dynamic shared-memory request: measured 0.476 ms against
            // 0.546 ms with the 40 warps per SM the register count alone allows (bf16, config-5 tensor)
            static PerDeviceOnce once_w;                                 // value = shared memory per SM
            int dev_w, smem_sm = 0;
            if (once_w.need(dev_w)) {
                int optin = 0;
                HCTR_CUDA(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev_w));
                HCTR_CUDA(cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev_w));
                HCTR_CUDA(cudaFuncSetAttribute(ctc_topk_warp_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - 10 * 1024));
                HCTR_CUDA(cudaFuncSetAttribute(ctc_topk_warp_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - 10 * 1024));
                once_w.mark(dev_w, smem_sm);
            } else {
                smem_sm = once_w.get(dev_w);
            }
            const size_t pad = (size_t)(smem_sm / 5) - 9 * 1024;          // static 8.3 KB + 1 KB reserved per CTA: 4 fit, 5 do not
            if (dtype == HCTR_F32)
                ctc_topk_warp_kernel<float><<<(int)blocks, kTopkWarps * 32, pad, s>>>(
                    static_cast<const float*>(logits), rows, B, C, stride_t, stride_b, k, topk_idx, topk_logp, lse);
            else
                ctc_topk_warp_kernel<__nv_bfloat16><<<(int)blocks, kTopkWarps * 32, pad, s>>>(
                    static_cast<const __nv_bfloat16*>(logits), rows, B, C, stride_t, stride_b, k, topk_idx, topk_logp, lse);
            HCTR_CUDA(cudaGetLastError());
            return HCTR_OK;
        }
    }
    const int buf_bytes = (int)(((size_t)C * esz + 32 + 127) & ~size_t(127));        // row + 16 bytes of misalignment either side
    // 128 threads per row (the default): one row buffer per CTA and as many CTAs as the shared memory holds (7 at C = 7375
    // fp32) - the other CTAs' copies are in flight while one reduces; 256 threads: two buffers per CTA (round 1).
    int threads = 128;
    if (const char* e = getenv("HCTR_TOPK_THREADS")) { if (atoi(e) == 256) threads = 256; }       // A/B measurements
    int nbuf = (threads == 256 && 2 * (size_t)buf_bytes <= 200 * 1024) ? 2 : 1;
    if (const char* e = getenv("HCTR_TOPK_NBUF")) { const int v = atoi(e); if ((v == 1 || v == 2) && (size_t)v * buf_bytes <= 200 * 1024) nbuf = v; }
    const size_t smem = (size_t)nbuf * buf_bytes;
    long long per_sm = (224 * 1024) / (long long)(smem + 6 * 1024);
    if (per_sm < 1) per_sm = 1;
    const long long per_sm_cap = threads == 256 ? 8 : 12;
    if (per_sm > per_sm_cap) per_sm = per_sm_cap;
    const long long grid = rows < per_sm * num_sms ? rows : per_sm * num_sms;
#define HCTR_TOPK_LAUNCH(TT, TH)                                                                                          \
    ctc_topk_logsoftmax_kernel<TT, TH><<<(int)grid, TH, smem, s>>>(static_cast<const TT*>(logits), rows, B, C, stride_t,   \
                                                                   stride_b, k, nbuf, buf_bytes, topk_idx, topk_logp, lse)
    if (dtype == HCTR_F32) { if (threads == 256) HCTR_TOPK_LAUNCH(float, 256); else HCTR_TOPK_LAUNCH(float, 128); }
    else { if (threads == 256) HCTR_TOPK_LAUNCH(__nv_bfloat16, 256); else HCTR_TOPK_LAUNCH(__nv_bfloat16, 128); }
#undef HCTR_TOPK_LAUNCH
    HCTR_CUDA(cudaGetLastError());
    return HCTR_OK;
}

static long long beam_ws_per_seq(int T, int beam) {
    const long long cap = (long long)T * beam + 1;
    long long bytes = 8ll * T + 8ll * cap + 8 /*align*/ + 8ll * cap;
    return (bytes + 15) & ~15ll;
}

long long hctr_ctc_beam_workspace_bytes(int T, int B, int beam_size) {
    if (T <= 0 || B <= 0 || beam_size <= 0) return 0;
    return beam_ws_per_seq(T, beam_size) * B;
}

int hctr_ctc_prefix_beam_search_lm(const int32_t* topk_idx, const float* topk_logp, int T, int B, int C, int k,
                                   int beam_size, double lm_penalty, double len_bonus, const double* lm_table,
                                   const hctr_ngram_lm* ngram, int32_t* out_idx, int32_t* out_len, int32_t* status,
                                   void* workspace, long long workspace_bytes, void* stream) {
    HCTR_CHECK(out_idx && out_len && status, HCTR_ERR_INVALID, "beam: null output");
    HCTR_CHECK(k >= 1 && k <= kMaxK, HCTR_ERR_INVALID, "beam: search depth must be in [1,%d] (got %d)", kMaxK, k);
    HCTR_CHECK(beam_size >= 1 && beam_size <= kMaxBeam, HCTR_ERR_INVALID, "beam: beam size must be in [1,%d] (got %d)", kMaxBeam, beam_size);
    HCTR_CHECK(T >= 0 && B >= 0 && C > 1, HCTR_ERR_INVALID, "beam: bad shape");
    HCTR_CHECK(!(ngram && lm_table), HCTR_ERR_INVALID, "beam: pass either a unigram table or an n-gram model");
    hctr_ngram_lm lm;
    memset(&lm, 0, sizeof(lm));
    if (ngram) {
        int rc = hctr::check_ngram(ngram, "beam");
        if (rc) return rc;
        HCTR_CHECK(ngram->num_ids >= C, HCTR_ERR_INVALID, "beam: the n-gram vocabulary map covers %d ids, the logits have %d classes", ngram->num_ids, C);
        lm = *ngram;
    }
    if (B == 0) return HCTR_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (T == 0) {        // no frames: the greedy path is empty for every sequence
        HCTR_CUDA(cudaMemsetAsync(out_len, 0, sizeof(int32_t) * B, s));
        HCTR_CUDA(cudaMemsetAsync(status, 0xff, sizeof(int32_t) * B, s));
        return HCTR_OK;
    }
    HCTR_CHECK(topk_idx && topk_logp, HCTR_ERR_INVALID, "beam: null input");
    HCTR_CHECK((long long)T * beam_size + 1 < (1ll << 31), HCTR_ERR_INVALID, "beam: sequence too long");
    const long long need = hctr_ctc_beam_workspace_bytes(T, B, beam_size);
    HCTR_CHECK(workspace && workspace_bytes >= need, HCTR_ERR_INVALID, "beam: workspace too small (%lld < %lld)", workspace_bytes, need);
    HCTR_CHECK((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, HCTR_ERR_INVALID, "beam: workspace must be 16-byte aligned");
    ctc_prefix_beam_kernel<<<B, kBeamThreads, 0, s>>>(topk_idx, topk_logp, T, B, C, k, beam_size, lm_penalty, len_bonus, lm_table, lm,
                                            out_idx, out_len, status, static_cast<unsigned char*>(workspace),
                                            beam_ws_per_seq(T, beam_size));
    HCTR_CUDA(cudaGetLastError());
    return HCTR_OK;
}

int hctr_ctc_prefix_beam_search(const int32_t* topk_idx, const float* topk_logp, int T, int B, int C, int k,
                                int beam_size, double lm_penalty, double len_bonus, const double* lm_table,
                                int32_t* out_idx, int32_t* out_len, int32_t* status, void* workspace,
                                long long workspace_bytes, void* stream) {
    return hctr_ctc_prefix_beam_search_lm(topk_idx, topk_logp, T, B, C, k, beam_size, lm_penalty, len_bonus, lm_table, nullptr,
                                          out_idx, out_len, status, workspace, workspace_bytes, stream);
}

}  // extern "C"
