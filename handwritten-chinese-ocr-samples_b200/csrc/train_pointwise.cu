// Train-mode pointwise passes of the backbone (HBM-bound), forward and backward.
//
// Reference semantics (models/handwritten_ctr_model.py, train()):
//   conv -> BatchNorm2d with batch statistics (:38,40,74-92) -> [SE gate (:26-30)] -> [+ residual (:57)] -> ReLU ->
//   [max_pool2d((2,1)) (:123-150)] -> Dropout (:45,96-99,59,130,...)
// Forward per layer:  z = conv(x)+bias (tcgen05 kernel, bf16)  ->  chan_stats(z)  ->  bn_finalize  ->  apply.
// Backward per layer: bwd_reduce(dout, z, ...) -> bwd_finalize (BN / SE-FC backward on [B,C] data) -> bwd_apply.
// With d_pre = gradient at the output of the affine (+gate, +residual) stage after the ReLU/pool/dropout masks,
//   dz = P[b,c]*d_pre + Q[b,c] + R[c]*z,   dres = d_pre
// where P,Q,R fold the BatchNorm backward (and the SE squeeze path) computed from per-(b,c) partial sums, so each
// direction touches every activation tensor exactly twice. All reductions are fixed-order (deterministic).
#include <cstring>

#include "common.cuh"
#include "../../include/hctr_b200.h"

namespace hctr {

constexpr int kStatPixMax = 2048;   // upper bound of pixels per reduction slice

// Pixels per slice so that B*slices gives ~8 blocks per SM even for the 2-lines-per-GPU training shape.
__host__ __device__ inline int stat_pix_per_slice(int B, int HW) {
    long long per = ((long long)B * HW + 1183) / 1184;
    per = (per + 31) / 32 * 32;
    if (per < 128) per = 128;
    if (per > kStatPixMax) per = kStatPixMax;
    return (int)per;
}

__device__ __forceinline__ uint32_t hash32(uint32_t x) {       // lowbias32
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
// Counter-based dropout: the keep-mask of output element (vector index `vec`, lane q in 0..7) is a pure function of
// (seed, vec, q), so forward and backward regenerate it instead of storing a mask. One full avalanche hash per 16-byte
// vector, then a 3-instruction finalizer per element.
__device__ __forceinline__ uint32_t drop_base(unsigned long long vec, uint32_t seed) {
    return hash32(static_cast<uint32_t>(vec) ^ hash32(seed + static_cast<uint32_t>(vec >> 32) * 0x9E3779B9u));
}
// keep-decision of element q of a vector: the vector's hashed base times one of eight odd constants (a bijection of the
// 32-bit base, so the keep rate is exact; pairs and the drop count of a vector measured binomial to sampling noise over 2^23
// bases). One IMAD + one compare per element: the first version ran a second full integer hash per element, seven
// half-rate ALU instructions each, and made the dropout variants of train_apply_fwd_kernel ALU-bound (66 % ALU pipe).
__device__ __forceinline__ bool drop_keep(uint32_t base, int q, uint32_t thresh) {
    constexpr uint32_t kMul[8] = {0x9E3779B1u, 0x85EBCA77u, 0xC2B2AE3Du, 0x27D4EB2Fu, 0x165667B1u, 0xD3A2646Du, 0xFD7046C5u, 0xB55A4F09u};
    return base * kMul[q] >= thresh;
}

__device__ __forceinline__ void unpack8(const uint4& q, float (&o)[8]) {
    o[0] = bf16_lo(q.x); o[1] = bf16_hi(q.x); o[2] = bf16_lo(q.y); o[3] = bf16_hi(q.y);
    o[4] = bf16_lo(q.z); o[5] = bf16_hi(q.z); o[6] = bf16_lo(q.w); o[7] = bf16_hi(q.w);
}
__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
    return make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
}

// ---------------------------------------------------------------- per-(b, slice, c) sum and sum of squares
__global__ void __launch_bounds__(256)
chan_stats_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ psum, float* __restrict__ psq, int HW, int C,
                  int slices, int pix_per_slice) {
    extern __shared__ float red[];                         // [2][groups][C]
    const int b = blockIdx.y, slice = blockIdx.x;
    const int vpp = C >> 3, groups = blockDim.x / vpp;
    const int g = threadIdx.x / vpp, v = threadIdx.x - g * vpp;
    const int p0 = slice * pix_per_slice, p1 = min(p0 + pix_per_slice, HW);
    float s[8] = {0, 0, 0, 0, 0, 0, 0, 0}, q[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const __nv_bfloat16* base = x + ((size_t)b * HW) * C + v * 8;
    for (int p = p0 + g; p < p1; p += groups) {
        float e[8];
        unpack8(ld_nc_v4(base + (size_t)p * C), e);
#pragma unroll
        for (int i = 0; i < 8; ++i) { s[i] += e[i]; q[i] = fmaf(e[i], e[i], q[i]); }
    }
    float* rs = red; float* rq = red + groups * C;
#pragma unroll
    for (int i = 0; i < 8; ++i) { rs[g * C + v * 8 + i] = s[i]; rq[g * C + v * 8 + i] = q[i]; }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float a = 0.f, d = 0.f;
        for (int gg = 0; gg < groups; ++gg) { a += rs[gg * C + c]; d += rq[gg * C + c]; }
        const size_t o = ((size_t)b * slices + slice) * C + c;
        psum[o] = a;
        if (psq) psq[o] = d;
    }
}

// Partial sums of two [B][slices][C] arrays over the slices `part, part + workers, ...` of lines b0 .. b0+LB-1 for channel c.
// All 2*LB*U loads of a step are issued before the first use, from clamped (always valid) addresses, and masked afterwards:
// written with `if (b0 + k < B)` around each load the compiler emits a branch per line and one round trip per load (measured:
// 0.9 us per loop iteration, 18-21 us for the 2-block launches of the 64-channel layers).
template <int LB, int U>
__device__ __forceinline__ void slice_partials(const float* __restrict__ a0, const float* __restrict__ a1, int b0, int B,
                                               int slices, int C, int c, int part, int workers, float (&x)[LB], float (&y)[LB],
                                               int line_stride = 0) {
    if (line_stride == 0) line_stride = slices;          // slices per line in memory (>= the `slices` that are summed)
#pragma unroll
    for (int k = 0; k < LB; ++k) { x[k] = 0.f; y[k] = 0.f; }
    size_t base[LB];
#pragma unroll
    for (int k = 0; k < LB; ++k) base[k] = (size_t)min(b0 + k, B - 1) * line_stride * C + c;
    for (int i = part; i < slices; i += U * workers) {
        float v[U][LB], w[U][LB];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int s = min(i + u * workers, slices - 1);
#pragma unroll
            for (int k = 0; k < LB; ++k) { v[u][k] = a0[base[k] + (size_t)s * C]; w[u][k] = a1[base[k] + (size_t)s * C]; }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const bool ok = i + u * workers < slices;
#pragma unroll
            for (int k = 0; k < LB; ++k) { x[k] += ok ? v[u][k] : 0.f; y[k] += ok ? w[u][k] : 0.f; }
        }
    }
}

// ---------------------------------------------------------------- BN batch statistics -> scale/shift (+running stats)
// One block = 32 channels (lanes) x 32 workers (warps); the workers split the slices of a line, their partials are combined
// in a fixed order. line_sum[b][c] = sum over (h,w) of z (used by the SE squeeze and by the backward).
// The first version had 8 workers per channel and C/32 blocks in all: at 2 lines per GPU a thread walked 32-73 dependent
// L2 round trips and the 33 launches of a step cost 0.48 ms (17-19 us each); 32 workers and four loads in flight per
// array leave two to five round trips.
constexpr int kFinWorkers = 32;

__global__ void __launch_bounds__(32 * kFinWorkers)
bn_finalize_kernel(const float* __restrict__ psum, const float* __restrict__ psq, int B, int slices, int line_stride,
                   int C, int HW, const float* __restrict__ gamma, const float* __restrict__ beta,
                   float eps, float momentum, float* __restrict__ running_mean,
                   float* __restrict__ running_var, float* __restrict__ mean_out,
                   float* __restrict__ invstd_out, float* __restrict__ scale_out,
                   float* __restrict__ shift_out, float* __restrict__ line_sum) {
    constexpr int LB = 2;                                  // lines reduced per round
    __shared__ float ps[LB][kFinWorkers][32];
    __shared__ double pq[LB][kFinWorkers][32];
    __shared__ float lt[LB][32];
    __shared__ double lu[LB][32];
    const int lane = threadIdx.x & 31, part = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + lane;
    double s = 0.0, q = 0.0;
    for (int b0 = 0; b0 < B; b0 += LB) {
        float ls[LB], lq[LB];
        // per-worker partials in fp32 (a worker adds at most slices/32 values), the cross-worker and cross-line sums of the
        // squares in fp64 below
        slice_partials<LB, 4>(psum, psq, b0, B, slices, C, min(c, C - 1), part, kFinWorkers, ls, lq, line_stride);
#pragma unroll
        for (int k = 0; k < LB; ++k) { ps[k][part][lane] = ls[k]; pq[k][part][lane] = (double)lq[k]; }
        __syncthreads();
        if (part < LB && b0 + part < B && c < C) {           // warp k combines the partials of line b0+k, fixed order
            float t = 0.f; double u = 0.0;
#pragma unroll
            for (int k = 0; k < kFinWorkers; ++k) { t += ps[part][k][lane]; u += pq[part][k][lane]; }
            if (line_sum) line_sum[(size_t)(b0 + part) * C + c] = t;
            lt[part][lane] = t; lu[part][lane] = u;
        }
        __syncthreads();
        if (part == 0 && c < C) {
            for (int k = 0; k < LB && b0 + k < B; ++k) { s += (double)lt[k][lane]; q += lu[k][lane]; }
        }
    }
    if (part != 0 || c >= C) return;
    const double n = (double)B * (double)HW;
    const double mean = s / n;
    double var = q / n - mean * mean;                      // biased variance (normalisation)
    if (var < 0.0) var = 0.0;
    const float invstd = (float)(1.0 / sqrt(var + (double)eps));
    const float sc = gamma[c] * invstd;
    mean_out[c] = (float)mean;
    invstd_out[c] = invstd;
    scale_out[c] = sc;
    shift_out[c] = beta[c] - (float)mean * sc;
    if (running_mean) {                                    // nn.BatchNorm2d: momentum update, unbiased variance
        const double unbiased = n > 1.0 ? var * n / (n - 1.0) : var;
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
    }
}

// ---------------------------------------------------------------- forward apply
struct ApplyParams {
    const __nv_bfloat16* z;      // [B][H][W][C]
    const float* scale;          // [C]
    const float* shift;          // [C]
    const float* gate;           // [B][C] or null
    const __nv_bfloat16* res;    // [B][H][W][C] or null
    int B, H, W, C;
    int relu, pool;
    float drop_p;                // 0 = off
    uint32_t seed;
    int vshift;                  // log2(C/8)
    uint8_t* mask;               // [B][H][W][C/8] one keep-bit per element at INPUT resolution (fwd: written; bwd: read)
};

// Work decomposition of the apply kernels: blockIdx.y walks the rows (b, h) of the iterated tensor, blockIdx.x splits a
// row of W * C/8 16-byte vectors into segments. C/8 is a power of two and divides the block size, so a thread keeps the
// same channel group for its whole life: the per-channel operands are loaded once, and no integer division runs in
// the element loop.
// Two vectors per thread and iteration, every load of both issued before the first use (one vector per iteration measured
// 3.6 TB/s), the per-(line, channel) SE gate loaded once per row of the tensor instead of once per vector: 10.5 -> 9.0 ms per
// 16-line step together with the cheaper dropout decision. Measured and dropped: four vectors per iteration (121 registers, two
// CTAs per SM: 13.8 ms; 80 registers with spills: 11.1 ms), 64 registers = four CTAs per SM with the affine operands in shared
// memory (10.5 ms) or spilled (10.1 ms).
constexpr int kApplyUnroll = 2;

__global__ void __launch_bounds__(256, 3)
train_apply_fwd_kernel(ApplyParams p, __nv_bfloat16* __restrict__ out) {
    const int vpp = p.C >> 3;
    const int Ho = p.pool ? p.H / 2 : p.H;
    const int rowlen = p.W * vpp;
    const uint32_t thresh = p.drop_p > 0.f ? (uint32_t)fminf(p.drop_p * 4294967296.f, 4294967295.f) : 0u;
    const float keep_scale = p.drop_p > 0.f ? 1.f / (1.f - p.drop_p) : 1.f;
    const int cv = threadIdx.x & (vpp - 1);
    const int c0 = cv * 8;
    float sc[8], sh[8];
    *reinterpret_cast<float4*>(sc) = __ldg(reinterpret_cast<const float4*>(p.scale + c0));
    *reinterpret_cast<float4*>(sc + 4) = __ldg(reinterpret_cast<const float4*>(p.scale + c0 + 4));
    *reinterpret_cast<float4*>(sh) = __ldg(reinterpret_cast<const float4*>(p.shift + c0));
    *reinterpret_cast<float4*>(sh + 4) = __ldg(reinterpret_cast<const float4*>(p.shift + c0 + 4));
    const int seg_len = ((rowlen + gridDim.x - 1) / gridDim.x + 255) / 256 * 256;
    const int j0 = blockIdx.x * seg_len, j1 = min(rowlen, j0 + seg_len);
    const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
    for (int row = blockIdx.y; row < p.B * Ho; row += gridDim.y) {
        const int b = row / Ho, ho = row - b * Ho;
        const size_t obase = (size_t)row * rowlen;                       // output vector index of (row, j=0)
        float g[8];
        if (p.gate) {
            *reinterpret_cast<float4*>(g) = __ldg(reinterpret_cast<const float4*>(p.gate + (size_t)b * p.C + c0));
            *reinterpret_cast<float4*>(g + 4) = __ldg(reinterpret_cast<const float4*>(p.gate + (size_t)b * p.C + c0 + 4));
        }
        for (int jb = j0 + threadIdx.x; jb < j1; jb += 256 * kApplyUnroll) {
            // ---- all loads of the kApplyUnroll vectors
            uint4 zq[kApplyUnroll], sq[kApplyUnroll];               // sq: the second pooled row, or the residual (never both)
            size_t zoff[kApplyUnroll];
#pragma unroll
            for (int u = 0; u < kApplyUnroll; ++u) {
                const int j = jb + 256 * u;
                zq[u] = zero4; sq[u] = zero4; zoff[u] = 0;
                if (j < j1) {
                    const int w = j >> p.vshift;
                    zoff[u] = p.pool ? (((size_t)b * p.H + 2 * ho) * p.W + w) * p.C + c0 : ((size_t)row * p.W + w) * p.C + c0;
                    zq[u] = ld_nc_v4(p.z + zoff[u]);
                    if (p.pool) sq[u] = ld_nc_v4(p.z + zoff[u] + (size_t)p.W * p.C);
                    else if (p.res) sq[u] = ld_nc_v4(p.res + zoff[u]);
                }
            }
            // ---- affine, gate, residual, ReLU / pool, dropout, store: the arithmetic of one vector at a time
#pragma unroll
            for (int u = 0; u < kApplyUnroll; ++u) {
                const int j = jb + 256 * u;
                if (j >= j1) break;
                const size_t i = obase + j;
                float v[8], e[8];
                unpack8(zq[u], e);
#pragma unroll
                for (int q = 0; q < 8; ++q) v[q] = fmaf(e[q], sc[q], sh[q]);
                if (p.gate) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) v[q] *= g[q];
                }
                if (p.res) {
                    float r[8];
                    unpack8(sq[u], r);
#pragma unroll
                    for (int q = 0; q < 8; ++q) v[q] += r[q];
                }
                const size_t mi = zoff[u] >> 3;                           // mask byte of this input-resolution vector
                if (p.pool) {
                    uint32_t keep = 0xffu;                               // dropout keep bits of the 8 output elements
                    if (p.drop_p > 0.f) {
                        const uint32_t base = drop_base(i, p.seed);
                        keep = 0u;
#pragma unroll
                        for (int q = 0; q < 8; ++q) keep |= drop_keep(base, q, thresh) ? (1u << q) : 0u;
                    }
                    float v1[8];
                    unpack8(sq[u], e);
#pragma unroll
                    for (int q = 0; q < 8; ++q) v1[q] = fmaf(e[q], sc[q], sh[q]);
                    uint32_t m0 = 0u, m1 = 0u;
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float r0 = p.relu ? fmaxf(v[q], 0.f) : v[q];
                        const float r1 = p.relu ? fmaxf(v1[q], 0.f) : v1[q];
                        const bool first = r0 >= r1;                     // torch max_pool2d: the first maximum takes the gradient
                        const bool pos = !p.relu || (first ? v[q] : v1[q]) > 0.f;
                        if (pos && first) m0 |= 1u << q;
                        if (pos && !first) m1 |= 1u << q;
                        v[q] = first ? r0 : r1;
                    }
                    if (p.mask) {
                        p.mask[mi] = (uint8_t)(m0 & keep);
                        p.mask[mi + (size_t)p.W * vpp] = (uint8_t)(m1 & keep);
                    }
                    if (p.drop_p > 0.f) {
#pragma unroll
                        for (int q = 0; q < 8; ++q) v[q] = ((keep >> q) & 1u) ? v[q] * keep_scale : 0.f;
                    }
                } else {
                    // ReLU and dropout are ONE decision per element: it passes iff it is positive and kept - one predicate
                    // (the integer compare chained onto the float one), one keep-bit, one select. Decided separately (ReLU
                    // mask + maximum, keep bits, then bit tests and selects) the pass spent ~10 integer-pipe instructions per
                    // element on them and was instruction-bound: ALU pipe 52 %, DRAM 43 % busy
                    // (profiles/r2_ncu_train_apply_fwd_kernel_B16_512ch.txt).
                    uint32_t m0 = 0xffu;
                    if (p.drop_p > 0.f) {
                        const uint32_t base = drop_base(i, p.seed);
                        m0 = 0u;
                        if (p.relu) {
#pragma unroll
                            for (int q = 0; q < 8; ++q) {
                                const bool pass = v[q] > 0.f && drop_keep(base, q, thresh);
                                m0 |= pass ? (1u << q) : 0u;
                                v[q] = pass ? v[q] * keep_scale : 0.f;
                            }
                        } else {
#pragma unroll
                            for (int q = 0; q < 8; ++q) {
                                const bool pass = drop_keep(base, q, thresh);
                                m0 |= pass ? (1u << q) : 0u;
                                v[q] = pass ? v[q] * keep_scale : 0.f;
                            }
                        }
                    } else if (p.relu) {
                        m0 = 0u;
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const bool pass = v[q] > 0.f;
                            m0 |= pass ? (1u << q) : 0u;
                            v[q] = pass ? v[q] : 0.f;
                        }
                    }
                    if (p.mask) p.mask[mi] = (uint8_t)m0;
                }
                *reinterpret_cast<uint4*>(out + i * 8) = pack8(v);
            }
        }
    }
}

// ---------------------------------------------------------------- backward: d_pre from the stored keep-mask
// d_pre = gradient behind the dropout / pool / ReLU masks for the 8 channels of one INPUT-resolution pixel. The forward
// stored one keep-bit per element (ReLU sign, pool winner and dropout keep folded together), so the backward reads
// 1 byte per 16-byte vector instead of recomputing the affine/gate/residual chain and the dropout hash.
__device__ __forceinline__ void bwd_dpre(const ApplyParams& p, const __nv_bfloat16* __restrict__ dout, int b, int h, int w,
                                         int c0, float keep_scale, float (&zv)[8], float (&d)[8]) {
    const int vpp = p.C >> 3;
    const size_t ivec = (((size_t)b * p.H + h) * p.W + w) * vpp + (c0 >> 3);
    unpack8(ld_nc_v4(p.z + ivec * 8), zv);
    const uint32_t m = p.mask[ivec];
    const int Ho = p.pool ? p.H / 2 : p.H;
    const int ho = p.pool ? (h >> 1) : h;
    const size_t ovec = (((size_t)b * Ho + ho) * p.W + w) * vpp + (c0 >> 3);
    float g[8];
    unpack8(ld_nc_v4(dout + ovec * 8), g);
#pragma unroll
    for (int i = 0; i < 8; ++i) d[i] = ((m >> i) & 1u) ? g[i] * keep_scale : 0.f;
}

// per-(b, slice, c): A2 = sum d_pre, A3 = sum d_pre * z
__global__ void __launch_bounds__(256, 4)
train_bwd_reduce_kernel(ApplyParams p, const __nv_bfloat16* __restrict__ dout, float* __restrict__ pA2,
                        float* __restrict__ pA3, int slices, int pix_per_slice) {
    extern __shared__ float red[];
    const int b = blockIdx.y, slice = blockIdx.x;
    const int vpp = p.C >> 3, groups = blockDim.x / vpp;
    const int g = threadIdx.x / vpp, cv = threadIdx.x - g * vpp;
    const int HW = p.H * p.W;
    const int p0 = slice * pix_per_slice, p1 = min(p0 + pix_per_slice, HW);
    const float keep_scale = p.drop_p > 0.f ? 1.f / (1.f - p.drop_p) : 1.f;
    const int c0 = cv * 8;
    float a2[8] = {0, 0, 0, 0, 0, 0, 0, 0}, a3[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int h = (p0 + g) / p.W, w = (p0 + g) - h * p.W;
    for (int px = p0 + g; px < p1; px += groups, w += groups) {
        while (w >= p.W) { w -= p.W; ++h; }
        float zv[8], d[8];
        bwd_dpre(p, dout, b, h, w, c0, keep_scale, zv, d);
#pragma unroll
        for (int i = 0; i < 8; ++i) { a2[i] += d[i]; a3[i] = fmaf(d[i], zv[i], a3[i]); }
    }
    float* r2 = red; float* r3 = red + groups * p.C;
#pragma unroll
    for (int i = 0; i < 8; ++i) { r2[g * p.C + c0 + i] = a2[i]; r3[g * p.C + c0 + i] = a3[i]; }
    __syncthreads();
    for (int c = threadIdx.x; c < p.C; c += blockDim.x) {
        float x = 0.f, y = 0.f;
        for (int gg = 0; gg < groups; ++gg) { x += r2[gg * p.C + c]; y += r3[gg * p.C + c]; }
        const size_t o = ((size_t)b * slices + slice) * p.C + c;
        pA2[o] = x; pA3[o] = y;
    }
}

// Combine the per-slice partial sums of every (b,c) item into slice 0, in a fixed order: one block = 32 items (lanes) x 32
// workers (warps). Only the SE layers need this pass (their finalize reads the totals of ALL channels); the other layers
// reduce inside the finalize kernel. (First version: 64 items x 4 workers, 64-146 dependent loads per thread, 13-26 us.)
__global__ void __launch_bounds__(32 * kFinWorkers)
bwd_slice_reduce_kernel(float* __restrict__ pA2, float* __restrict__ pA3, int slices, int B, int C) {
    __shared__ float s2[kFinWorkers][32], s3[kFinWorkers][32];
    const int lane = threadIdx.x & 31, part = threadIdx.x >> 5;
    const int item = blockIdx.x * 32 + lane;                // C % 32 == 0: the 32 items of a block share their line
    const int items = B * C;
    float x = 0.f, y = 0.f;
    size_t o0 = 0;
    if (item < items) {
        const int b = item / C, c = item - b * C;
        o0 = ((size_t)b * slices) * C + c;
#pragma unroll 4
        for (int s = part; s < slices; s += kFinWorkers) { x += pA2[o0 + (size_t)s * C]; y += pA3[o0 + (size_t)s * C]; }
    }
    s2[part][lane] = x; s3[part][lane] = y;
    __syncthreads();
    if (part < 2 && item < items) {
        const float (*src)[32] = part ? s3 : s2;
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < kFinWorkers; ++k) t += src[k][lane];
        (part ? pA3 : pA2)[o0] = t;
    }
}

// BN (+SE) backward on [B,C]-sized data.
struct BwdFinalizeParams {
    const float* pA2; const float* pA3; int slices;       // SE layers: slice 0 holds the totals (bwd_slice_reduce_kernel ran first)
    int B, C, HW;
    const float* gamma; const float* mean; const float* invstd; const float* scale; const float* shift;
    const float* line_sum;       // [B][C] sum_hw z (forward)
    // SE (all null when the layer has no gate)
    const float* gate;           // [B][C] sigmoid output
    const float* se_hidden;      // [B][Cr] relu(W1 m)
    const float* se_mean;        // [B][C] m = mean_hw(bn(z))
    const float* w1; const float* w2; int Cr;
    float* dw1; float* dw2;      // [Cr][C], [C][Cr]
    // outputs
    float* dgamma; float* dbeta; float* dbias;     // [C]; dbias = sum dz (conv bias gradient)
    float* P; float* Q;          // [B][C]
    float* R;                    // [C]
};

// One block per 32 channels (the first version was ONE block for the layer: 47 us for a 512-channel SE layer, 6 us otherwise,
// behind a 13-26 us slice reduction). Everything after the slice sums is local to a channel except the SE hidden-layer
// gradient da[b][r] = sum_c W2[c][r] du[b][c], which every block recomputes from the [B][C] totals (B*Cr*C MACs: microseconds)
// instead of exchanging partials. Layers without a gate reduce their own channels' slices here: one launch per layer.
// shared memory: A2o, A3o, dmo [B][32] | t1, t2 [B][32] doubles | sc [2][32] doubles | red [2*LB][32][32] (no gate)
//                or du [B][C], da [B][Cr], pda [<=1024] (gate)
__global__ void __launch_bounds__(32 * kFinWorkers)
train_bwd_finalize_kernel(BwdFinalizeParams p) {
    extern __shared__ __align__(16) unsigned char fin_smem[];
    const int B = p.B, C = p.C, Cr = p.Cr;
    double* t1 = reinterpret_cast<double*>(fin_smem);
    double* t2 = t1 + B * 32;
    double* sc = t2 + B * 32;                               // S1[32], R[32]
    float* A2o = reinterpret_cast<float*>(sc + 64);
    float* A3o = A2o + B * 32;
    float* dmo = A3o + B * 32;
    float* extra = dmo + B * 32;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int lane = tid & 31, part = tid >> 5;
    const int c0 = blockIdx.x * 32;
    const int c = c0 + lane;

    if (!p.gate) {
        // slice sums of this block's 32 channels, LB lines per round, fixed-order combine
        constexpr int LB = 4;
        float (*red)[kFinWorkers][32] = reinterpret_cast<float (*)[kFinWorkers][32]>(extra);     // [2*LB]
        for (int b0 = 0; b0 < B; b0 += LB) {
            float x[LB], y[LB];
            slice_partials<LB, 2>(p.pA2, p.pA3, b0, B, p.slices, C, c, part, kFinWorkers, x, y);
#pragma unroll
            for (int k = 0; k < LB; ++k) { red[k][part][lane] = x[k]; red[LB + k][part][lane] = y[k]; }
            __syncthreads();
            if (part < 2 * LB) {
                const int k = part < LB ? part : part - LB;
                if (b0 + k < B) {
                    float t = 0.f;
#pragma unroll
                    for (int w = 0; w < kFinWorkers; ++w) t += red[part][w][lane];
                    (part < LB ? A2o : A3o)[(b0 + k) * 32 + lane] = t;
                }
            }
            __syncthreads();
        }
        for (int i = tid; i < B * 32; i += nt) dmo[i] = 0.f;
        __syncthreads();
    } else {
        float* du = extra; float* da = du + (size_t)B * C; float* pda = da + B * Cr;
        // dgate[b,c] = sum_hw d_pre * bn(z) = scale*A3 + shift*A2 ; du = dgate * g (1-g)     (all channels)
        for (int i = tid; i < B * C; i += nt) {
            const int b = i / C, cc = i - b * C;
            const size_t o = ((size_t)b * p.slices) * C + cc;
            const float a2 = p.pA2[o], a3 = p.pA3[o];
            const float dg = p.scale[cc] * a3 + p.shift[cc] * a2;
            const float g = p.gate[i];
            du[i] = dg * g * (1.f - g);
            if (cc >= c0 && cc < c0 + 32) { A2o[b * 32 + cc - c0] = a2; A3o[b * 32 + cc - c0] = a3; }
        }
        __syncthreads();
        // da[b,r] = [hidden>0] * sum_c W2[c][r] du[b,c]: item = (b, r), the channel range split over nt/items chunks
        const int items = B * Cr;
        for (int i0 = 0; i0 < items; i0 += nt) {
            const int n_here = min(items - i0, nt);
            int chunks = nt / n_here;
            if (chunks > C) chunks = C;
            const int per = (C + chunks - 1) / chunks;
            const int ch = tid / n_here, li = tid - ch * n_here;
            if (ch < chunks) {
                const int it = i0 + li, b = it / Cr, r = it - b * Cr;
                const int ca = ch * per, cb = min(C, ca + per);
                float s = 0.f;
                for (int cc = ca; cc < cb; ++cc) s = fmaf(__ldg(p.w2 + (size_t)cc * Cr + r), du[b * C + cc], s);
                pda[ch * n_here + li] = s;
            }
            __syncthreads();
            if (tid < n_here) {
                float s = 0.f;
                for (int k = 0; k < chunks; ++k) s += pda[k * n_here + tid];
                da[i0 + tid] = p.se_hidden[i0 + tid] > 0.f ? s : 0.f;
            }
            __syncthreads();
        }
        // own channels: dW2[c][r] = sum_b du[b,c] * hidden[b,r]
        for (int i = tid; i < 32 * Cr; i += nt) {
            const int cl = i / Cr, r = i - cl * Cr;
            float s = 0.f;
            for (int b = 0; b < B; ++b) s = fmaf(du[b * C + c0 + cl], p.se_hidden[b * Cr + r], s);
            p.dw2[(size_t)(c0 + cl) * Cr + r] = s;
        }
        // dW1[r][c] = sum_b da[b,r] m[b,c] ; dm[b,c] = sum_r W1[r][c] da[b,r]
        for (int i = tid; i < Cr * 32; i += nt) {
            const int r = i >> 5, cl = i & 31;
            float s = 0.f;
            for (int b = 0; b < B; ++b) s = fmaf(da[b * Cr + r], p.se_mean[(size_t)b * C + c0 + cl], s);
            p.dw1[(size_t)r * C + c0 + cl] = s;
        }
        for (int i = tid; i < B * 32; i += nt) {
            const int b = i >> 5, cl = i & 31;
            float s = 0.f;
            for (int r = 0; r < Cr; ++r) s = fmaf(p.w1[(size_t)r * C + c0 + cl], da[b * Cr + r], s);
            dmo[i] = s;
        }
        __syncthreads();
    }

    // per-(b, channel) terms in fp64, summed over the lines in line order by one thread per channel
    const double n = (double)B * (double)p.HW;
    const double mu = p.mean[c], is = p.invstd[c];
    for (int i = tid; i < B * 32; i += nt) {
        const int b = i >> 5;                               // i & 31 == lane
        const size_t gi = (size_t)b * C + c;
        const double g = p.gate ? (double)p.gate[gi] : 1.0;
        const double a2 = A2o[i], a3 = A3o[i];
        const double dmb = dmo[i];
        const double x1 = ((double)p.line_sum[gi] - (double)p.HW * mu) * is;     // sum_hw xhat
        t1[i] = g * a2 + dmb;
        t2[i] = g * (a3 - mu * a2) * is + dmb / (double)p.HW * x1;
    }
    __syncthreads();
    const double gis = (double)p.gamma[c] * is;
    if (part == 0) {
        double S1 = 0.0, S2 = 0.0;
        for (int b = 0; b < B; ++b) { S1 += t1[b * 32 + lane]; S2 += t2[b * 32 + lane]; }
        p.dgamma[c] = (float)S2;
        p.dbeta[c] = (float)S1;
        const double R = -gis * is * S2 / n;
        p.R[c] = (float)R;
        sc[lane] = S1; sc[32 + lane] = R;
    }
    __syncthreads();
    {
        const double S1 = sc[lane], R = sc[32 + lane];
        for (int i = tid; i < B * 32; i += nt) {
            const int b = i >> 5;
            const size_t gi = (size_t)b * C + c;
            const double g = p.gate ? (double)p.gate[gi] : 1.0;
            const double Pv = gis * g;
            const double Qv = gis * ((double)dmo[i] / (double)p.HW - S1 / n) - R * mu;
            p.P[gi] = (float)Pv;
            p.Q[gi] = (float)Qv;
            t1[i] = Pv * (double)A2o[i] + Qv * (double)p.HW + R * (double)p.line_sum[gi];
        }
    }
    __syncthreads();
    if (part == 0 && p.dbias) {
        double dbias = 0.0;
        for (int b = 0; b < B; ++b) dbias += t1[b * 32 + lane];
        p.dbias[c] = (float)dbias;
    }
}

// dz = P*d_pre + Q + R*z ; dres = d_pre
__global__ void __launch_bounds__(256, 4)
train_bwd_apply_kernel(ApplyParams p, const __nv_bfloat16* __restrict__ dout, const float* __restrict__ P,
                       const float* __restrict__ Q, const float* __restrict__ R, __nv_bfloat16* __restrict__ dz,
                       __nv_bfloat16* __restrict__ dres) {
    const int vpp = p.C >> 3;
    const int rowlen = p.W * vpp;
    const float keep_scale = p.drop_p > 0.f ? 1.f / (1.f - p.drop_p) : 1.f;
    const int cv = threadIdx.x & (vpp - 1);
    const int c0 = cv * 8;
    float Rv[8];
    *reinterpret_cast<float4*>(Rv) = __ldg(reinterpret_cast<const float4*>(R + c0));
    *reinterpret_cast<float4*>(Rv + 4) = __ldg(reinterpret_cast<const float4*>(R + c0 + 4));
    const int seg_len = ((rowlen + gridDim.x - 1) / gridDim.x + 255) / 256 * 256;
    const int j0 = blockIdx.x * seg_len, j1 = min(rowlen, j0 + seg_len);
    for (int row = blockIdx.y; row < p.B * p.H; row += gridDim.y) {
        const int b = row / p.H, h = row - b * p.H;
        float Pv[8], Qv[8];
        *reinterpret_cast<float4*>(Pv) = __ldg(reinterpret_cast<const float4*>(P + (size_t)b * p.C + c0));
        *reinterpret_cast<float4*>(Pv + 4) = __ldg(reinterpret_cast<const float4*>(P + (size_t)b * p.C + c0 + 4));
        *reinterpret_cast<float4*>(Qv) = __ldg(reinterpret_cast<const float4*>(Q + (size_t)b * p.C + c0));
        *reinterpret_cast<float4*>(Qv + 4) = __ldg(reinterpret_cast<const float4*>(Q + (size_t)b * p.C + c0 + 4));
        const size_t ibase = (size_t)row * rowlen;
        const int Ho = p.pool ? p.H / 2 : p.H;
        const size_t obase = ((size_t)b * Ho + (p.pool ? (h >> 1) : h)) * rowlen;    // the row of dout this input row reads
        // two vectors per thread and iteration, all four 16-byte loads issued before the first use: the pass is latency-bound
        // (one vector per iteration kept 16-32 KB in flight per SM; DRAM 46 % busy)
        for (int j = j0 + threadIdx.x; j < j1; j += 512) {
            const bool two = j + 256 < j1;
            const size_t i0 = ibase + j, i1 = i0 + 256;
            const uint4 zq0 = ld_nc_v4(p.z + i0 * 8);
            const uint4 gq0 = ld_nc_v4(dout + (obase + j) * 8);
            const uint32_t m0 = p.mask[i0];
            uint4 zq1 = zq0, gq1 = gq0;
            uint32_t m1 = 0u;
            if (two) { zq1 = ld_nc_v4(p.z + i1 * 8); gq1 = ld_nc_v4(dout + (obase + j + 256) * 8); m1 = p.mask[i1]; }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                if (u == 1 && !two) break;
                float zv[8], g[8], d[8], o[8];
                unpack8(u ? zq1 : zq0, zv);
                unpack8(u ? gq1 : gq0, g);
                const uint32_t m = u ? m1 : m0;
#pragma unroll
                for (int q = 0; q < 8; ++q) d[q] = ((m >> q) & 1u) ? g[q] * keep_scale : 0.f;
#pragma unroll
                for (int q = 0; q < 8; ++q) o[q] = fmaf(Pv[q], d[q], fmaf(Rv[q], zv[q], Qv[q]));
                const size_t i = u ? i1 : i0;
                *reinterpret_cast<uint4*>(dz + i * 8) = pack8(o);
                if (dres) *reinterpret_cast<uint4*>(dres + i * 8) = pack8(d);
            }
        }
    }
}

// SE gate from the BN-folded line means: m = scale*mean_hw(z)+shift ; hidden = relu(W1 m) ; gate = sigmoid(W2 hidden)
__global__ void __launch_bounds__(512)
se_excite_train_kernel(const float* __restrict__ line_sum, const float* __restrict__ scale, const float* __restrict__ shift,
                       const float* __restrict__ w1, const float* __restrict__ w2, float* __restrict__ se_mean,
                       float* __restrict__ hidden_out, float* __restrict__ gate, int C, int Cr, float inv_hw) {
    extern __shared__ float sm[];
    float* mean = sm; float* hidden = sm + C;
    const int b = blockIdx.x;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const float m = fmaf(line_sum[(size_t)b * C + c] * inv_hw, scale[c], shift[c]);
        mean[c] = m; se_mean[(size_t)b * C + c] = m;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int r = warp; r < Cr; r += nwarps) {
        float s = 0.f;
        for (int c = lane; c < C; c += 32) s = fmaf(w1[(size_t)r * C + c], mean[c], s);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) { hidden[r] = fmaxf(s, 0.f); hidden_out[(size_t)b * Cr + r] = fmaxf(s, 0.f); }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float s = 0.f;
        for (int r = 0; r < Cr; ++r) s = fmaf(w2[(size_t)c * Cr + r], hidden[r], s);
        gate[(size_t)b * C + c] = 1.f / (1.f + expf(-s));
    }
}

static bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
// grid (segments per row, rows): enough blocks for ~8 per SM, at least 256 vectors per block
static dim3 grid_rows(long long rows, int rowlen) {
    long long gy = rows < 32768 ? rows : 32768;
    long long segs = (1184 + gy - 1) / gy;
    const long long max_segs = (rowlen + 255) / 256;
    if (segs > max_segs) segs = max_segs;
    if (segs < 1) segs = 1;
    return dim3((unsigned)segs, (unsigned)gy);
}

}  // namespace hctr

using namespace hctr;

extern "C" {

int hctr_stat_slices(int B, int H, int W) {
    const int per = stat_pix_per_slice(B, H * W);
    return (H * W + per - 1) / per;
}

int hctr_chan_stats(const void* x, float* psum, float* psq, int B, int H, int W, int C, void* stream) {
    HCTR_CHECK(x && psum, HCTR_ERR_INVALID, "chan_stats: null pointer");
    HCTR_CHECK(C % 8 == 0 && C >= 8 && C <= 2048 && 256 % (C / 8) == 0, HCTR_ERR_INVALID, "chan_stats: C/8 must divide 256 (C=%d)", C);
    HCTR_CHECK(al16(x), HCTR_ERR_INVALID, "chan_stats: x must be 16-byte aligned");
    const int slices = hctr_stat_slices(B, H, W);
    const int groups = 256 / (C / 8);
    dim3 grid(slices, B);
    chan_stats_kernel<<<grid, 256, 2 * (size_t)groups * C * sizeof(float), static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(x), psum, psq, H * W, C, slices, stat_pix_per_slice(B, H * W));
    HCTR_CUDA(cudaGetLastError());
    return HCTR_OK;
}

int hctr_bn_finalize_train(const float* psum, const float* psq, int B, int slices, int C, int HW, const float* gamma,
                           const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                           float* mean, float* invstd, float* scale, float* shift, float* line_sum, void* stream) {
    HCTR_CHECK(psum && psq && gamma && beta && mean && invstd && scale && shift, HCTR_ERR_INVALID, "bn_finalize: null pointer");
    int used = slices;
    if (slices > 2 * kFinWorkers) {
        // many partials per line (the conv epilogue's per-(row, span, warp) slots, or long lines): one block per (line, 32
        // channels) first adds a line's partials into its slice 0, in a fixed order
        bwd_slice_reduce_kernel<<<(B * C + 31) / 32, 32 * kFinWorkers, 0, static_cast<cudaStream_t>(stream)>>>(
            const_cast<float*>(psum), const_cast<float*>(psq), slices, B, C);
        HCTR_CUDA(cudaGetLastError());
        used = 1;
    }
    bn_finalize_kernel<<<(C + 31) / 32, 32 * kFinWorkers, 0, static_cast<cudaStream_t>(stream)>>>(
        psum, psq, B, used, slices, C, HW, gamma, beta, eps, momentum, running_mean, running_var, mean, invstd, scale, shift, line_sum);
    HCTR_CUDA(cudaGetLastError());
    return HCTR_OK;
}

int hctr_se_excite_train(const float* line_sum, const float* scale, const float* shift, const float* w1, const float* w2,
                         float* se_mean, float* hidden, float* gate, int B, int C, int Cr, int HW, void* stream) {
    HCTR_CHECK(line_sum && scale && shift && w1 && w2 && se_mean && hidden && gate, HCTR_ERR_INVALID, "se_excite_train: null pointer");
    se_excite_train_kernel<<<B, 512, (size_t)(C + Cr) * sizeof(float), static_cast<cudaStream_t>(stream)>>>(
        line_sum, scale, shift, w1, w2, se_mean, hidden, gate, C, Cr, 1.0f / (float)HW);
    HCTR_CUDA(cudaGetLastError());
    return HCTR_OK;
}

static int fill_apply(ApplyParams& p, const void* z, const float* scale, const float* shift, const float* gate,
                      const void* res, int B, int H, int W, int C, int relu, int pool, float drop_p, unsigned seed) {
    HCTR_CHECK(z && scale && shift, HCTR_ERR_INVALID, "train_apply: null pointer");
    HCTR_CHECK(C % 8 == 0 && C >= 8 && 256 % (C / 8) == 0, HCTR_ERR_INVALID, "train_apply: bad channel count %d", C);
    HCTR_CHECK(!pool || (H % 2 == 0 && !gate && !res), HCTR_ERR_INVALID, "train_apply: pooling needs even H and no gate/residual");
    HCTR_CHECK(drop_p >= 0.f && drop_p < 1.f, HCTR_ERR_INVALID, "train_apply: dropout p must be in [0,1)");
    HCTR_CHECK(al16(z) && al16(scale) && al16(shift) && al16(gate) && al16(res), HCTR_ERR_INVALID, "train_apply: 16-byte alignment");
    p.z = static_cast<const __nv_bfloat16*>(z); p.scale = scale; p.shift = shift; p.gate = gate;
    p.res = static_cast<const __nv_bfloat16*>(res);
    p.B = B; p.H = H; p.W = W; p.C = C; p.relu = relu; p.pool = pool; p.drop_p = drop_p; p.seed = seed;
    p.vshift = 0;
    while ((1 << p.vshift) < C / 8) ++p.vshift;
    HCTR_CHECK((1 << p.vshift) == C / 8, HCTR_ERR_INVALID, "train_apply: C/8 must be a power of two (C=%d)", C);
    return HCTR_OK;
}

int hctr_train_apply_fwd(const void* z, const float* scale, const float* shift, const float* gate, const void* res,
                         void* out, void* mask, int B, int H, int W, int C, int relu, int pool, float drop_p, unsigned seed,
                         void* stream) {
    ApplyParams p;
    int rc = fill_apply(p, z, scale, shift, gate, res, B, H, W, C, relu, pool, drop_p, seed);
    if (rc) return rc;
    p.mask = static_cast<uint8_t*>(mask);
    HCTR_CHECK(out && al16(out), HCTR_ERR_INVALID, "train_apply_fwd: bad output");
    train_apply_fwd_kernel<<<grid_rows((long long)B * (pool ? H / 2 : H), W * (C / 8)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        p, static_cast<__nv_bfloat16*>(out));
    HCTR_CUDA(cudaGetLastError());
    return HCTR_OK;
}

static int fill_bwd(ApplyParams& p, const void* z, const void* mask, int B, int H, int W, int C, int pool, float drop_p) {
    HCTR_CHECK(z && mask, HCTR_ERR_INVALID, "train_bwd: null pointer");
    HCTR_CHECK(C % 8 == 0 && C >= 8 && 256 % (C / 8) == 0, HCTR_ERR_INVALID, "train_bwd: bad channel count %d", C);
    HCTR_CHECK(!pool || H % 2 == 0, HCTR_ERR_INVALID, "train_bwd: pooling needs an even H");
    HCTR_CHECK(drop_p >= 0.f && drop_p < 1.f, HCTR_ERR_INVALID, "train_bwd: dropout p must be in [0,1)");
    HCTR_CHECK(al16(z), HCTR_ERR_INVALID, "train_bwd: 16-byte alignment");
    memset(&p, 0, sizeof(p));
    p.z = static_cast<const __nv_bfloat16*>(z);
    p.mask = const_cast<uint8_t*>(static_cast<const uint8_t*>(mask));
    p.B = B; p.H = H; p.W = W; p.C = C; p.pool = pool; p.drop_p = drop_p;
    p.vshift = 0;
    while ((1 << p.vshift) < C / 8) ++p.vshift;
    HCTR_CHECK((1 << p.vshift) == C / 8, HCTR_ERR_INVALID, "train_bwd: C/8 must be a power of two (C=%d)", C);
    return HCTR_OK;
}

int hctr_train_bwd_reduce(const void* dout, const void* z, const void* mask, float* pA2, float* pA3, int B, int H, int W,
                          int C, int pool, float drop_p, void* stream) {
    ApplyParams p;
    int rc = fill_bwd(p, z, mask, B, H, W, C, pool, drop_p);
    if (rc) return rc;
    HCTR_CHECK(dout && pA2 && pA3 && al16(dout), HCTR_ERR_INVALID, "train_bwd_reduce: null pointer");
    const int slices = hctr_stat_slices(B, H, W);
    const int groups = 256 / (C / 8);
    dim3 grid(slices, B);
    train_bwd_reduce_kernel<<<grid, 256, 2 * (size_t)groups * C * sizeof(float), static_cast<cudaStream_t>(stream)>>>(
        p, static_cast<const __nv_bfloat16*>(dout), pA2, pA3, slices, stat_pix_per_slice(B, H * W));
    HCTR_CUDA(cudaGetLastError());
    return HCTR_OK;
}

int hctr_train_bwd_finalize(const float* pA2, const float* pA3, int slices, int B, int C, int HW, const float* gamma,
                            const float* mean, const float* invstd, const float* scale, const float* shift,
                            const float* line_sum, const float* gate, const float* se_hidden, const float* se_mean,
                            const float* w1, const float* w2, int Cr, float* dw1, float* dw2, float* dgamma,
                            float* dbeta, float* dbias, float* P, float* Q, float* R, void* stream) {
    HCTR_CHECK(pA2 && pA3 && gamma && mean && invstd && scale && shift && line_sum && dgamma && dbeta && P && Q && R,
               HCTR_ERR_INVALID, "train_bwd_finalize: null pointer");
    HCTR_CHECK(!gate || (se_hidden && se_mean && w1 && w2 && dw1 && dw2 && Cr > 0), HCTR_ERR_INVALID, "train_bwd_finalize: SE arguments");
    BwdFinalizeParams p;
    p.pA2 = pA2; p.pA3 = pA3; p.slices = slices; p.B = B; p.C = C; p.HW = HW;
    p.gamma = gamma; p.mean = mean; p.invstd = invstd; p.scale = scale; p.shift = shift; p.line_sum = line_sum;
    p.gate = gate; p.se_hidden = se_hidden; p.se_mean = se_mean; p.w1 = w1; p.w2 = w2; p.Cr = gate ? Cr : 0;
    p.dw1 = dw1; p.dw2 = dw2; p.dgamma = dgamma; p.dbeta = dbeta; p.dbias = dbias; p.P = P; p.Q = Q; p.R = R;
    HCTR_CHECK(C % 32 == 0, HCTR_ERR_INVALID, "train_bwd_finalize: C must be a multiple of 32 (C=%d)", C);
    const size_t smem = (size_t)(2 * B * 32 + 64) * sizeof(double) + (size_t)3 * B * 32 * sizeof(float) +
                        (gate ? ((size_t)B * C + (size_t)B * Cr + 1024) : (size_t)8 * kFinWorkers * 32) * sizeof(float);
    HCTR_CHECK(smem <= 200 * 1024, HCTR_ERR_INVALID, "train_bwd_finalize: B*C too large for a block's shared memory (%zu bytes)", smem);
    static PerDeviceOnce once;
    int dev;
    if (once.need(dev)) {
        HCTR_CUDA(cudaFuncSetAttribute(train_bwd_finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        once.mark(dev);
    }
    if (gate) {
        bwd_slice_reduce_kernel<<<(B * C + 31) / 32, 32 * kFinWorkers, 0, static_cast<cudaStream_t>(stream)>>>(
            const_cast<float*>(pA2), const_cast<float*>(pA3), slices, B, C);
        HCTR_CUDA(cudaGetLastError());
    }
    train_bwd_finalize_kernel<<<C / 32, 32 * kFinWorkers, smem, static_cast<cudaStream_t>(stream)>>>(p);
    HCTR_CUDA(cudaGetLastError());
    return HCTR_OK;
}

int hctr_train_bwd_apply(const void* dout, const void* z, const void* mask, const float* P, const float* Q, const float* R,
                         void* dz, void* dres, int B, int H, int W, int C, int pool, float drop_p, void* stream) {
    ApplyParams p;
    int rc = fill_bwd(p, z, mask, B, H, W, C, pool, drop_p);
    if (rc) return rc;
    HCTR_CHECK(dout && P && Q && R && dz && al16(dout) && al16(dz) && al16(dres) && al16(P) && al16(Q) && al16(R),
               HCTR_ERR_INVALID, "train_bwd_apply: bad pointer");
    train_bwd_apply_kernel<<<grid_rows((long long)B * H, W * (C / 8)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        p, static_cast<const __nv_bfloat16*>(dout), P, Q, R, static_cast<__nv_bfloat16*>(dz), static_cast<__nv_bfloat16*>(dres));
    HCTR_CUDA(cudaGetLastError());
    return HCTR_OK;
}

}  // extern "C"
