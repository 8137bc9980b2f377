// HBM-bound parts of the backbone: the 1->64 stem conv (CUDA cores, 9 FLOP/B), the SE squeeze /
// excite / scale+residual+ReLU passes. All activations are NHWC bf16; loads/stores are 16-byte
// vectors over the channel dimension so that a warp always touches whole 128-byte lines.
#include "common.cuh"
#include "../../include/hctr_b200.h"

namespace hctr {

// ---------------------------------------------------------------- stem: conv0_1 + bn0_1 + relu
// reference: models/handwritten_ctr_model.py:116-118. One block walks one image row in 256-column segments: the
// 3 x 258 input halo of a segment sits in shared memory, the thread's 8 x 9 weights stay in registers for the whole row;
// one thread = one pixel x 8 output channels, 8 consecutive lanes cover the 64 channels of a pixel (one 128-byte
// line), so a warp writes 512 contiguous bytes per iteration. No integer division in the pixel loop.
constexpr int kStemSeg = 256;

__global__ void __launch_bounds__(256, 2)
stem_conv_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ scale,
                 const float* __restrict__ shift, __nv_bfloat16* __restrict__ y, int B, int H, int W, int relu) {
    __shared__ float tile[3][kStemSeg + 2];
    const int cg = threadIdx.x & 7, pl = threadIdx.x >> 3;
    // Channel pairs as packed fp32x2 operands: 36 + 4 FFMA2 per pixel instead of 72 + 8 FFMA, the ReLU inside the bf16
    // conversion. Bit-identical to the scalar form (each half is an IEEE fma); the first version was issue-bound at 16
    // thread-instructions per output element (57 % issue-active, 2.5 TB/s, profiles/r2_ncu_stem.txt).
    unsigned long long wr2[4][9], sc2[4], sh2[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
#pragma unroll
        for (int t = 0; t < 9; ++t) wr2[c][t] = f32x2_pack(__ldg(w + (cg * 8 + 2 * c) * 9 + t), __ldg(w + (cg * 8 + 2 * c + 1) * 9 + t));
        sc2[c] = f32x2_pack(__ldg(scale + cg * 8 + 2 * c), __ldg(scale + cg * 8 + 2 * c + 1));
        sh2[c] = f32x2_pack(__ldg(shift + cg * 8 + 2 * c), __ldg(shift + cg * 8 + 2 * c + 1));
    }
    for (int row = blockIdx.x; row < B * H; row += gridDim.x) {
        const int b = row / H, h = row - b * H;
        const float* img = x + (size_t)b * H * W;
        __nv_bfloat16* yrow = y + ((size_t)row * W) * 64;
        for (int w0 = 0; w0 < W; w0 += kStemSeg) {
            __syncthreads();                                     // previous segment fully consumed
            for (int i = threadIdx.x; i < 3 * (kStemSeg + 2); i += blockDim.x) {
                const int r = i / (kStemSeg + 2), c = i - r * (kStemSeg + 2);
                const int hh = h + r - 1, ww = w0 + c - 1;
                tile[r][c] = (hh >= 0 && hh < H && ww >= 0 && ww < W) ? __ldg(img + (size_t)hh * W + ww) : 0.f;
            }
            __syncthreads();
#pragma unroll 1
            for (int it = 0; it < kStemSeg / 32; ++it) {
                const int px = it * 32 + pl;
                const int wq = w0 + px;
                if (wq >= W) break;
                unsigned long long in2[9];
#pragma unroll
                for (int kh = 0; kh < 3; ++kh)
#pragma unroll
                    for (int kw = 0; kw < 3; ++kw) { const float v = tile[kh][px + kw]; in2[kh * 3 + kw] = f32x2_pack(v, v); }
                uint32_t pk[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    unsigned long long a = f32x2_pack(0.f, 0.f);
#pragma unroll
                    for (int t = 0; t < 9; ++t) a = f32x2_fma(in2[t], wr2[c][t], a);
                    a = f32x2_fma(a, sc2[c], sh2[c]);
                    float a0, a1;
                    f32x2_unpack(a, a0, a1);
                    pk[c] = relu ? pack_bf16x2_relu(a0, a1) : pack_bf16x2(a0, a1);
                }
                *reinterpret_cast<uint4*>(yrow + (size_t)wq * 64 + cg * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
        }
    }
}

// ---------------------------------------------------------------- SE squeeze (stage 1 of 2)
// reference: SELayer.forward avg_pool (models/handwritten_ctr_model.py:27-28). Grid (slices, B); every block
// sums a contiguous pixel range of one line for all channels in a fixed order -> partial[b][slice][c].
constexpr int kSePixPerSlice = 2048;

__global__ void __launch_bounds__(256)
se_squeeze_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ partial, int HW, int C, int slices) {
    extern __shared__ float red[];                         // [groups][C]
    const int b = blockIdx.y, slice = blockIdx.x;
    const int vec_per_pix = C >> 3;                        // 16-byte vectors per pixel
    const int groups = blockDim.x / vec_per_pix;           // pixels processed in parallel
    const int g = threadIdx.x / vec_per_pix;
    const int v = threadIdx.x - g * vec_per_pix;
    const int p0 = slice * kSePixPerSlice;
    const int p1 = min(p0 + kSePixPerSlice, HW);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (g < groups) {
        const __nv_bfloat16* base = x + ((size_t)b * HW) * C + v * 8;
        for (int p = p0 + g; p < p1; p += groups) {
            const uint4 q = ld_nc_v4(base + (size_t)p * C);
            acc[0] += bf16_lo(q.x); acc[1] += bf16_hi(q.x);
            acc[2] += bf16_lo(q.y); acc[3] += bf16_hi(q.y);
            acc[4] += bf16_lo(q.z); acc[5] += bf16_hi(q.z);
            acc[6] += bf16_lo(q.w); acc[7] += bf16_hi(q.w);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) red[g * C + v * 8 + i] = acc[i];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float s = 0.f;
        for (int gg = 0; gg < groups; ++gg) s += red[gg * C + c];     // fixed order: deterministic
        partial[((size_t)b * slices + slice) * C + c] = s;
    }
}

// ---------------------------------------------------------------- SE excite (stage 2 + the two FCs)
// reference: SELayer.fc (models/handwritten_ctr_model.py:19-24,29). One block per line.
__global__ void __launch_bounds__(512)
se_excite_kernel(const float* __restrict__ partial, int slices, const float* __restrict__ w1,
                 const float* __restrict__ w2, float* __restrict__ gate, int C, int Cr, float inv_hw) {
    extern __shared__ float sm[];                          // mean[C] + hidden[Cr] + part[groups][C]
    float* mean = sm;
    float* hidden = sm + C;
    float* part = sm + ((C + Cr + 3) & ~3);                 // 16-byte aligned rows
    const int b = blockIdx.x;
    // stage 2 of the squeeze: the block's threads form `groups` workers per float4 column; worker g adds the slices
    // i = g, g+groups, ... (four loads in flight), the workers' sums are combined in a fixed order
    const int ncol4 = C >> 2;
    const int groups = blockDim.x / ncol4;
    const int col = threadIdx.x % ncol4, g = threadIdx.x / ncol4;
    if (g < groups) {
        const float4* src = reinterpret_cast<const float4*>(partial + (size_t)b * slices * C) + col;
        float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, a2 = a0, a3 = a0;
        int i = g;
        for (; i + 3 * groups < slices; i += 4 * groups) {
            const float4 v0 = __ldg(src + (size_t)i * ncol4), v1 = __ldg(src + (size_t)(i + groups) * ncol4);
            const float4 v2 = __ldg(src + (size_t)(i + 2 * groups) * ncol4), v3 = __ldg(src + (size_t)(i + 3 * groups) * ncol4);
            a0.x += v0.x; a0.y += v0.y; a0.z += v0.z; a0.w += v0.w;
            a1.x += v1.x; a1.y += v1.y; a1.z += v1.z; a1.w += v1.w;
            a2.x += v2.x; a2.y += v2.y; a2.z += v2.z; a2.w += v2.w;
            a3.x += v3.x; a3.y += v3.y; a3.z += v3.z; a3.w += v3.w;
        }
        for (; i < slices; i += groups) {
            const float4 v0 = __ldg(src + (size_t)i * ncol4);
            a0.x += v0.x; a0.y += v0.y; a0.z += v0.z; a0.w += v0.w;
        }
        float4 t;
        t.x = (a0.x + a1.x) + (a2.x + a3.x); t.y = (a0.y + a1.y) + (a2.y + a3.y);
        t.z = (a0.z + a1.z) + (a2.z + a3.z); t.w = (a0.w + a1.w) + (a2.w + a3.w);
        reinterpret_cast<float4*>(part + (size_t)g * C)[col] = t;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float s = 0.f;
        for (int k = 0; k < groups; ++k) s += part[(size_t)k * C + c];
        mean[c] = s * inv_hw;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int r = warp; r < Cr; r += nwarps) {
        float s = 0.f;
        for (int c = lane; c < C; c += 32) s = fmaf(w1[(size_t)r * C + c], mean[c], s);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) hidden[r] = fmaxf(s, 0.f);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float s = 0.f;
        for (int r = 0; r < Cr; ++r) s = fmaf(w2[(size_t)c * Cr + r], hidden[r], s);
        gate[(size_t)b * C + c] = 1.f / (1.f + expf(-s));
    }
}


// ---------------------------------------------------------------- SE gate computed from the conv's INPUT
// BasicBlock.forward (models/handwritten_ctr_model.py:47-58): out = relu(se(bn2(conv2(t))) + residual), and the SE gate needs
// mean_{h,w} of z = bn2(conv2(t)) (:27-28) - a full-tensor dependency that forces z to be written and read again. But the
// mean of a convolution output is linear in its input:
//     sum_{h,w} z[co] = scale[co] * sum_{tap,ci} W[co,tap,ci] * S_tap[ci] + H*W*shift[co],
//     S_tap[ci] = sum of t[.., ci] over the pixels that tap (dh,dw) reads for some output pixel (zero padding excluded)
//               = total - (first or last row) - (first or last column) + (the corner removed twice).
// So the gate is known BEFORE conv2 runs, from per-channel sums of t (accumulated in conv1's epilogue), the border rows
// and columns of t and a [C x 9C] mat-vec with conv2's own bf16 weights; conv2's epilogue then applies gate, residual
// and ReLU and z never exists in memory. Three small kernels: sums (grid B x kGateSplit), mean of z (grid B x C/64), FCs.
constexpr int kGateSplit = 16;        // blocks per line in the sums kernel
constexpr int kGateCo = 64;           // output channels per block in the mean kernel

// psum[b][g][5][C]: partial {total, first row, last row, first column, last column} sums of t over block g's share
__global__ void __launch_bounds__(256)
se_gate_sums_kernel(const __nv_bfloat16* __restrict__ t, const float* __restrict__ partial, int slices,
                    float* __restrict__ psum, int H, int W, int C) {
    extern __shared__ float part[];                  // [groups][C]
    const int b = blockIdx.x, g = blockIdx.y, tid = threadIdx.x;
    const int nv = C >> 3;                           // 16-byte vectors per pixel
    const int groups = blockDim.x / nv;
    const int vec = tid % nv, grp = tid / nv;
    const __nv_bfloat16* tb = t + (size_t)b * H * W * C;
    float* out = psum + ((size_t)b * kGateSplit + g) * 5 * C;
    // totals: this block adds the producer's slices g, g + kGateSplit, ... (fixed order)
    {
        const int ncol4 = C >> 2, g4 = blockDim.x / ncol4;
        const int col = tid % ncol4, sub = tid / ncol4;
        if (sub < g4) {
            const float4* src = reinterpret_cast<const float4*>(partial + (size_t)b * slices * C) + col;
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int i = g + sub * kGateSplit; i < slices; i += g4 * kGateSplit) {
                const float4 v = __ldg(src + (size_t)i * ncol4);
                a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
            }
            reinterpret_cast<float4*>(part + (size_t)sub * C)[col] = a;
        }
        __syncthreads();
        for (int c = tid; c < C; c += blockDim.x) {
            float s = 0.f;
            for (int k = 0; k < g4; ++k) s += part[(size_t)k * C + c];
            out[c] = s;
        }
        __syncthreads();
    }
    // border rows (first, last) and columns (first, last): this block's share of the pixels
    for (int which = 0; which < 4; ++which) {
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
        const int n = which < 2 ? W : H;
        if (grp < groups) {
            for (int i = g + grp * kGateSplit; i < n; i += groups * kGateSplit) {
                const size_t pix = which == 0 ? (size_t)i : which == 1 ? (size_t)(H - 1) * W + i
                                 : which == 2 ? (size_t)i * W : (size_t)i * W + (W - 1);
                const uint4 q = ld_nc_v4(tb + pix * C + vec * 8);
                acc[0] += bf16_lo(q.x); acc[1] += bf16_hi(q.x); acc[2] += bf16_lo(q.y); acc[3] += bf16_hi(q.y);
                acc[4] += bf16_lo(q.z); acc[5] += bf16_hi(q.z); acc[6] += bf16_lo(q.w); acc[7] += bf16_hi(q.w);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) part[(size_t)grp * C + vec * 8 + j] = acc[j];
        }
        __syncthreads();
        for (int c = tid; c < C; c += blockDim.x) {
            float s = 0.f;
            for (int k = 0; k < groups; ++k) s += part[(size_t)k * C + c];
            out[(1 + which) * C + c] = s;
        }
        __syncthreads();
    }
}

// mean_z[b][co] for this block's kGateCo output channels
__global__ void __launch_bounds__(512)
se_gate_mean_kernel(const __nv_bfloat16* __restrict__ t, const float* __restrict__ psum, const __nv_bfloat16* __restrict__ wp,
                    const float* __restrict__ scale, const float* __restrict__ shift, float* __restrict__ mean_z,
                    int H, int W, int C) {
    extern __shared__ float sm[];
    float* S = sm;                                   // [9][C] window sums
    float* bord = S + 9 * C;                         // [8][C]: R0, RL, C0, CL, corners 00, 0L, L0, LL
    const int b = blockIdx.x, tid = threadIdx.x;
    const __nv_bfloat16* tb = t + (size_t)b * H * W * C;
    const float* ps = psum + (size_t)b * kGateSplit * 5 * C;
    for (int i = tid; i < 5 * C; i += blockDim.x) {
        float s = 0.f;
        for (int g = 0; g < kGateSplit; ++g) s += ps[(size_t)g * 5 * C + i];     // fixed order
        if (i < C) S[4 * C + i] = s;                 // centre tap (dh = dw = 0) reads every pixel
        else bord[i - C] = s;
    }
    for (int c = tid; c < C; c += blockDim.x) {
        bord[4 * C + c] = __bfloat162float(tb[c]);
        bord[5 * C + c] = __bfloat162float(tb[(size_t)(W - 1) * C + c]);
        bord[6 * C + c] = __bfloat162float(tb[(size_t)(H - 1) * W * C + c]);
        bord[7 * C + c] = __bfloat162float(tb[((size_t)(H - 1) * W + (W - 1)) * C + c]);
    }
    __syncthreads();
    // window sums per tap: tap = kh*3 + kw reads input (h + kh - 1, w + kw - 1)
    for (int i = tid; i < 9 * C; i += blockDim.x) {
        const int tap = i / C, c = i - tap * C;
        if (tap == 4) continue;
        const int dh = tap / 3 - 1, dw = tap % 3 - 1;
        float s = S[4 * C + c];
        if (dh == -1) s -= bord[1 * C + c];          // the last row is never read
        if (dh == 1) s -= bord[0 * C + c];           // the first row is never read
        if (dw == -1) s -= bord[3 * C + c];
        if (dw == 1) s -= bord[2 * C + c];
        if (dh != 0 && dw != 0)                      // the corner removed with its row and again with its column
            s += bord[(4 + (dh == -1 ? 2 : 0) + (dw == -1 ? 1 : 0)) * C + c];
        S[tap * C + c] = s;
    }
    __syncthreads();
    // one warp per output channel: 9*C-long dot product with conv2's packed bf16 weights [co][tap][ci]
    const int warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
    const int K = 9 * C;
    const float inv_hw = 1.0f / ((float)H * (float)W);
    const int co_end = min(C, (int)(blockIdx.y + 1) * kGateCo);
    for (int co = blockIdx.y * kGateCo + warp; co < co_end; co += nwarps) {
        const __nv_bfloat16* wr = wp + (size_t)co * K;
        float s = 0.f;
        for (int k = lane * 8; k < K; k += 256) {
            const uint4 q = ld_nc_v4(wr + k);
            s = fmaf(bf16_lo(q.x), S[k + 0], s); s = fmaf(bf16_hi(q.x), S[k + 1], s);
            s = fmaf(bf16_lo(q.y), S[k + 2], s); s = fmaf(bf16_hi(q.y), S[k + 3], s);
            s = fmaf(bf16_lo(q.z), S[k + 4], s); s = fmaf(bf16_hi(q.z), S[k + 5], s);
            s = fmaf(bf16_lo(q.w), S[k + 6], s); s = fmaf(bf16_hi(q.w), S[k + 7], s);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) mean_z[(size_t)b * C + co] = fmaf(s * inv_hw, scale[co], shift[co]);
    }
}

// ---------------------------------------------------------------- SE scale + residual + ReLU
// reference: BasicBlock.forward tail (models/handwritten_ctr_model.py:30,54-58); dropout is identity in eval().
__global__ void __launch_bounds__(256)
se_scale_residual_relu_kernel(const uint4* __restrict__ x, const float* __restrict__ gate,
                              const uint4* __restrict__ res, uint4* __restrict__ y, long long nvec, int vec_per_line,
                              int vec_per_pix, int C) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
        const int b = (int)(i / vec_per_line);
        const int c0 = (int)(i % vec_per_pix) * 8;
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gate + (size_t)b * C + c0));
        const float4 g1 = __ldg(reinterpret_cast<const float4*>(gate + (size_t)b * C + c0 + 4));
        const uint4 a = ld_nc_v4(x + i);
        const uint4 r = ld_nc_v4(res + i);
        uint4 o;
        o.x = pack_bf16x2(fmaxf(fmaf(bf16_lo(a.x), g0.x, bf16_lo(r.x)), 0.f), fmaxf(fmaf(bf16_hi(a.x), g0.y, bf16_hi(r.x)), 0.f));
        o.y = pack_bf16x2(fmaxf(fmaf(bf16_lo(a.y), g0.z, bf16_lo(r.y)), 0.f), fmaxf(fmaf(bf16_hi(a.y), g0.w, bf16_hi(r.y)), 0.f));
        o.z = pack_bf16x2(fmaxf(fmaf(bf16_lo(a.z), g1.x, bf16_lo(r.z)), 0.f), fmaxf(fmaf(bf16_hi(a.z), g1.y, bf16_hi(r.z)), 0.f));
        o.w = pack_bf16x2(fmaxf(fmaf(bf16_lo(a.w), g1.z, bf16_lo(r.w)), 0.f), fmaxf(fmaf(bf16_hi(a.w), g1.w, bf16_hi(r.w)), 0.f));
        y[i] = o;
    }
}

static bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace hctr

using namespace hctr;

extern "C" {

int hctr_stem_conv_fwd(const float* x, const float* w, const float* scale, const float* shift, void* y, int B, int H,
                       int W, int relu, void* stream) {
    HCTR_CHECK(x && w && scale && shift && y, HCTR_ERR_INVALID, "stem: null pointer");
    HCTR_CHECK(B > 0 && H > 0 && W > 0, HCTR_ERR_INVALID, "stem: empty tensor");
    HCTR_CHECK(al16(y), HCTR_ERR_INVALID, "stem: output must be 16-byte aligned");
    long long rows = (long long)B * H;
    HCTR_CHECK(rows < (1ll << 31), HCTR_ERR_INVALID, "stem: too many rows");
    const int grid = (int)(rows < 148 * 16 ? rows : 148 * 16);
    stem_conv_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        x, w, scale, shift, static_cast<__nv_bfloat16*>(y), B, H, W, relu);
    HCTR_CUDA(cudaGetLastError());
    return HCTR_OK;
}

int hctr_se_slices(int H, int W) {
    const long long hw = (long long)H * W;
    return (int)((hw + kSePixPerSlice - 1) / kSePixPerSlice);
}

int hctr_se_squeeze(const void* x, float* partial, int B, int H, int W, int C, void* stream) {
    HCTR_CHECK(x && partial, HCTR_ERR_INVALID, "se_squeeze: null pointer");
    HCTR_CHECK(C % 8 == 0 && C >= 8 && C <= 2048 && 256 % (C / 8) == 0, HCTR_ERR_INVALID,
               "se_squeeze: C/8 must divide 256 (got C=%d)", C);
    HCTR_CHECK(al16(x), HCTR_ERR_INVALID, "se_squeeze: x must be 16-byte aligned");
    const int slices = hctr_se_slices(H, W);
    const int groups = 256 / (C / 8);
    const size_t smem = (size_t)groups * C * sizeof(float);
    dim3 grid(slices, B);
    se_squeeze_kernel<<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(x), partial, H * W, C, slices);
    HCTR_CUDA(cudaGetLastError());
    return HCTR_OK;
}

int hctr_se_excite(const float* partial, int slices, const float* w1, const float* w2, float* gate, int B, int C,
                   int Cr, int HW, void* stream) {
    HCTR_CHECK(partial && w1 && w2 && gate, HCTR_ERR_INVALID, "se_excite: null pointer");
    HCTR_CHECK(C > 0 && Cr > 0 && slices > 0 && HW > 0, HCTR_ERR_INVALID, "se_excite: bad shape");
    HCTR_CHECK(C % 4 == 0 && C / 4 <= 512 && al16(partial), HCTR_ERR_INVALID, "se_excite: C must be a multiple of 4, <= 2048, partials 16-byte aligned");
    const int groups = 512 / (C / 4);
    const size_t smem = ((size_t)((C + Cr + 3) & ~3) + (size_t)groups * C) * sizeof(float);
    se_excite_kernel<<<B, 512, smem, static_cast<cudaStream_t>(stream)>>>(partial, slices, w1, w2, gate, C, Cr,
                                                                           1.0f / (float)HW);
    HCTR_CUDA(cudaGetLastError());
    return HCTR_OK;
}

long long hctr_se_gate_workspace_bytes(int B, int C) {
    if (B <= 0 || C <= 0) return 0;
    return ((long long)B * kGateSplit * 5 * C + (long long)B * C) * (long long)sizeof(float);
}

int hctr_se_gate_from_input(const void* t, const float* partial, int slices, const void* conv_w_packed, const float* scale,
                            const float* shift, const float* w1, const float* w2, float* gate, int B, int H, int W, int C,
                            int Cr, void* workspace, long long workspace_bytes, void* stream) {
    HCTR_CHECK(t && partial && conv_w_packed && scale && shift && w1 && w2 && gate, HCTR_ERR_INVALID, "se_gate_from_input: null pointer");
    HCTR_CHECK(B > 0 && H > 0 && W > 0 && slices > 0 && Cr > 0, HCTR_ERR_INVALID, "se_gate_from_input: bad shape");
    HCTR_CHECK(C % 32 == 0 && C >= 32 && 256 % (C / 8) == 0, HCTR_ERR_INVALID, "se_gate_from_input: C/8 must divide 256 (C=%d)", C);
    HCTR_CHECK(al16(t) && al16(partial) && al16(conv_w_packed), HCTR_ERR_INVALID, "se_gate_from_input: 16-byte alignment");
    HCTR_CHECK(workspace && al16(workspace) && workspace_bytes >= hctr_se_gate_workspace_bytes(B, C), HCTR_ERR_INVALID,
               "se_gate_from_input: workspace too small or misaligned");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    float* psum = static_cast<float*>(workspace);
    float* mean_z = psum + (size_t)B * kGateSplit * 5 * C;
    const int groups = 256 / (C / 8), g4 = 256 / (C / 4);
    const size_t sm1 = (size_t)(groups > g4 ? groups : g4) * C * sizeof(float);
    se_gate_sums_kernel<<<dim3(B, kGateSplit), 256, sm1, s>>>(static_cast<const __nv_bfloat16*>(t), partial, slices, psum, H, W, C);
    HCTR_CUDA(cudaGetLastError());
    const size_t sm2 = (size_t)17 * C * sizeof(float);
    HCTR_CHECK(sm2 <= 200 * 1024, HCTR_ERR_INVALID, "se_gate: %d channels do not fit shared memory", C);
    static PerDeviceOnce once;
    int dev;
    if (once.need(dev)) {
        HCTR_CUDA(cudaFuncSetAttribute(se_gate_mean_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        once.mark(dev);
    }
    se_gate_mean_kernel<<<dim3(B, (C + kGateCo - 1) / kGateCo), 512, sm2, s>>>(
        static_cast<const __nv_bfloat16*>(t), psum, static_cast<const __nv_bfloat16*>(conv_w_packed), scale, shift, mean_z, H, W, C);
    HCTR_CUDA(cudaGetLastError());
    // SELayer.fc on the means (the squeeze stage of se_excite degenerates to one slice with weight 1)
    return hctr_se_excite(mean_z, 1, w1, w2, gate, B, C, Cr, 1, stream);
}

int hctr_se_scale_residual_relu(const void* x, const float* gate, const void* residual, void* y, int B, int H, int W,
                                int C, void* stream) {
    HCTR_CHECK(x && gate && residual && y, HCTR_ERR_INVALID, "se_scale: null pointer");
    HCTR_CHECK(C % 8 == 0, HCTR_ERR_INVALID, "se_scale: C must be a multiple of 8");
    HCTR_CHECK(al16(x) && al16(residual) && al16(y) && al16(gate), HCTR_ERR_INVALID, "se_scale: 16-byte alignment");
    const long long vec_per_line = (long long)H * W * C / 8;
    HCTR_CHECK(vec_per_line < (1ll << 31), HCTR_ERR_INVALID, "se_scale: line too large");
    const long long nvec = vec_per_line * B;
    long long blocks = (nvec + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    se_scale_residual_relu_kernel<<<(int)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const uint4*>(x), gate, static_cast<const uint4*>(residual), static_cast<uint4*>(y), nvec,
        (int)vec_per_line, C / 8, C);
    HCTR_CUDA(cudaGetLastError());
    return HCTR_OK;
}

}  // extern "C"
