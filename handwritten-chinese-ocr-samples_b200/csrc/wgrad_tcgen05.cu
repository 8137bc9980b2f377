// Weight-gradient GEMM on tcgen05 for the backbone convolutions and the classifier (training backward):
//
//   dW[tap][m][n] = sum over pixels (b,h,w) of  G[b, h, w, m] * X[b, h + dh(tap), w + dw(tap), n]
//
// G = gradient wrt the conv output (NHWC bf16, M = Cout channels), X = the conv input (NHWC bf16, N = Cin). The
// reduction dimension K is the PIXEL index, so both operands are MN-major in shared memory exactly as TMA delivers an
// NHWC box: a (64 channels x 64 pixels) box is 64 rows (pixels = K) of 128 bytes (channels = M/N) with SWIZZLE_128B;
// boxes for the next 64 channels follow at +8 KB (descriptor LBO), groups of 8 pixels at +1 KB (SBO). Zero padding of
// the conv (and ragged widths) is again the TMA out-of-bounds fill. One CTA owns NUM_SUB accumulators of 128 x BLOCK_N
// in TMEM for one (tap, m-tile, n-tile, k-split) work item and writes an fp32 partial; a second kernel reduces the
// splits in a fixed order and scatters to the reference's OIHW layout (deterministic).
// Reference op: autograd of nn.Conv2d / nn.Linear weights (main.py:426 loss.backward()).
#include <cstring>
#include <mutex>

#include "common.cuh"
#include "../../include/hctr_b200.h"

namespace hctr {

constexpr int kWgPix = 64;           // pixels (K) per pipeline stage
constexpr int kWgThreads = 192;

struct WgradParams {
    int B, H, W;                     // pixel grid of G
    int w_blocks;                    // ceil(W / 64)
    int kblocks_total;               // B * H * w_blocks
    int kblocks_per_split, splits;
    int ntaps;
    int8_t tap_dh[9], tap_dw[9];
    int M, N;                        // Cout, Cin
    int m_tiles, n_tiles;            // tiles of NUM_SUB*128 and BLOCK_N
    int total_items;                 // ntaps * m_tiles * n_tiles * splits
    float* partial;                  // [splits][ntaps][M][N] fp32
};

__device__ __forceinline__ uint64_t make_sw128_mnmajor_desc(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3ffffu) >> 4);
    d |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;       // LBO: next 64-element group along M/N
    d |= static_cast<uint64_t>(1024 >> 4) << 32;            // SBO: next group of 8 rows along K
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;                    // SWIZZLE_128B
    return d;
}
__host__ __device__ constexpr uint32_t make_idesc_bf16_mn(int m, int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) /*A MN-major*/ | (1u << 16) /*B MN-major*/ |
           (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}

template <int BLOCK_N, int NUM_SUB, int STAGES>
struct WgSmem {
    static constexpr int kBoxBytes = 64 * kWgPix * 2;                  // 8 KB: 64 channels x 64 pixels
    static constexpr int kABytes = NUM_SUB * 2 * kBoxBytes;            // M = NUM_SUB * 128 channels
    static constexpr int kBBytes = (BLOCK_N / 64) * kBoxBytes;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kTotal = STAGES * kStageBytes + 1024 + 1024;
};

template <int BLOCK_N, int NUM_SUB, int STAGES>
__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_tcgen05_kernel(const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmX, const WgradParams p) {
    using L = WgSmem<BLOCK_N, NUM_SUB, STAGES>;
    constexpr int kTmemCols = (NUM_SUB * BLOCK_N <= 32) ? 32 : (NUM_SUB * BLOCK_N <= 64) ? 64 : (NUM_SUB * BLOCK_N <= 128) ? 128
                              : (NUM_SUB * BLOCK_N <= 256) ? 256 : 512;
    static_assert(NUM_SUB * BLOCK_N <= 512, "TMEM");
    extern __shared__ uint8_t smem_raw[];
    // 1 KB alignment by POINTER arithmetic on the __shared__ array: rounding the address as an integer and casting it back made
    // every later access through `smem` a generic LD/ST (the compiler no longer knew the address space) - 495 generic loads and
    // 380 generic stores in the conv kernels' epilogues, long-scoreboard stalls at the staged stores (ncu source view)
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * L::kStageBytes);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* acc_full = empty_bar + STAGES;
    uint64_t* acc_empty = acc_full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmG); tma_prefetch_desc(&tmX);
        for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        mbar_init(acc_full, 1); mbar_init(acc_empty, 4);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    auto decode = [&](int item, int& split, int& tap, int& mt, int& nt) {
        nt = item % p.n_tiles; item /= p.n_tiles;
        mt = item % p.m_tiles; item /= p.m_tiles;
        tap = item % p.ntaps; split = item / p.ntaps;
    };

    if (warp == 0) {
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0;
            for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
                int split, tap, mt, nt; decode(item, split, tap, mt, nt);
                const int kb0 = split * p.kblocks_per_split;
                const int kb1 = min(kb0 + p.kblocks_per_split, p.kblocks_total);
                for (int kb = kb0; kb < kb1; ++kb) {
                    const int wb = kb % p.w_blocks; int r = kb / p.w_blocks;
                    const int h = r % p.H; const int b = r / p.H;
                    const int w0 = wb * kWgPix;
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    mbar_arrive_expect_tx(&full_bar[stage], L::kStageBytes);
                    uint8_t* st = smem + stage * L::kStageBytes;
#pragma unroll
                    for (int j = 0; j < NUM_SUB * 2; ++j)
                        tma_load_4d(st + j * L::kBoxBytes, &tmG, &full_bar[stage], mt * NUM_SUB * 128 + j * 64, w0, h, b);
#pragma unroll
                    for (int j = 0; j < BLOCK_N / 64; ++j)
                        tma_load_4d(st + L::kABytes + j * L::kBoxBytes, &tmX, &full_bar[stage], nt * BLOCK_N + j * 64,
                                    w0 + p.tap_dw[tap], h + p.tap_dh[tap], b);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = make_idesc_bf16_mn(128, BLOCK_N);
        int stage = 0; uint32_t phase = 0, acc_phase = 0;
        for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
            int split, tap, mt, nt; decode(item, split, tap, mt, nt);
            const int kb0 = split * p.kblocks_per_split;
            const int kb1 = min(kb0 + p.kblocks_per_split, p.kblocks_total);
            mbar_wait(acc_empty, acc_phase ^ 1);
            tc_fence_after();
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t a_addr = smem_u32(smem + stage * L::kStageBytes);
                    const uint32_t b_addr = a_addr + L::kABytes;
#pragma unroll
                    for (int s = 0; s < NUM_SUB; ++s) {
#pragma unroll
                        for (int k = 0; k < kWgPix / 16; ++k) {
                            const uint64_t da = make_sw128_mnmajor_desc(a_addr + s * 2 * L::kBoxBytes + k * 2048, L::kBoxBytes);
                            const uint64_t db = make_sw128_mnmajor_desc(b_addr + k * 2048, L::kBoxBytes);
                            umma_bf16(tmem_base + s * BLOCK_N, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                        }
                    }
                    umma_commit(&empty_bar[stage]);
                    if (kb == kb1 - 1) umma_commit(acc_full);
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            if (kb1 <= kb0 && lane == 0) mbar_arrive(acc_full);      // empty split: publish (zeros are written below)
            acc_phase ^= 1;
        }
    } else {
        const int quad = warp & 3;
        uint32_t acc_phase = 0;
        for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
            int split, tap, mt, nt; decode(item, split, tap, mt, nt);
            const bool empty = (split * p.kblocks_per_split >= p.kblocks_total);
            mbar_wait(acc_full, acc_phase);
            tc_fence_after();
#pragma unroll 1
            for (int s = 0; s < NUM_SUB; ++s) {
                const int m = (mt * NUM_SUB + s) * 128 + quad * 32 + lane;
#pragma unroll 1
                for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
                    float v[32];
                    tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + s * BLOCK_N + c0, v);
                    const int n0 = nt * BLOCK_N + c0;
                    if (m < p.M && n0 < p.N) {
                        float* dst = p.partial + (((size_t)split * p.ntaps + tap) * p.M + m) * p.N + n0;
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            if (n0 + j < p.N)
                                *reinterpret_cast<float4*>(dst + j) = empty ? make_float4(0.f, 0.f, 0.f, 0.f)
                                                                            : make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty);
            acc_phase ^= 1;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, kTmemCols); }
}

// out[m][n][tap] (OIHW fp32, the reference's parameter layout) = sum over splits, fixed order. One thread per (m, n): the
// reads of every (split, tap) plane are coalesced across the warp and independent (splits * ntaps loads in flight), and the
// thread's ntaps results are one contiguous run of the OIHW tensor - a warp writes 32 * ntaps consecutive floats. (The first
// version walked (tap, m, n) and wrote single floats 4 * ntaps bytes apart: eight times the sector traffic; 23 us per
// launch at 2 lines per GPU, 66 launches per step.)
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float* __restrict__ partial, int splits, int ntaps, int M, int N, float* __restrict__ out,
                    int out_tap_major) {
    const long long plane = (long long)M * N;
    const long long total = (long long)ntaps * plane;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < plane; i += (long long)gridDim.x * blockDim.x) {
        float s[9];
#pragma unroll
        for (int t = 0; t < 9; ++t) s[t] = 0.f;
        for (int sp = 0; sp < splits; ++sp) {
            const float* src = partial + (size_t)sp * total + i;
#pragma unroll
            for (int t = 0; t < 9; ++t)
                if (t < ntaps) s[t] += __ldg(src + (size_t)t * plane);
        }
        if (out_tap_major) {
#pragma unroll
            for (int t = 0; t < 9; ++t)
                if (t < ntaps) out[(size_t)t * plane + i] = s[t];            // [tap][M][N]
        } else {
            float* dst = out + (size_t)i * ntaps;                            // [M][N][tap] == OIHW
#pragma unroll
            for (int t = 0; t < 9; ++t)
                if (t < ntaps) dst[t] = s[t];
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn wg_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}
// [B][H][W][C] bf16 with an explicit pixel pitch (elements) -> (C, W, H, B) map, box (64, 64, 1, 1)
static int make_map(CUtensorMap* m, const void* x, int B, int H, int W, int C, long long pitch) {
    EncodeTiledFn enc = wg_encode_fn();
    HCTR_CHECK(enc != nullptr, HCTR_ERR_CUDA, "cuTensorMapEncodeTiled not available");
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)pitch * 2, (cuuint64_t)W * pitch * 2, (cuuint64_t)H * W * pitch * 2};
    cuuint32_t box[4] = {64, (cuuint32_t)kWgPix, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    HCTR_CHECK(r == CUDA_SUCCESS, HCTR_ERR_CUDA, "cuTensorMapEncodeTiled(wgrad %dx%dx%dx%d pitch %lld) failed: %d", B, H, W, C, pitch, (int)r);
    return HCTR_OK;
}

template <int BLOCK_N, int NUM_SUB, int STAGES>
static int launch_wgrad(const CUtensorMap& g, const CUtensorMap& x, const WgradParams& p, cudaStream_t s) {
    using L = WgSmem<BLOCK_N, NUM_SUB, STAGES>;
    auto kern = wgrad_tcgen05_kernel<BLOCK_N, NUM_SUB, STAGES>;
    static PerDeviceOnce once;        // per instantiation
    int cfg_dev;
    if (once.need(cfg_dev)) {
        HCTR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
        once.mark(cfg_dev);
    }
    int sms = 0, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
    const int grid = p.total_items < sms ? p.total_items : sms;
    kern<<<grid, kWgThreads, L::kTotal, s>>>(g, x, p);
    HCTR_CUDA(cudaGetLastError());
    return HCTR_OK;
}

}  // namespace hctr

using namespace hctr;

extern "C" {

static int wgrad_plan(int B, int H, int W, int M, int N, int ntaps, WgradParams& p) {
    memset(&p, 0, sizeof(p));
    p.B = B; p.H = H; p.W = W;
    p.w_blocks = (W + kWgPix - 1) / kWgPix;
    const long long kb = (long long)B * H * p.w_blocks;
    HCTR_CHECK(kb < (1ll << 31), HCTR_ERR_INVALID, "wgrad: too many pixel blocks");
    p.kblocks_total = (int)kb;
    p.ntaps = ntaps; p.M = M; p.N = N;
    const int num_sub = M > 128 ? 2 : 1;
    const int block_n = N >= 256 ? 256 : (N >= 128 ? 128 : 64);
    p.m_tiles = (M + num_sub * 128 - 1) / (num_sub * 128);
    p.n_tiles = (N + block_n - 1) / block_n;
    const int base_items = ntaps * p.m_tiles * p.n_tiles;
    // Split K so that the work items fill whole waves of the 148 persistent CTAs with as few splits as possible: every split
    // writes and the reduction re-reads a full fp32 copy of dW (9.4 MB for 512x512x9), which at small batches costs more
    // than an idle tail (13 splits at 2 lines per GPU = 122 MB per layer and a 49 us reduction behind a 200 us GEMM).
    // cost(s) = waves(s) * (K blocks per item + per-item overhead) + s * (traffic of one partial copy, in K-block times)
    const int max_splits = (p.kblocks_total + 15) / 16 > 64 ? 64 : (p.kblocks_total + 15) / 16;
    int splits = 1;
    double best = 1e30;
    for (int s = 1; s <= (max_splits < 1 ? 1 : max_splits); ++s) {
        const int per = (p.kblocks_total + s - 1) / s;
        const int items = base_items * ((p.kblocks_total + per - 1) / per);
        const int waves = (items + 147) / 148;
        const double cost = (double)waves * (per + 6.0) + 5.0 * s;
        if (cost < best - 1e-9) { best = cost; splits = s; }
    }
    p.kblocks_per_split = (p.kblocks_total + splits - 1) / splits;
    p.splits = (p.kblocks_total + p.kblocks_per_split - 1) / p.kblocks_per_split;
    p.total_items = base_items * p.splits;
    return HCTR_OK;
}

long long hctr_wgrad_workspace_bytes(int B, int H, int W, int M, int N, int ntaps) {
    WgradParams p;
    if (wgrad_plan(B, H, W, M, N, ntaps, p) != HCTR_OK) return -1;
    return (long long)p.splits * ntaps * M * N * 4;
}

static int wgrad_run(const void* grad_out, long long g_pitch, const void* x, long long x_pitch, int x_H, float* dw,
                     int B, int H, int W, int M, int N, int ntaps, const int8_t* dh, const int8_t* dw_off,
                     int tap_major_out, void* workspace, long long workspace_bytes, void* stream) {
    HCTR_CHECK(grad_out && x && dw && workspace, HCTR_ERR_INVALID, "wgrad: null pointer");
    HCTR_CHECK(B > 0 && H > 0 && W > 0 && M > 0 && N > 0, HCTR_ERR_INVALID, "wgrad: bad shape");
    HCTR_CHECK(N % 64 == 0 && g_pitch % 8 == 0 && x_pitch % 8 == 0, HCTR_ERR_INVALID, "wgrad: N %% 64 and 16-byte pitches required");
    HCTR_CHECK(((uintptr_t)grad_out & 15) == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)workspace & 15) == 0,
               HCTR_ERR_INVALID, "wgrad: 16-byte alignment");
    WgradParams p;
    int rc = wgrad_plan(B, H, W, M, N, ntaps, p);
    if (rc) return rc;
    for (int t = 0; t < ntaps; ++t) { p.tap_dh[t] = dh[t]; p.tap_dw[t] = dw_off[t]; }
    const long long need = (long long)p.splits * ntaps * M * N * 4;
    HCTR_CHECK(workspace_bytes >= need, HCTR_ERR_INVALID, "wgrad: workspace too small (%lld < %lld)", workspace_bytes, need);
    p.partial = static_cast<float*>(workspace);
    CUtensorMap tg, tx;
    rc = make_map(&tg, grad_out, B, H, W, M, g_pitch);
    if (rc) return rc;
    rc = make_map(&tx, x, B, x_H, W, N, x_pitch);
    if (rc) return rc;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int num_sub = M > 128 ? 2 : 1;
    const int block_n = N >= 256 ? 256 : (N >= 128 ? 128 : 64);
    if (num_sub == 2) {
        if (block_n == 256) rc = launch_wgrad<256, 2, 3>(tg, tx, p, s);
        else if (block_n == 128) rc = launch_wgrad<128, 2, 4>(tg, tx, p, s);
        else rc = launch_wgrad<64, 2, 4>(tg, tx, p, s);
    } else {
        if (block_n == 256) rc = launch_wgrad<256, 1, 4>(tg, tx, p, s);
        else if (block_n == 128) rc = launch_wgrad<128, 1, 4>(tg, tx, p, s);
        else rc = launch_wgrad<64, 1, 4>(tg, tx, p, s);
    }
    if (rc) return rc;
    HCTR_CHECK(ntaps <= 9, HCTR_ERR_INVALID, "wgrad: at most 9 taps");
    const long long plane = (long long)M * N;
    int blocks = (int)((plane + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    wgrad_reduce_kernel<<<blocks, 256, 0, s>>>(p.partial, p.splits, ntaps, M, N, dw, tap_major_out);
    HCTR_CUDA(cudaGetLastError());
    return HCTR_OK;
}

int hctr_conv_wgrad(const void* dz, const void* x, float* dw, int B, int H, int W, int Cout, int Cin, int ksize,
                    void* workspace, long long workspace_bytes, void* stream) {
    HCTR_CHECK(ksize == 1 || ksize == 3, HCTR_ERR_INVALID, "wgrad: ksize must be 1 or 3");
    int8_t dh[9] = {0}, dwo[9] = {0};
    if (ksize == 3)
        for (int kh = 0; kh < 3; ++kh)
            for (int kw = 0; kw < 3; ++kw) { dh[kh * 3 + kw] = (int8_t)(kh - 1); dwo[kh * 3 + kw] = (int8_t)(kw - 1); }
    // dw: fp32 [Cout][Cin][k][k] (OIHW, the reference parameter layout)
    return wgrad_run(dz, Cout, x, Cin, H, dw, B, H, W, Cout, Cin, ksize * ksize, dh, dwo, 0, workspace, workspace_bytes, stream);
}

int hctr_linear_wgrad(const void* dlogits, long long pitch, const void* feat, float* dw, int B, int Hf, int W, int Cf,
                      int num_classes, void* workspace, long long workspace_bytes, void* stream) {
    HCTR_CHECK(Hf >= 1 && Hf <= 9, HCTR_ERR_INVALID, "linear_wgrad: feature rows must be in [1,9]");
    int8_t dh[9] = {0}, dwo[9] = {0};
    for (int h = 0; h < Hf; ++h) dh[h] = (int8_t)h;
    // dw: fp32 [num_classes][Cf][Hf] == the reference's linear.weight [num_classes][c*Hf + h]
    return wgrad_run(dlogits, pitch, feat, Cf, Hf, dw, B, 1, W, num_classes, Cf, Hf, dh, dwo, 0, workspace, workspace_bytes, stream);
}

long long hctr_linear_wgrad_workspace_bytes(int B, int Hf, int W, int Cf, int num_classes) {
    return hctr_wgrad_workspace_bytes(B, 1, W, num_classes, Cf, Hf);
}

}  // extern "C"
