// C-ABI entry points for the tensor-core path: conv3x3 / conv1x1 (+BN+ReLU+pool) and the
// column classifier. Declared in include/hctr_b200.h.
#include <atomic>
#include <cstdarg>
#include <cstring>
#include <mutex>

#include <cstdlib>

#include "igemm2_tcgen05.cuh"
#include "../../include/hctr_b200.h"
#include "../../include/hctr_b200_testing.h"

namespace hctr {

// ---------------------------------------------------------------- error plumbing
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_cuda(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return HCTR_OK;
    set_error("CUDA error %s (%d) in %s", cudaGetErrorString(e), static_cast<int>(e), what);
    return HCTR_ERR_CUDA;
}

// ---------------------------------------------------------------- tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

// NHWC bf16 activation [B,H,W,C] -> 4-D map (C, W, H, B), box (64, 128, 1, 1), 128B swizzle, zero OOB fill.
static int make_act_map(CUtensorMap* m, const void* x, int B, int H, int W, int C, long long pitch = 0, int box_w = kTileM) {
    EncodeTiledFn enc = get_encode_fn();
    HCTR_CHECK(enc != nullptr, HCTR_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
    if (pitch <= 0) pitch = C;                    // elements between consecutive pixels
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)pitch * 2, (cuuint64_t)W * pitch * 2, (cuuint64_t)H * W * pitch * 2};
    cuuint32_t box[4] = {(cuuint32_t)kBlockK, (cuuint32_t)box_w, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    HCTR_CHECK(r == CUDA_SUCCESS, HCTR_ERR_CUDA, "cuTensorMapEncodeTiled(activation %dx%dx%dx%d) failed: %d", B, H, W, C,
               (int)r);
    return HCTR_OK;
}

// Packed weights [N][K] bf16 (K-major) -> 2-D map (K, N), box (64, block_n).
static int make_weight_map(CUtensorMap* m, const void* w, int N, int K, int block_n, long long pitch = 0) {
    EncodeTiledFn enc = get_encode_fn();
    HCTR_CHECK(enc != nullptr, HCTR_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
    if (pitch <= 0) pitch = K;
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)N};
    cuuint64_t strides[1] = {(cuuint64_t)pitch * 2};
    cuuint32_t box[2] = {(cuuint32_t)kBlockK, (cuuint32_t)block_n};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    HCTR_CHECK(r == CUDA_SUCCESS, HCTR_ERR_CUDA, "cuTensorMapEncodeTiled(weights %dx%d) failed: %d", N, K, (int)r);
    return HCTR_OK;
}

static int sm_count() {
    static PerDeviceOnce once;
    int dev;
    if (once.need(dev)) {
        int n = 0;
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
        once.mark(dev, n);
        return n;
    }
    return once.get(dev);
}

template <int BLOCK_N, int NUM_SUB, int STAGES, int ACC_STAGES, int EPI, int KWF = 0>
static int launch_igemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const IgemmParams& p, cudaStream_t stream) {
    using L = IgemmSmem<BLOCK_N, NUM_SUB, STAGES, KWF>;
    auto kern = igemm_tcgen05_kernel<BLOCK_N, NUM_SUB, STAGES, ACC_STAGES, EPI, KWF>;
    static PerDeviceOnce once;        // per instantiation
    int dev;
    if (once.need(dev)) {
        HCTR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
        once.mark(dev);
    }
    int grid = p.total_tiles < sm_count() ? p.total_tiles : sm_count();
    IgemmParams q = p;
    // whole (line, span) columns per CTA when a column has several tiles and the columns balance (<= 4 % idle tail);
    // HCTR_IGEMM_COLS=0 turns it off (A/B measurements)
    static const bool cols_off = getenv("HCTR_IGEMM_COLS") && getenv("HCTR_IGEMM_COLS")[0] == '0';
    const long long ncols = (long long)p.B * p.w_tiles;
    const long long padded = (ncols + grid - 1) / grid * grid;
    q.col_mode = (!cols_off && EPI == EPI_CONV && p.h_tiles > 1 && padded * 100 <= ncols * 104) ? 1 : 0;
    kern<<<grid, kIgemmThreads, L::kTotal, stream>>>(tmA, tmB, q);
    HCTR_CUDA(cudaGetLastError());
    return HCTR_OK;
}

// kw-fused activation slab (one 136-pixel box per (kh, 64-channel chunk) serves the three kw taps; the UMMA descriptor
// start address is shifted by whole 128-byte rows, the swizzle XOR comes from the absolute shared-memory address).
// 1 = on (default), 0 = one TMA box per tap. Process-wide switches for A/B measurements and the variant tests only:
// HCTR_IGEMM_KWF / HCTR_IGEMM_PAIR in the environment, or the test hooks of include/hctr_b200_testing.h.
static std::atomic<int> g_kwf_mode{-1};
static int kwf_mode() {
    int m = g_kwf_mode.load(std::memory_order_relaxed);
    if (m < 0) {
        const char* e = getenv("HCTR_IGEMM_KWF");
        m = (e && e[0] == '0') ? 0 : 1;
        g_kwf_mode.store(m, std::memory_order_relaxed);
    }
    return m;
}

// CTA-pair kernel (cta_group::2): used for the wide (Cout % 256 == 0), un-pooled convolutions. HCTR_IGEMM_PAIR=0 falls
// back to the single-CTA kernel (A/B measurements).
static std::atomic<int> g_pair_mode{-1};
static bool pair_enabled() {
    int m = g_pair_mode.load(std::memory_order_relaxed);
    if (m < 0) {
        const char* e = getenv("HCTR_IGEMM_PAIR");
        m = (e && e[0] == '0') ? 0 : 1;
        g_pair_mode.store(m, std::memory_order_relaxed);
    }
    return m != 0;
}
static bool test_hooks_enabled() {
    const char* e = getenv("HCTR_TEST_HOOKS");
    return e && e[0] == '1';
}
// (every layer of the model with Cout % 256 == 0 has Cin % 128 == 0, so the K blocks pair up)
// Pooled layers on the pair kernel: the two rows of a (2,1) window live in the two CTAs of a pair, which combine them
// with red.global.max.v4.bf16x2 into a zero-filled output (needs ReLU: all operands >= 0). HCTR_IGEMM_PAIR_POOL=0 keeps
// them on the single-CTA kernel (A/B measurements).
static bool pair_pool_enabled() {
    static int mode = -1;
    if (mode < 0) {
        const char* e = getenv("HCTR_IGEMM_PAIR_POOL");
        mode = (e && e[0] == '0') ? 0 : 1;
    }
    return mode == 1;
}
static bool use_pair(int H, int Cin, int Cout, int ksize, int pool, int relu = 0, bool plain = false) {
    // (measured: at Cout = 128 the pair kernel is slower than the single-CTA slab kernel, with or without the slab -
    //  722 / 763 vs 990 TFLOP/s on 128->128: K = 1152 gives 4 us tiles and the cluster-wide accumulator hand-over
    //  per tile dominates)
    const bool slab = ksize == 3 && kwf_mode() != 0;         // one stage per (kh, chunk); otherwise K blocks go in pairs
    if (pool && !(relu && plain && pair_pool_enabled())) return false;
    return pair_enabled() && Cout % 256 == 0 && H % 2 == 0 && (slab || (ksize * ksize * (Cin / 64)) % kPairKSub == 0);
}

template <int BLOCK_N, int STAGES, int KWF, int ADD>
static int launch_igemm_pair(const CUtensorMap& tmA, const CUtensorMap& tmB, const IgemmParams& p, cudaStream_t stream) {
    using L = PairSmem<BLOCK_N, STAGES, KWF>;
    auto kern = igemm_pair_kernel<BLOCK_N, STAGES, KWF, ADD>;
    static PerDeviceOnce once;        // per instantiation; value = max active clusters on that device
    int dev;
    int max_clusters = 0;
    if (!once.need(dev)) {
        max_clusters = once.get(dev);
    } else {
        HCTR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * sm_count());
        cfg.blockDim = dim3(kIgemmThreads);
        cfg.dynamicSmemBytes = L::kTotal;
        cudaLaunchAttribute attr;
        attr.id = cudaLaunchAttributeClusterDimension;
        attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
        cfg.attrs = &attr; cfg.numAttrs = 1;
        int n = 0;
        HCTR_CUDA(cudaOccupancyMaxActiveClusters(&n, kern, &cfg));
        HCTR_CHECK(n > 0, HCTR_ERR_CUDA, "conv: the CTA-pair kernel does not fit this device");
        max_clusters = n;
        once.mark(dev, n);
        if (getenv("HCTR_DEBUG")) fprintf(stderr, "hctr_b200: igemm_pair_kernel<%d> max active clusters = %d (SMs %d)\n", BLOCK_N, n, sm_count());
    }
    const int pairs = p.total_tiles < max_clusters ? p.total_tiles : max_clusters;
    IgemmParams q = p;
    // whole columns per pair when they balance (<= 4 % idle tail); HCTR_PAIR_COLS=0 turns it off
    static const bool cols_off = getenv("HCTR_PAIR_COLS") && getenv("HCTR_PAIR_COLS")[0] == '0';
    const long long ncols = (long long)p.B * p.w_tiles;
    const long long padded = (ncols + pairs - 1) / pairs * pairs;
    q.col_mode = (!cols_off && padded * 100 <= ncols * 104) ? 1 : 0;
    kern<<<2 * pairs, kIgemmThreads, L::kTotal, stream>>>(tmA, tmB, q);
    HCTR_CUDA(cudaGetLastError());
    return HCTR_OK;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// log-sum-exp fix-up of the classifier's per-(row, half-tile) softmax partials (fixed order)
__global__ void lse_combine_kernel(const float2* __restrict__ partial, long long rows, int slots, float* __restrict__ lse) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const float2* p = partial + r * slots;
    float m = -3.0e38f;
    for (int i = 0; i < slots; ++i) m = fmaxf(m, p[i].x);
    float s = 0.f;
    for (int i = 0; i < slots; ++i) s += p[i].y * __expf(p[i].x - m);
    lse[r] = m + logf(s);
}

// arg-max fix-up of the classifier's per-(row, half-tile) partials, in class order: numpy.argmax over the row
__global__ void argmax_combine_kernel(const int2* __restrict__ partial, long long rows, int slots, int32_t* __restrict__ out) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const int2* p = partial + r * slots;
    float bv = 0.f;
    int bi = -1;
    for (int i = 0; i < slots; ++i) {
        const int2 q = p[i];
        if (q.y < 0) continue;                                  // a half-tile beyond the last class
        const float x = __int_as_float(q.x);
        if (bi < 0 || x > bv || (x != x && bv == bv)) { bv = x; bi = q.y; }
    }
    out[r] = bi < 0 ? 0 : bi;
}

}  // namespace hctr

using namespace hctr;

extern "C" {

const char* hctr_last_error(void) { return g_err; }

int hctr_abi_version(void) { return HCTR_ABI_VERSION; }

int hctr_device_supported(int device) {
    cudaDeviceProp prop;
    HCTR_CUDA(cudaGetDeviceProperties(&prop, device));
    HCTR_CHECK(prop.major == 10, HCTR_ERR_UNSUPPORTED,
               "hctr_b200 kernels need an sm_100a device (B200); device %d is sm_%d%d", device, prop.major, prop.minor);
    return HCTR_OK;
}

static int conv_launch(const void* x, const void* w_packed, const float* scale, const float* shift, const void* add,
                       void* y, int B, int H, int W, int Cin, int Cout, int ksize, int relu, int pool, int flip,
                       void* stream, float* se_partial = nullptr, int sum_stored = 0, const float* gate = nullptr,
                       float* sq_partial = nullptr) {
    HCTR_CHECK(x && w_packed && scale && shift && y, HCTR_ERR_INVALID, "conv: null pointer");
    HCTR_CHECK(ksize == 1 || ksize == 3, HCTR_ERR_INVALID, "conv: ksize must be 1 or 3 (got %d)", ksize);
    HCTR_CHECK(B > 0 && H > 0 && W > 0, HCTR_ERR_INVALID, "conv: empty tensor %dx%dx%d", B, H, W);
    HCTR_CHECK(Cin % 64 == 0 && Cin >= 64, HCTR_ERR_INVALID, "conv: Cin must be a multiple of 64 (got %d)", Cin);
    HCTR_CHECK(Cout == 64 || Cout == 128 || Cout % 256 == 0, HCTR_ERR_INVALID,
               "conv: Cout must be 64, 128 or a multiple of 256 (got %d)", Cout);
    HCTR_CHECK(!pool || (H % 2 == 0), HCTR_ERR_INVALID, "conv: (2,1) pooling needs an even height (got %d)", H);
    HCTR_CHECK(aligned16(x) && aligned16(w_packed) && aligned16(y) && aligned16(scale) && aligned16(shift),
               HCTR_ERR_INVALID, "conv: pointers must be 16-byte aligned");

    HCTR_CHECK(!(se_partial && sum_stored && pool), HCTR_ERR_INVALID, "conv: sums of the stored tensor are not taken on pooled layers");
    IgemmParams p;
    memset(&p, 0, sizeof(p));
    p.B = B; p.H = H; p.W = W;
    p.cin_chunks = Cin / 64;
    if (ksize == 3) {
        p.ntaps = 9;
        for (int kh = 0; kh < 3; ++kh)
            for (int kw = 0; kw < 3; ++kw) {
                // flip: transposed convolution for the data gradient reads dz at (h - (kh-1), w - (kw-1))
                p.tap_dh[kh * 3 + kw] = (int8_t)(flip ? 1 - kh : kh - 1);
                p.tap_dw[kh * 3 + kw] = (int8_t)(flip ? 1 - kw : kw - 1);
            }
    } else {
        p.ntaps = 1;
    }
    p.sub_dh = 1; p.sub_dw = 0;
    p.add = add;
    p.se_partial = se_partial;
    p.sum_stored = sum_stored;
    p.sq_partial = sq_partial;
    p.gate = gate;
    p.N = Cout;
    p.w_tiles = (W + kTileM - 1) / kTileM;
    p.h_tiles = (H + 1) / 2;
    p.scale = scale; p.shift = shift; p.out = y;
    p.relu = relu; p.pool = pool;
    p.out_H = pool ? H / 2 : H;

    const int block_n = Cout >= 256 ? 256 : Cout;
    p.n_tiles = Cout / block_n;
    const long long total = (long long)B * p.h_tiles * p.w_tiles * p.n_tiles;
    HCTR_CHECK(total < (1ll << 31), HCTR_ERR_INVALID, "conv: too many tiles");
    p.total_tiles = (int)total;

    if (use_pair(H, Cin, Cout, ksize, pool, relu, !add && !se_partial && !gate)) {
        // one tile = rows (2*h_tile, 2*h_tile+1) x 128 pixels x 256 channels on a CTA pair
        const bool slab = ksize == 3 && kwf_mode() != 0;          // kw-fused activation slab (HCTR_IGEMM_KWF=0 turns it off)
        CUtensorMap tmA, tmB;
        int rc = make_act_map(&tmA, x, B, H, W, Cin, 0, slab ? kSlabPix : kTileM);
        if (rc) return rc;
        rc = make_weight_map(&tmB, w_packed, Cout, p.ntaps * Cin, block_n / 2);
        if (rc) return rc;
        cudaStream_t cs = static_cast<cudaStream_t>(stream);
        if (pool)        // identity of the max reduction over ReLU outputs
            HCTR_CUDA(cudaMemsetAsync(y, 0, (size_t)B * (H / 2) * W * Cout * sizeof(__nv_bfloat16), cs));
        if (add) {
            HCTR_CHECK(se_partial == nullptr, HCTR_ERR_INVALID, "conv: channel sums and a residual cannot be combined");
            return slab ? launch_igemm_pair<256, 3, 1, 1>(tmA, tmB, p, cs) : launch_igemm_pair<256, 3, 0, 1>(tmA, tmB, p, cs);
        }
        return slab ? launch_igemm_pair<256, 3, 1, 0>(tmA, tmB, p, cs) : launch_igemm_pair<256, 3, 0, 0>(tmA, tmB, p, cs);
    }

    // thin layers (Cout <= 128) are bound by the L2->SMEM re-reads of the activation tile: fuse the three kw taps
    const int kwf = (ksize == 3 && block_n <= 128) ? kwf_mode() : 0;
    CUtensorMap tmA, tmB;
    int rc = make_act_map(&tmA, x, B, H, W, Cin, 0, kwf ? kSlabPix : kTileM);
    if (rc) return rc;
    rc = make_weight_map(&tmB, w_packed, Cout, p.ntaps * Cin, block_n);
    if (rc) return rc;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (kwf) {
        if (block_n == 64) return launch_igemm<64, 2, 3, 2, EPI_CONV, 1>(tmA, tmB, p, s);
        return launch_igemm<128, 2, 2, 2, EPI_CONV, 1>(tmA, tmB, p, s);
    }
    switch (block_n) {
        case 64:  return launch_igemm<64, 2, 4, 2, EPI_CONV>(tmA, tmB, p, s);
        case 128: return launch_igemm<128, 2, 4, 2, EPI_CONV>(tmA, tmB, p, s);
        default:  return launch_igemm<256, 2, 3, 1, EPI_CONV>(tmA, tmB, p, s);
    }
}

// Test hook (include/hctr_b200_testing.h): selects the kernel variant of the NEXT conv launches of this process.
// Refused unless HCTR_TEST_HOOKS=1 is in the environment, so a product process cannot change variants by accident.
int hctr_testing_set_conv_variant(int kw_fused_slab, int cta_pairs) {
    HCTR_CHECK(test_hooks_enabled(), HCTR_ERR_UNSUPPORTED, "test hooks are disabled (set HCTR_TEST_HOOKS=1)");
    g_kwf_mode.store(kw_fused_slab ? 1 : 0, std::memory_order_relaxed);
    g_pair_mode.store(cta_pairs ? 1 : 0, std::memory_order_relaxed);
    return HCTR_OK;
}

int hctr_conv_bn_act_fwd(const void* x, const void* w_packed, const float* scale, const float* shift, void* y, int B,
                         int H, int W, int Cin, int Cout, int ksize, int relu, int pool, void* stream) {
    return conv_launch(x, w_packed, scale, shift, nullptr, y, B, H, W, Cin, Cout, ksize, relu, pool, 0, stream);
}

int hctr_conv_se_slices(int H, int W, int Cout) {
    // one slot per (tile row, 128-px span, epilogue warp quarter); the CTA-pair kernel keeps the two rows of a tile apart
    const int rows = use_pair(H, Cout, Cout, 3, 0) ? H : (H + 1) / 2;        // BasicBlock conv2: 3x3, Cin == Cout
    return rows * ((W + kTileM - 1) / kTileM) * 4;
}

int hctr_conv_bn_se_fwd(const void* x, const void* w_packed, const float* scale, const float* shift, void* y,
                        float* se_partial, int B, int H, int W, int Cin, int Cout, int ksize, void* stream) {
    HCTR_CHECK(se_partial != nullptr, HCTR_ERR_INVALID, "conv_bn_se: null partial buffer");
    HCTR_CHECK(Cin == Cout && ksize == 3, HCTR_ERR_INVALID, "conv_bn_se: BasicBlock conv2 is 3x3 with Cin == Cout (got %d -> %d, k=%d)", Cin, Cout, ksize);
    return conv_launch(x, w_packed, scale, shift, nullptr, y, B, H, W, Cin, Cout, ksize, 0, 0, 0, stream, se_partial);
}

int hctr_conv_sum_slices(int H, int W, int Cin, int Cout, int ksize) {
    // sums of the values as stored: one slot per (row, 128-px span, epilogue warp quarter) in either kernel; the single-CTA
    // kernel's tiles are row pairs, so an odd H has one more (zero) row of slots
    const int rows = use_pair(H, Cin, Cout, ksize, 0) ? H : (H + 1) / 2 * 2;
    return rows * ((W + kTileM - 1) / kTileM) * 4;
}

int hctr_conv_bn_act_sum_fwd(const void* x, const void* w_packed, const float* scale, const float* shift, void* y,
                             float* partial, int B, int H, int W, int Cin, int Cout, int ksize, int relu, void* stream) {
    HCTR_CHECK(partial != nullptr, HCTR_ERR_INVALID, "conv_bn_act_sum: null partial buffer");
    return conv_launch(x, w_packed, scale, shift, nullptr, y, B, H, W, Cin, Cout, ksize, relu, 0, 0, stream, partial, 1);
}

int hctr_conv_stats_fwd(const void* x, const void* w_packed, const float* scale, const float* shift, void* y, float* psum,
                        float* psq, int B, int H, int W, int Cin, int Cout, int ksize, void* stream) {
    HCTR_CHECK(psum != nullptr && psq != nullptr, HCTR_ERR_INVALID, "conv_stats: null partial buffer");
    return conv_launch(x, w_packed, scale, shift, nullptr, y, B, H, W, Cin, Cout, ksize, 0, 0, 0, stream, psum, 1, nullptr, psq);
}

int hctr_conv_bn_gate_res_fwd(const void* x, const void* w_packed, const float* scale, const float* shift, const float* gate,
                              const void* residual, void* y, int B, int H, int W, int Cin, int Cout, int ksize, int relu,
                              void* stream) {
    HCTR_CHECK(gate && residual, HCTR_ERR_INVALID, "conv_bn_gate_res: null gate / residual");
    HCTR_CHECK(aligned16(gate) && aligned16(residual), HCTR_ERR_INVALID, "conv_bn_gate_res: gate and residual must be 16-byte aligned");
    return conv_launch(x, w_packed, scale, shift, residual, y, B, H, W, Cin, Cout, ksize, relu, 0, 0, stream, nullptr, 0, gate);
}

int hctr_conv_dgrad(const void* dz, const void* w_packed_t, const float* ones, const float* zeros, const void* add,
                    void* dx, int B, int H, int W, int Cout, int Cin, int ksize, void* stream) {
    HCTR_CHECK(!add || aligned16(add), HCTR_ERR_INVALID, "dgrad: add tensor must be 16-byte aligned");
    // the data gradient is the same implicit GEMM with the roles of Cin/Cout swapped and mirrored taps
    return conv_launch(dz, w_packed_t, ones, zeros, add, dx, B, H, W, Cout, Cin, ksize, 0, 0, 1, stream);
}

int hctr_classifier_dgrad(const void* dlogits, long long pitch, const void* w_t, const float* ones, const float* zeros,
                          void* dfeat, int B, int Hf, int W, int Cf, int num_classes, void* stream) {
    HCTR_CHECK(dlogits && w_t && ones && zeros && dfeat, HCTR_ERR_INVALID, "classifier_dgrad: null pointer");
    HCTR_CHECK(pitch % 8 == 0 && pitch >= num_classes, HCTR_ERR_INVALID, "classifier_dgrad: pitch must be a multiple of 8");
    HCTR_CHECK(Cf == 64 || Cf == 128 || Cf % 256 == 0, HCTR_ERR_INVALID, "classifier_dgrad: bad Cf %d", Cf);
    HCTR_CHECK(aligned16(dlogits) && aligned16(w_t) && aligned16(dfeat), HCTR_ERR_INVALID, "classifier_dgrad: alignment");
    // dfeat[b, h, w, c] = sum_n dlogits[b, w, n] * Wt[h*Cf + c, n]   (bf16 in, fp32 accumulate, bf16 out)
    const int kchunks = (num_classes + 63) / 64;
    const int block_n = Cf >= 256 ? 256 : Cf;
    for (int h = 0; h < Hf; ++h) {
        IgemmParams p;
        memset(&p, 0, sizeof(p));
        p.B = B; p.H = 1; p.W = W;
        p.cin_chunks = kchunks; p.ntaps = 1;
        p.sub_dh = 0; p.sub_dw = 1;
        p.N = Cf;
        p.w_tiles = (W + 2 * kTileM - 1) / (2 * kTileM);
        p.h_tiles = 1;
        p.n_tiles = Cf / block_n;
        p.scale = ones; p.shift = zeros;
        p.out = static_cast<__nv_bfloat16*>(dfeat) + (size_t)h * W * Cf;      // row h of every line; line pitch = Hf*W*Cf
        p.out_H = 1;
        // out index = ((b*out_H + 0)*W + w)*N + n  -> we need a line pitch of Hf*W*Cf: fold Hf into W of the output
        // by treating the output as [B][Hf*W][Cf] with w offset h*W (pointer offset above) and out_H = 1, W_out = Hf*W.
        p.total_tiles = B * p.w_tiles * p.n_tiles;
        p.out_line_pitch = (long long)Hf * W * Cf;
        CUtensorMap tmA, tmB;
        int rc = make_act_map(&tmA, dlogits, B, 1, W, num_classes, pitch);
        if (rc) return rc;
        rc = make_weight_map(&tmB, static_cast<const __nv_bfloat16*>(w_t) + (size_t)h * Cf * pitch, Cf, num_classes, block_n, pitch);
        if (rc) return rc;
        cudaStream_t s = static_cast<cudaStream_t>(stream);
        switch (block_n) {
            case 64:  rc = launch_igemm<64, 2, 4, 2, EPI_CONV>(tmA, tmB, p, s); break;
            case 128: rc = launch_igemm<128, 2, 4, 2, EPI_CONV>(tmA, tmB, p, s); break;
            default:  rc = launch_igemm<256, 2, 3, 1, EPI_CONV>(tmA, tmB, p, s); break;
        }
        if (rc) return rc;
    }
    return HCTR_OK;
}

// Column tile of the classifier GEMM: 256 (one TMEM accumulator stage: the epilogue - 1.9 GB of logits at B=64 - is not
// overlapped with the next tile's main loop) or 128 (two stages; twice the activation re-reads). HCTR_CLS_N picks (A/B).
static int cls_block_n() {
    static int n = 0;
    if (n == 0) {
        const char* e = getenv("HCTR_CLS_N");
        n = (e && atoi(e) == 128) ? 128 : 256;
    }
    return n;
}

static int classifier_launch(const void* feat, const void* w_packed, const float* bias, void* logits, int out_dtype,
                             long long out_pitch, int B, int Hf, int W, int Cf, int num_classes, float2* lse_partial,
                             int2* argmax_partial, void* stream) {
    HCTR_CHECK(feat && w_packed && bias && (logits || (argmax_partial && !lse_partial)), HCTR_ERR_INVALID, "classifier: null pointer");
    HCTR_CHECK(B > 0 && W > 0 && Hf > 0 && Hf <= kMaxTaps, HCTR_ERR_INVALID, "classifier: bad shape B=%d Hf=%d W=%d", B, Hf, W);
    HCTR_CHECK(Cf % 64 == 0 && Cf >= 64, HCTR_ERR_INVALID, "classifier: feature channels must be a multiple of 64 (got %d)", Cf);
    HCTR_CHECK(num_classes > 0 && out_pitch >= num_classes, HCTR_ERR_INVALID, "classifier: pitch %lld < classes %d", out_pitch, num_classes);
    HCTR_CHECK(out_dtype == HCTR_F32 || out_dtype == HCTR_BF16, HCTR_ERR_INVALID, "classifier: bad out dtype %d", out_dtype);
    HCTR_CHECK(aligned16(feat) && aligned16(w_packed) && aligned16(bias), HCTR_ERR_INVALID, "classifier: pointers must be 16-byte aligned");

    IgemmParams p;
    memset(&p, 0, sizeof(p));
    p.B = B; p.H = Hf; p.W = W;
    p.cin_chunks = Cf / 64;
    p.ntaps = Hf;                               // one "tap" per feature row: K index = h*Cf + c
    for (int h = 0; h < Hf; ++h) { p.tap_dh[h] = (int8_t)h; p.tap_dw[h] = 0; }
    p.sub_dh = 0; p.sub_dw = 1;
    p.N = num_classes;
    p.w_tiles = (W + 2 * kTileM - 1) / (2 * kTileM);
    p.h_tiles = 1;
    const int block_n = cls_block_n();
    p.n_tiles = (num_classes + block_n - 1) / block_n;
    p.shift = bias; p.out = logits;
    p.out_H = 1;
    p.out_dtype = out_dtype; p.out_pitch = out_pitch;
    p.lse_partial = lse_partial;
    p.argmax_partial = argmax_partial;
    const long long total = (long long)B * p.w_tiles * p.n_tiles;
    HCTR_CHECK(total < (1ll << 31), HCTR_ERR_INVALID, "classifier: too many tiles");
    p.total_tiles = (int)total;

    CUtensorMap tmA, tmB;
    int rc = make_act_map(&tmA, feat, B, Hf, W, Cf);
    if (rc) return rc;
    rc = make_weight_map(&tmB, w_packed, num_classes, Hf * Cf, block_n);
    if (rc) return rc;
    if (block_n == 128) return launch_igemm<128, 2, 4, 2, EPI_LINEAR>(tmA, tmB, p, static_cast<cudaStream_t>(stream));
    return launch_igemm<256, 2, 3, 1, EPI_LINEAR>(tmA, tmB, p, static_cast<cudaStream_t>(stream));
}


int hctr_classifier_fwd(const void* feat, const void* w_packed, const float* bias, void* logits, int out_dtype,
                        long long out_pitch, int B, int Hf, int W, int Cf, int num_classes, void* stream) {
    return classifier_launch(feat, w_packed, bias, logits, out_dtype, out_pitch, B, Hf, W, Cf, num_classes, nullptr, nullptr, stream);
}

long long hctr_classifier_lse_workspace_bytes(int B, int W, int num_classes) {
    return (long long)B * W * ((num_classes + 127) / 128) * 2 * (long long)sizeof(float2);      // sized for the 128-column tile
}

int hctr_classifier_lse_fwd(const void* feat, const void* w_packed, const float* bias, void* logits, int out_dtype,
                            long long out_pitch, int B, int Hf, int W, int Cf, int num_classes, float* row_lse,
                            void* workspace, long long workspace_bytes, void* stream) {
    HCTR_CHECK(row_lse && workspace, HCTR_ERR_INVALID, "classifier_lse: null pointer");
    HCTR_CHECK(workspace_bytes >= hctr_classifier_lse_workspace_bytes(B, W, num_classes), HCTR_ERR_INVALID, "classifier_lse: workspace too small");
    HCTR_CHECK(aligned16(workspace), HCTR_ERR_INVALID, "classifier_lse: workspace alignment");
    int rc = classifier_launch(feat, w_packed, bias, logits, out_dtype, out_pitch, B, Hf, W, Cf, num_classes,
                               static_cast<float2*>(workspace), nullptr, stream);
    if (rc) return rc;
    const long long rows = (long long)B * W;
    const int slots = ((num_classes + cls_block_n() - 1) / cls_block_n()) * 2;
    lse_combine_kernel<<<(int)((rows + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const float2*>(workspace), rows, slots, row_lse);
    HCTR_CUDA(cudaGetLastError());
    return HCTR_OK;
}

long long hctr_classifier_greedy_workspace_bytes(int B, int W, int num_classes) {
    return (long long)B * W * ((num_classes + 127) / 128) * 2 * (long long)sizeof(int2);        // sized for the 128-column tile
}

int hctr_classifier_greedy_fwd(const void* feat, const void* w_packed, const float* bias, void* logits, int out_dtype,
                               long long out_pitch, int B, int Hf, int W, int Cf, int num_classes, int32_t* argmax_bt,
                               int32_t* out_idx, int32_t* out_len, void* workspace, long long workspace_bytes, void* stream) {
    HCTR_CHECK(argmax_bt && out_idx && out_len && workspace, HCTR_ERR_INVALID, "classifier_greedy: null pointer");
    HCTR_CHECK(workspace_bytes >= hctr_classifier_greedy_workspace_bytes(B, W, num_classes), HCTR_ERR_INVALID,
               "classifier_greedy: workspace too small");
    HCTR_CHECK(aligned16(workspace), HCTR_ERR_INVALID, "classifier_greedy: workspace alignment");
    int rc = classifier_launch(feat, w_packed, bias, logits, out_dtype, out_pitch, B, Hf, W, Cf, num_classes, nullptr,
                               static_cast<int2*>(workspace), stream);
    if (rc) return rc;
    const long long rows = (long long)B * W;
    const int slots = ((num_classes + cls_block_n() - 1) / cls_block_n()) * 2;
    argmax_combine_kernel<<<(int)((rows + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const int2*>(workspace), rows, slots, argmax_bt);
    HCTR_CUDA(cudaGetLastError());
    return hctr_ctc_collapse(argmax_bt, W, B, num_classes, out_idx, out_len, stream);
}

}  // extern "C"
