// Small training-side kernels: column sums (classifier bias gradient), the stem's weight gradient, and the fused
// optimizer tail (global grad-norm clip + SGD with momentum / weight decay; reference main.py:210-213,430-438).
#include "common.cuh"
#include "../../include/hctr_b200.h"

namespace hctr {

// ---------------------------------------------------------------- column sums of a bf16 [rows][pitch] matrix
constexpr int kColRows = 512;     // rows per slice
__global__ void __launch_bounds__(256)
colsum_partial_kernel(const __nv_bfloat16* __restrict__ x, long long rows, int cols, long long pitch, float* __restrict__ partial) {
    const int c = blockIdx.x * 256 + threadIdx.x;
    const long long r0 = (long long)blockIdx.y * kColRows;
    const long long r1 = r0 + kColRows < rows ? r0 + kColRows : rows;
    if (c >= cols) return;
    float s = 0.f;
    for (long long r = r0; r < r1; ++r) s += __bfloat162float(x[r * pitch + c]);
    partial[(size_t)blockIdx.y * cols + c] = s;
}
__global__ void colsum_final_kernel(const float* __restrict__ partial, int slices, int cols, float* __restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    float s = 0.f;
    for (int i = 0; i < slices; ++i) s += partial[(size_t)i * cols + c];
    out[c] = s;
}

// out[g][c] = sum over the slices i = g, g + groups, ... (fixed order): first stage of a two-stage column reduction
__global__ void colsum_strided_kernel(const float* __restrict__ partial, int slices, int cols, int groups, float* __restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int g = blockIdx.y;
    if (c >= cols) return;
    float s = 0.f;
    for (int i = g; i < slices; i += groups) s += partial[(size_t)i * cols + c];
    out[(size_t)g * cols + c] = s;
}

// ---------------------------------------------------------------- stem weight gradient (Cin = 1)
// dW[co][tap] = sum_pix dz[pix][co] * x[pix + tap]; dz is NHWC bf16 [B][H][W][64], x fp32 [B][1][H][W].
// One block = one run of up to kStemRun pixels of one image row: the three input rows it touches sit in shared memory
// (zero-padded), 32-bit indexing, no division in the pixel loop (the first version walked 4096 flat pixels per block with
// two 64-bit divisions per pixel: 0.86 ms at 2 lines per GPU on 128 of the 148 SMs).
constexpr int kStemRun = 1024;
constexpr int kStemGroups = 64;
__host__ __device__ inline int stem_runs(int W) { return (W + kStemRun - 1) / kStemRun; }

__global__ void __launch_bounds__(256)
stem_wgrad_partial_kernel(const __nv_bfloat16* __restrict__ dz, const float* __restrict__ x, int B, int H, int W,
                          float* __restrict__ partial) {
    __shared__ float rows[3][kStemRun + 2];
    __shared__ float red[4][64][9];
    const int co = threadIdx.x & 63, g = threadIdx.x >> 6;
    const int runs = stem_runs(W);
    const int run = blockIdx.x % runs;
    const int bh = blockIdx.x / runs;                       // b * H + h
    const int h = bh % H;
    const int w0 = run * kStemRun;
    const int n = min(kStemRun, W - w0);
    const float* img = x + (size_t)(bh - h) * W;            // image of line b
    for (int i = threadIdx.x; i < 3 * (kStemRun + 2); i += blockDim.x) {
        const int r = i / (kStemRun + 2), c = i - r * (kStemRun + 2);
        const int hh = h + r - 1, ww = w0 + c - 1;
        rows[r][c] = (hh >= 0 && hh < H && ww >= 0 && ww < W && c < n + 2) ? __ldg(img + (size_t)hh * W + ww) : 0.f;
    }
    __syncthreads();
    float acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    const __nv_bfloat16* dzr = dz + ((size_t)bh * W + w0) * 64 + co;
#pragma unroll 4
    for (int w = g; w < n; w += 4) {
        const float d = __bfloat162float(dzr[(size_t)w * 64]);
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) acc[kh * 3 + kw] = fmaf(d, rows[kh][w + kw], acc[kh * 3 + kw]);
    }
#pragma unroll
    for (int t = 0; t < 9; ++t) red[g][co][t] = acc[t];
    __syncthreads();
    for (int i = threadIdx.x; i < 64 * 9; i += blockDim.x) {
        const int c = i / 9, t = i - c * 9;
        partial[(size_t)blockIdx.x * 576 + i] = red[0][c][t] + red[1][c][t] + red[2][c][t] + red[3][c][t];
    }
}

// ---------------------------------------------------------------- optimizer tail
// stage 1: per-block partial sums of g^2 over a flat fp32 gradient buffer
__global__ void __launch_bounds__(256)
sqnorm_partial_kernel(const float* __restrict__ g, long long n, float* __restrict__ partial) {
    __shared__ float red[8];
    float s = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float v = g[i];
        s = fmaf(v, v, s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < 8; ++i) t += red[i];
        partial[blockIdx.x] = t;
    }
}
// stage 2: total norm (fixed order) -> clip coefficient (torch.nn.utils.clip_grad_norm_: max_norm / (norm + 1e-6), <= 1)
// A non-finite norm (NaN/Inf gradient) publishes out[2] = 1 and coef = 0: the update is then skipped entirely, as the
// reference does twice over (main.py:413 skips a batch whose loss is not finite; GradScaler.step skips optimizer.step()
// when it found an Inf) - clipping alone would write NaN into every parameter (inf * 0).
__global__ void clip_coef_kernel(const float* __restrict__ partial, int n, float grad_scale, float max_norm,
                                 float* __restrict__ out /* [0]=total_norm, [1]=coef, [2]=1 if the step is skipped */) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < n; ++i) s += (double)partial[i];
        const float norm = sqrtf((float)s) * grad_scale;
        const bool finite = isfinite(norm);
        float coef = max_norm / (norm + 1e-6f);
        if (coef > 1.f) coef = 1.f;
        out[0] = norm;
        out[1] = !finite ? 0.f : (max_norm > 0.f ? coef : 1.f);
        out[2] = finite ? 0.f : 1.f;
    }
}
// stage 3: SGD(momentum, weight_decay) exactly as torch.optim.SGD (dampening 0, no nesterov):
//   g = grad*grad_scale*coef + wd*p ; buf = first ? g : momentum*buf + g ; p -= lr*buf
__global__ void __launch_bounds__(256)
sgd_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ buf, long long n,
                const float* __restrict__ coef, float grad_scale, float lr, float momentum, float wd, int first) {
    if (coef[2] != 0.f) {
        // skipped step: parameters and momentum stay as they are; a skipped FIRST step leaves a zero momentum buffer, so the
        // next step (momentum*0 + g) is the first step torch would have taken
        if (first)
            for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) buf[i] = 0.f;
        return;
    }
    const float c = coef[1] * grad_scale;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float w = p[i];
        const float gv = fmaf(wd, w, g[i] * c);
        const float b = first ? gv : fmaf(momentum, buf[i], gv);
        buf[i] = b;
        p[i] = w - lr * b;
    }
}


// ---------------------------------------------------------------- weight packing, all layers in one launch
// fp32 parameters in the reference layout (OIHW; nn.Linear [N][Cf*Hf]) -> the bf16 operand layouts of the implicit GEMMs:
// forward [Cout][tap][Cin] and data-gradient [Cin][tap][Cout] (classifier: [tap][Cin][pitch]). The optimizer rewrites every
// parameter each step, so this runs once per training step: one launch over a descriptor table instead of ~200 small torch
// permute / contiguous / cast kernels (0.76 ms of GPU time and most of the host time of a 2-line step). A tile is 32 output
// channels x 32 input channels x all taps, staged through shared memory so that reads and both writes are contiguous runs.
constexpr int kPackTile = 32;
constexpr int kPackMaxTaps = 9;

// TAPS is a compile-time constant (1, 4 or 9) so that the tile index arithmetic is multiplications and shifts: with a run-time
// divisor the three loops spent ~100 instructions per element on integer division (0.36 ms per step for 424 MB of traffic).
template <int TAPS>
__device__ __forceinline__ void pack_tile(const hctr_pack_desc& d, int lt, float (*tile)[kPackTile * kPackMaxTaps + 1]) {
    const int cin = d.cin, cout = d.cout;
    const int ctiles = (cin + kPackTile - 1) / kPackTile;
    const int o0 = (lt / ctiles) * kPackTile, c0 = (lt % ctiles) * kPackTile;
    constexpr int ncol = kPackTile * TAPS;                          // floats of one output channel inside the tile
    const int cw = min(kPackTile, cin - c0) * TAPS;                 // valid ones
    for (int i = threadIdx.x; i < kPackTile * ncol; i += blockDim.x) {
        const int row = i / ncol, col = i - row * ncol;
        float v = 0.f;
        if (o0 + row < cout && col < cw) v = __ldg(d.src + ((long long)(o0 + row) * cin + c0) * TAPS + col);
        tile[row][col] = v;
    }
    __syncthreads();
    __nv_bfloat16* fwd = static_cast<__nv_bfloat16*>(d.dst_fwd);
    for (int i = threadIdx.x; i < kPackTile * ncol; i += blockDim.x) {
        const int cl = i & (kPackTile - 1), rest = i >> 5;
        const int row = rest / TAPS, tap = rest - row * TAPS;
        if (o0 + row < cout && c0 + cl < cin)
            fwd[((long long)(o0 + row) * TAPS + tap) * cin + c0 + cl] = __float2bfloat16_rn(tile[row][cl * TAPS + tap]);
    }
    if (d.dst_bwd != nullptr) {
        __nv_bfloat16* bwd = static_cast<__nv_bfloat16*>(d.dst_bwd);
        for (int i = threadIdx.x; i < kPackTile * ncol; i += blockDim.x) {
            const int row = i & (kPackTile - 1), rest = i >> 5;
            const int cl = rest / TAPS, tap = rest - cl * TAPS;
            if (o0 + row < cout && c0 + cl < cin) {
                const long long o = d.bwd_mode == 0 ? ((long long)(c0 + cl) * TAPS + tap) * cout + o0 + row
                                                    : ((long long)tap * cin + c0 + cl) * d.bwd_pitch + o0 + row;
                bwd[o] = __float2bfloat16_rn(tile[row][cl * TAPS + tap]);
            }
        }
    }
}

__global__ void __launch_bounds__(256)
pack_weights_kernel(const hctr_pack_desc* __restrict__ descs, int ndesc) {
    __shared__ float tile[kPackTile][kPackTile * kPackMaxTaps + 1];
    const long long gt = blockIdx.x;
    int lo = 0, hi = ndesc - 1;                                     // last descriptor whose tile_start <= gt
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (descs[mid].tile_start <= gt) lo = mid; else hi = mid - 1;
    }
    const hctr_pack_desc d = descs[lo];
    const int lt = (int)(gt - d.tile_start);
    switch (d.taps) {                                               // block-uniform
        case 9: pack_tile<9>(d, lt, tile); break;
        case 4: pack_tile<4>(d, lt, tile); break;
        case 1: pack_tile<1>(d, lt, tile); break;
        default: break;                                             // rejected on the host
    }
}

}  // namespace hctr

using namespace hctr;

extern "C" {

int hctr_colsum_bf16(const void* x, long long rows, int cols, long long pitch, float* out, float* workspace,
                     long long workspace_bytes, void* stream) {
    HCTR_CHECK(x && out && workspace, HCTR_ERR_INVALID, "colsum: null pointer");
    const long long slices = (rows + kColRows - 1) / kColRows;
    HCTR_CHECK(workspace_bytes >= slices * cols * 4, HCTR_ERR_INVALID, "colsum: workspace too small");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    dim3 grid((cols + 255) / 256, (unsigned)slices);
    colsum_partial_kernel<<<grid, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(x), rows, cols, pitch, workspace);
    HCTR_CUDA(cudaGetLastError());
    colsum_final_kernel<<<(cols + 255) / 256, 256, 0, s>>>(workspace, (int)slices, cols, out);
    HCTR_CUDA(cudaGetLastError());
    return HCTR_OK;
}
long long hctr_colsum_workspace_bytes(long long rows, int cols) { return ((rows + kColRows - 1) / kColRows) * cols * 4; }

int hctr_stem_wgrad(const void* dz, const float* x, float* dw, int B, int H, int W, float* workspace,
                    long long workspace_bytes, void* stream) {
    HCTR_CHECK(dz && x && dw && workspace, HCTR_ERR_INVALID, "stem_wgrad: null pointer");
    const long long blocks = (long long)B * H * stem_runs(W);
    HCTR_CHECK(blocks < (1ll << 31), HCTR_ERR_INVALID, "stem_wgrad: too many rows");
    HCTR_CHECK(workspace_bytes >= (blocks + kStemGroups) * 576 * 4, HCTR_ERR_INVALID, "stem_wgrad: workspace too small");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    stem_wgrad_partial_kernel<<<(int)blocks, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(dz), x, B, H, W, workspace);
    HCTR_CUDA(cudaGetLastError());
    // two-stage fixed-order reduction of the per-block partials (up to B*H*2 of them)
    float* stage = workspace + blocks * 576;
    colsum_strided_kernel<<<dim3((576 + 255) / 256, kStemGroups), 256, 0, s>>>(workspace, (int)blocks, 576, kStemGroups, stage);
    HCTR_CUDA(cudaGetLastError());
    colsum_final_kernel<<<(576 + 255) / 256, 256, 0, s>>>(stage, kStemGroups, 576, dw);
    HCTR_CUDA(cudaGetLastError());
    return HCTR_OK;
}
long long hctr_stem_wgrad_workspace_bytes(int B, int H, int W) {
    return ((long long)B * H * stem_runs(W) + kStemGroups) * 576 * 4;
}

int hctr_sgd_clip_step(float* params, const float* grads, float* momentum_buf, long long n, float grad_scale,
                       float max_norm, float lr, float momentum, float weight_decay, int first_step, float* norm_out,
                       float* workspace, void* stream) {
    HCTR_CHECK(params && grads && momentum_buf && norm_out && workspace, HCTR_ERR_INVALID, "sgd: null pointer");
    HCTR_CHECK(n > 0, HCTR_ERR_INVALID, "sgd: empty parameter buffer");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int blocks = 148 * 8;
    sqnorm_partial_kernel<<<blocks, 256, 0, s>>>(grads, n, workspace);
    HCTR_CUDA(cudaGetLastError());
    clip_coef_kernel<<<1, 32, 0, s>>>(workspace, blocks, grad_scale, max_norm, norm_out);
    HCTR_CUDA(cudaGetLastError());
    sgd_step_kernel<<<blocks, 256, 0, s>>>(params, grads, momentum_buf, n, norm_out, grad_scale, lr, momentum, weight_decay, first_step);
    HCTR_CUDA(cudaGetLastError());
    return HCTR_OK;
}
long long hctr_sgd_workspace_bytes(void) { return 148 * 8 * 4; }

int hctr_pack_weights(const hctr_pack_desc* descs_device, int ndesc, long long total_tiles, void* stream) {
    HCTR_CHECK(descs_device != nullptr && ndesc > 0, HCTR_ERR_INVALID, "pack_weights: empty descriptor table");
    HCTR_CHECK(total_tiles > 0 && total_tiles < (1ll << 31), HCTR_ERR_INVALID, "pack_weights: bad tile count %lld", total_tiles);
    // (taps must be 1, 4 or 9: 1x1 / 3x3 convolutions and the 4-row classifier; other descriptors are skipped by the kernel)
    pack_weights_kernel<<<(unsigned)total_tiles, 256, 0, static_cast<cudaStream_t>(stream)>>>(descs_device, ndesc);
    HCTR_CUDA(cudaGetLastError());
    return HCTR_OK;
}

}  // extern "C"
