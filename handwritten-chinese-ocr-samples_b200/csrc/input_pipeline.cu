// Device-side input pipeline for the recognition path: ToTensor + Normalize + right border padding of
// a batch of ragged grayscale text lines, written straight into the [B,1,H,Wb] fp32 tensor the stem conv reads.
// Reference: utils/dataset.py:78-93 (NormalizePAD: img/255, sub 0.5, div 0.5, pad with the last column) as used by
// test.py:170-186 and AlignCollate (utils/dataset.py:96-132).
#include <cmath>

#include "common.cuh"
#include "../../include/hctr_b200.h"

namespace hctr {

__global__ void __launch_bounds__(256)
normalize_pad_kernel(const uint8_t* __restrict__ pixels, const long long* __restrict__ offsets,
                     const int32_t* __restrict__ widths, float* __restrict__ out, int B, int H, int Wb) {
    const long long total = (long long)B * H * Wb;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int w = (int)(i % Wb);
        const long long bh = i / Wb;
        const int h = (int)(bh % H);
        const int b = (int)(bh / H);
        const int wi = widths[b];
        const int ws = w < wi ? w : wi - 1;                      // border replication of the last real column
        const float v = (float)pixels[offsets[b] + (long long)h * wi + ws];
        // same IEEE operations as torchvision ToTensor + sub_(0.5).div_(0.5): bit-exact
        out[i] = __fdiv_rn(__fsub_rn(__fdiv_rn(v, 255.0f), 0.5f), 0.5f);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// cv2.resize(..., interpolation=cv2.INTER_AREA) for 8-bit single-channel lines: the resize the reference applies in front
// of NormalizePAD (utils/dataset.py:53-57, test.py:206-214: height -> 128, width -> int(width * 128 / height)).
// One thread per destination pixel regenerates OpenCV's tables for its row and column with the same IEEE operations
// (double for the table geometry, float32 for the weights and sums, no FMA contraction) and accumulates in table order,
// so the result is bit-identical to OpenCV's C++ paths (oracle/resize.py restates them, pinned by cv2's own outputs):
//   both scales >= 1:  resizeArea_ (area-weighted mean, float32), or resizeAreaFast_ for integer scales
//   otherwise:         the bilinear fixed-point path with INTER_AREA's coefficient rule.
struct AreaSpan {            // computeResizeAreaTab entries of one destination index, in table order:
    int s1, s2;              //   [s1-1 with a_first]  s1 .. s2-1 with a_mid  [s2 with a_last]
    float a_first, a_mid, a_last;
    bool has_first, has_last;
};
__device__ __forceinline__ AreaSpan area_span(int d, double scale, int ssize) {
    AreaSpan t;
    const double f1 = __dmul_rn((double)d, scale);
    const double f2 = __dadd_rn(f1, scale);
    const double cell = fmin(scale, __dsub_rn((double)ssize, f1));
    int s1 = (int)ceil(f1);
    int s2 = (int)floor(f2);
    s2 = min(s2, ssize - 1);
    s1 = min(s1, s2);
    t.s1 = s1; t.s2 = s2;
    const double left = __dsub_rn((double)s1, f1);
    t.has_first = left > 1e-3;
    t.a_first = __double2float_rn(__ddiv_rn(left, cell));
    t.a_mid = __double2float_rn(__ddiv_rn(1.0, cell));
    const double right = __dsub_rn(f2, (double)s2);
    t.has_last = right > 1e-3;
    t.a_last = __double2float_rn(__ddiv_rn(fmin(fmin(right, 1.0), cell), cell));
    return t;
}

__device__ __forceinline__ uint8_t sat_u8_rn(float v) {           // saturate_cast<uchar>(float): round half to even, clamp
    const int i = __float2int_rn(v);
    return (uint8_t)min(max(i, 0), 255);
}

__device__ __forceinline__ float area_row(const uint8_t* __restrict__ row, const AreaSpan& x) {
    float buf = 0.f;
    if (x.has_first) buf = __fadd_rn(buf, __fmul_rn((float)row[x.s1 - 1], x.a_first));
    for (int sx = x.s1; sx < x.s2; ++sx) buf = __fadd_rn(buf, __fmul_rn((float)row[sx], x.a_mid));
    if (x.has_last) buf = __fadd_rn(buf, __fmul_rn((float)row[x.s2], x.a_last));
    return buf;
}

// coefficients of the bilinear path with INTER_AREA's rule for destination index d
struct LinTap { int s; int c0, c1; bool two; };
__device__ __forceinline__ LinTap lin_tap(int d, double scale, double inv, int ssize) {
    LinTap t;
    int s = (int)floor(__dmul_rn((double)d, scale));
    float f = __double2float_rn(__dsub_rn((double)(d + 1), __dmul_rn((double)(s + 1), inv)));
    f = f <= 0.f ? 0.f : __fsub_rn(f, floorf(f));
    if (s < 0) { f = 0.f; s = 0; }
    t.two = s + 1 < ssize;
    if (s >= ssize - 1) { f = 0.f; s = ssize - 1; }
    t.s = s;
    t.c0 = min(max(__float2int_rn(__fmul_rn(__fsub_rn(1.f, f), 2048.f)), -32768), 32767);
    t.c1 = min(max(__float2int_rn(__fmul_rn(f, 2048.f)), -32768), 32767);
    return t;
}

__global__ void __launch_bounds__(256)
resize_area_kernel(const uint8_t* __restrict__ src, int sh, int sw, long long spitch, uint8_t* __restrict__ dst, int dh, int dw,
                   long long dpitch, double scale_x, double scale_y, double inv_x, double inv_y, int mode, int ix, int iy) {
    const int dx = blockIdx.x * blockDim.x + threadIdx.x;
    const int dy = blockIdx.y;
    if (dx >= dw) return;
    uint8_t out;
    if (mode == 0) {                                       // resizeArea_
        const AreaSpan xs = area_span(dx, scale_x, sw);
        const AreaSpan ys = area_span(dy, scale_y, sh);
        float acc = 0.f;
        if (ys.has_first) acc = __fadd_rn(acc, __fmul_rn(ys.a_first, area_row(src + (long long)(ys.s1 - 1) * spitch, xs)));
        for (int sy = ys.s1; sy < ys.s2; ++sy) acc = __fadd_rn(acc, __fmul_rn(ys.a_mid, area_row(src + (long long)sy * spitch, xs)));
        if (ys.has_last) acc = __fadd_rn(acc, __fmul_rn(ys.a_last, area_row(src + (long long)ys.s2 * spitch, xs)));
        out = sat_u8_rn(acc);
    } else if (mode == 1) {                                // resizeAreaFast_: integer scale factors
        int sum = 0;
        for (int y = 0; y < iy; ++y) {
            const uint8_t* row = src + (long long)(dy * iy + y) * spitch + dx * ix;
            for (int x = 0; x < ix; ++x) sum += row[x];
        }
        out = (ix == 2 && iy == 2) ? (uint8_t)((sum + 2) >> 2) : sat_u8_rn(__fmul_rn((float)sum, 1.f / (float)(ix * iy)));
    } else {                                               // bilinear fixed point, INTER_AREA coefficients
        const LinTap tx = lin_tap(dx, scale_x, inv_x, sw);
        const LinTap ty = lin_tap(dy, scale_y, inv_y, sh);
        const int r0 = min(max(ty.s, 0), sh - 1), r1 = min(max(ty.s + 1, 0), sh - 1);
        const uint8_t* p0 = src + (long long)r0 * spitch;
        const uint8_t* p1 = src + (long long)r1 * spitch;
        const int h0 = tx.two ? p0[tx.s] * tx.c0 + p0[tx.s + 1] * tx.c1 : p0[tx.s] * 2048;
        const int h1 = tx.two ? p1[tx.s] * tx.c0 + p1[tx.s + 1] * tx.c1 : p1[tx.s] * 2048;
        const int v = (((ty.c0 * (h0 >> 4)) >> 16) + ((ty.c1 * (h1 >> 4)) >> 16) + 2) >> 2;
        out = (uint8_t)min(max(v, 0), 255);
    }
    dst[(long long)dy * dpitch + dx] = out;
}

}  // namespace hctr

using namespace hctr;

extern "C" int hctr_resize_area_u8(const void* src, int src_h, int src_w, long long src_pitch, void* dst, int dst_h, int dst_w,
                                   long long dst_pitch, void* stream) {
    HCTR_CHECK(src && dst, HCTR_ERR_INVALID, "resize_area: null pointer");
    HCTR_CHECK(src_h > 0 && src_w > 0 && dst_h > 0 && dst_w > 0, HCTR_ERR_INVALID, "resize_area: bad shape %dx%d -> %dx%d",
               src_h, src_w, dst_h, dst_w);
    HCTR_CHECK(src_pitch >= src_w && dst_pitch >= dst_w, HCTR_ERR_INVALID, "resize_area: pitch smaller than the row");
    HCTR_CHECK(dst_h <= 65535, HCTR_ERR_INVALID, "resize_area: destination too high (%d)", dst_h);
    // cv::resize: inv_scale = (double)dsize/ssize, scale = 1./inv_scale
    const double inv_x = (double)dst_w / src_w, inv_y = (double)dst_h / src_h;
    const double scale_x = 1.0 / inv_x, scale_y = 1.0 / inv_y;
    int mode = 2, ix = 0, iy = 0;
    if (scale_x >= 1 && scale_y >= 1) {
        ix = (int)lrint(scale_x); iy = (int)lrint(scale_y);
        const bool fast = fabs(scale_x - ix) < 2.220446049250313e-16 && fabs(scale_y - iy) < 2.220446049250313e-16;
        mode = fast ? 1 : 0;
    }
    dim3 grid((dst_w + 255) / 256, dst_h);
    resize_area_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const uint8_t*>(src), src_h, src_w, src_pitch, static_cast<uint8_t*>(dst), dst_h, dst_w, dst_pitch, scale_x,
        scale_y, inv_x, inv_y, mode, ix, iy);
    HCTR_CUDA(cudaGetLastError());
    return HCTR_OK;
}



extern "C" int hctr_normalize_pad(const void* pixels, const long long* offsets, const int32_t* widths, float* out, int B,
                                  int H, int Wb, void* stream) {
    HCTR_CHECK(pixels && offsets && widths && out, HCTR_ERR_INVALID, "normalize_pad: null pointer");
    HCTR_CHECK(B > 0 && H > 0 && Wb > 0, HCTR_ERR_INVALID, "normalize_pad: bad shape");
    const long long total = (long long)B * H * Wb;
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    normalize_pad_kernel<<<(int)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const uint8_t*>(pixels), offsets, widths, out, B, H, Wb);
    HCTR_CUDA(cudaGetLastError());
    return HCTR_OK;
}
