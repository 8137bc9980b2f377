// Device-side input pipeline for the recognition path: ToTensor + Normalize + right border padding of
// a batch of ragged grayscale text lines, written straight into the [B,1,H,Wb] fp32 tensor the stem conv reads.
// Reference: utils/dataset.py:78-93 (NormalizePAD: img/255, sub 0.5, div 0.5, pad with the last column) as used by
// test.py:170-186 and AlignCollate (utils/dataset.py:96-132).
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

}  // namespace hctr

using namespace hctr;

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
