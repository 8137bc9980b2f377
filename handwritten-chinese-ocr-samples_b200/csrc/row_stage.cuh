// Helpers shared by the row-wise codec kernels (log-softmax / top-k, CTC loss): bulk copy of a row's aligned interior,
// row geometry for rows that are not 16-byte aligned (C = 7375 floats contiguous), warp reductions, exp2 on the MUFU pipe.
#pragma once
#include "common.cuh"

namespace hctr {

__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// geometry of one row in global memory; in its staging area element 0 lives at byte 16 - head_bytes
struct RowGeom {
    const unsigned char* a0;           // first byte of the row
    int head_bytes;                    // [a0, lo): bytes in front of the 16-byte-aligned interior (< 16)
    int body_bytes;                    // [lo, hi): multiple of 16 (0 for tiny rows: everything travels through registers)
};

__device__ __forceinline__ RowGeom row_geom(const void* row_ptr, int row_bytes) {
    RowGeom r;
    r.a0 = static_cast<const unsigned char*>(row_ptr);
    const uintptr_t a = reinterpret_cast<uintptr_t>(r.a0);
    const uintptr_t lo = (a + 15) & ~uintptr_t(15), hi = (a + row_bytes) & ~uintptr_t(15);
    if (hi > lo) { r.head_bytes = (int)(lo - a); r.body_bytes = (int)(hi - lo); }
    else { r.head_bytes = 0; r.body_bytes = 0; }
    return r;
}

// exp2 on the MUFU pipe; ex2.approx.ftz flushes results below 2^-126 to zero, which is what a sum of exponentials wants
__device__ __forceinline__ float ex2_fast(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

}  // namespace hctr
