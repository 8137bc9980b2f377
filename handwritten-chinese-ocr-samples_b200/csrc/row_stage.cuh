// Warp-per-row staging of logits rows in shared memory (used by the CTC-loss row pass and the log-softmax/top-k pass).
//
// A row of C logits (fp32: 29.5 KB at C = 7375, bf16: 14.75 KB) is pulled into a per-warp shared-memory buffer by the
// bulk-copy engine (cp.async.bulk global -> shared, completion on a per-warp mbarrier), then one warp makes all its
// passes over the shared copy: HBM is read exactly once per row, and a warp needs no block barrier at all - the per-row
// bookkeeping of a 256-thread CTA per row (four block barriers, shared-memory hand-overs, idle lanes in the ranking)
// was the larger half of the instructions of the first top-k kernel. Rows need not be 16-byte aligned (C = 7375 floats
// contiguous): the bulk copy moves the 16-byte-aligned interior [lo, hi) of the row, the < 16-byte head and tail go through
// registers of the first lanes.
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

// One warp stages one row into `stage` (16-byte aligned; needs 16 + round16(row_bytes) + 16 bytes): lane 0 arms the
// barrier and starts the bulk copy of the interior, lanes < 16 carry the head / tail elements (ES = element size).
// The caller guarantees that nobody still reads the buffer (single buffer per warp) and has issued
// fence.proxy.async + __syncwarp() after its last generic-proxy access to it.
template <int ES>
__device__ __forceinline__ void warp_stage_row(unsigned char* stage, const RowGeom& r, int row_bytes, uint64_t* bar, int lane) {
    if (lane == 0) {
        if (r.body_bytes > 0) {
            mbar_arrive_expect_tx(bar, (uint32_t)r.body_bytes);
            bulk_g2s(stage + 16, r.a0 + r.head_bytes, (uint32_t)r.body_bytes, bar);
        } else {
            mbar_arrive(bar);
        }
    }
    const int nhead = r.head_bytes / ES;
    const int off = lane < nhead ? lane * ES : r.head_bytes + r.body_bytes + (lane - nhead) * ES;
    if (lane < 16 && off < row_bytes) {
        if (ES == 4) *reinterpret_cast<uint32_t*>(stage + 16 - r.head_bytes + off) = __ldg(reinterpret_cast<const uint32_t*>(r.a0 + off));
        else         *reinterpret_cast<unsigned short*>(stage + 16 - r.head_bytes + off) = __ldg(reinterpret_cast<const unsigned short*>(r.a0 + off));
    }
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
