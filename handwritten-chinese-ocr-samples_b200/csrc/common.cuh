// Shared device helpers for the HCTR B200 kernels: error plumbing, bf16 packing,
// and thin inline-PTX wrappers for mbarrier / TMA / tcgen05 (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <atomic>
#include <cstdint>
#include <cstdio>

#if defined(__CUDA_ARCH__) && !defined(__CUDA_ARCH_FEAT_SM100_ALL)
#error "hctr_b200 kernels must be compiled for sm_100a (-gencode arch=compute_100a,code=sm_100a)"
#endif

namespace hctr {

// ---------------------------------------------------------------- host error plumbing
void set_error(const char* fmt, ...);
int  check_cuda(cudaError_t e, const char* what);

#define HCTR_CHECK(cond, code, ...)                 \
    do {                                            \
        if (!(cond)) {                              \
            ::hctr::set_error(__VA_ARGS__);         \
            return (code);                          \
        }                                           \
    } while (0)

#define HCTR_CUDA(expr)                                               \
    do {                                                              \
        int _rc = ::hctr::check_cuda((expr), #expr);                  \
        if (_rc != 0) return _rc;                                     \
    } while (0)

enum : int {
    HCTR_OK = 0,
    HCTR_ERR_INVALID = -1,   // bad argument (shape / alignment / null)
    HCTR_ERR_CUDA = -2,      // CUDA runtime / driver error
    HCTR_ERR_UNSUPPORTED = -3,
    HCTR_ERR_INDEX = -4,     // mirrors the reference's IndexError (empty greedy path in beam mode)
};

enum : int { HCTR_F32 = 0, HCTR_BF16 = 1 };

// ---------------------------------------------------------------- per-device one-time setup (host)
// cudaFuncSetAttribute and occupancy numbers belong to a device; a process may drive several (the ABI promises
// re-entrancy across devices). Each launcher keeps one of these per kernel: `int dev; if (once.need(dev)) {...; once.mark(dev);}`.
// Racing threads at worst configure twice (idempotent).
struct PerDeviceOnce {
    static constexpr int kMaxDev = 64;
    std::atomic<unsigned long long> done{0};
    int value[kMaxDev] = {};                     // optional per-device result (occupancy, SM count)
    bool need(int& dev) {
        dev = 0;
        cudaGetDevice(&dev);
        return dev < 0 || dev >= kMaxDev || !((done.load(std::memory_order_acquire) >> dev) & 1ull);
    }
    void mark(int dev, int v = 0) {
        if (dev < 0 || dev >= kMaxDev) return;
        value[dev] = v;
        done.fetch_or(1ull << dev, std::memory_order_release);
    }
    int get(int dev) const { return (dev >= 0 && dev < kMaxDev) ? value[dev] : 0; }
};

// ---------------------------------------------------------------- small device utils
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}

// the same with ReLU inside the conversion (cvt.rn.relu: negative -> +0, NaN stays NaN as in torch.relu): the epilogues'
// separate FMNMX per element were 32 of a chunk's ~250 instructions
__device__ __forceinline__ uint32_t pack_bf16x2_relu(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
// one pixel's 32 channels of a chunk -> 64 bytes of bf16 at `dst` (shared memory), ReLU folded into the conversion
template <bool RELU>
__device__ __forceinline__ void stage_chunk_row(const float (&v)[32], uint8_t* dst) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        uint32_t pk[4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
            pk[j] = RELU ? pack_bf16x2_relu(v[8 * q + 2 * j], v[8 * q + 2 * j + 1]) : pack_bf16x2(v[8 * q + 2 * j], v[8 * q + 2 * j + 1]);
        *reinterpret_cast<uint4*>(dst + q * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
}

__device__ __forceinline__ float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }

// Packed fp32 pairs (Blackwell FFMA2: two IEEE fp32 fused multiply-adds per instruction, each lane rounded exactly like fmaf;
// an operand built from the same scalar twice becomes the instruction's broadcast form, no move)
__device__ __forceinline__ unsigned long long f32x2_pack(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void f32x2_unpack(unsigned long long v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long f32x2_fma(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

// Channel sums of a staged chunk "as stored": the store loop of the conv epilogues reads the staged 32-pixel x 32-channel bf16
// chunk back as 16-byte pieces (8 channels of one pixel row; lane = (row & 7) * 4 + piece, rows r, r+8, r+16, r+24). Adding the
// pieces a lane holds and then the eight lanes of a piece (xor 4, 8, 16) gives the per-channel sums over the chunk's pixels in
// 24 shuffles - against 31 shuffles + 62 selects + the separate ReLU / bf16 rounding of 32 values per lane for the
// transpose-reduce butterfly on the fp32 registers. Lanes 0..3 end up holding channels piece*8 .. piece*8+7.
__device__ __forceinline__ void piece_add(const uint4& q, float (&s)[8]) {
    s[0] += bf16_lo(q.x); s[1] += bf16_hi(q.x); s[2] += bf16_lo(q.y); s[3] += bf16_hi(q.y);
    s[4] += bf16_lo(q.z); s[5] += bf16_hi(q.z); s[6] += bf16_lo(q.w); s[7] += bf16_hi(q.w);
}
__device__ __forceinline__ void piece_add_sq(const uint4& q, float (&t)[8]) {
    float x;
    x = bf16_lo(q.x); t[0] = fmaf(x, x, t[0]); x = bf16_hi(q.x); t[1] = fmaf(x, x, t[1]);
    x = bf16_lo(q.y); t[2] = fmaf(x, x, t[2]); x = bf16_hi(q.y); t[3] = fmaf(x, x, t[3]);
    x = bf16_lo(q.z); t[4] = fmaf(x, x, t[4]); x = bf16_hi(q.z); t[5] = fmaf(x, x, t[5]);
    x = bf16_lo(q.w); t[6] = fmaf(x, x, t[6]); x = bf16_hi(q.w); t[7] = fmaf(x, x, t[7]);
}
__device__ __forceinline__ void piece_rows_reduce(float (&s)[8]) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        s[k] += __shfl_xor_sync(0xffffffffu, s[k], 4);
        s[k] += __shfl_xor_sync(0xffffffffu, s[k], 8);
        s[k] += __shfl_xor_sync(0xffffffffu, s[k], 16);
    }
}
__device__ __forceinline__ void piece_store(float* dst, const float (&s)[8]) {     // dst: 32-byte aligned run of 8 floats
    *reinterpret_cast<float4*>(dst) = make_float4(s[0], s[1], s[2], s[3]);
    *reinterpret_cast<float4*>(dst + 4) = make_float4(s[4], s[5], s[6], s[7]);
}

__device__ __forceinline__ uint4 ld_nc_v4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "elect.sync _|p, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(pred));
    return pred != 0;
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                 :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must surface as a trapped kernel, never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 26)) {
            printf("hctr_b200: mbarrier wait timed out (block %d thread %d)\n", (int)blockIdx.x, (int)threadIdx.x);
            __trap();
        }
    }
}

// ---------------------------------------------------------------- TMA (cp.async.bulk.tensor)
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" :: "l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar,
                                            int32_t c0, int32_t c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4}], [%2];"
        :: "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)),
           "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar,
                                            int32_t c0, int32_t c1, int32_t c2, int32_t c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6}], [%2];"
        :: "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)),
           "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

// ---------------------------------------------------------------- tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 :: "r"(smem_u32(smem_slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after()  { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T, bf16 inputs, fp32 accumulate, single-CTA.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                          uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" :: "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// Arrive on an mbarrier once all previously issued MMAs have completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                 :: "r"(smem_u32(bar)) : "memory");
}

// K-major, 128-byte-swizzled operand tile: rows of 64 bf16 (128 B), 8-row swizzle atoms of
// 1024 B stacked along M/N (SBO = 1024 B). Bit layout follows the sm_100 shared-memory
// matrix descriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48),
// layout type [61,64) with SWIZZLE_128B = 2.
__device__ __forceinline__ uint64_t make_sw128_kmajor_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3ffffu) >> 4);
    d |= static_cast<uint64_t>(1) << 16;             // LBO (unused for swizzled K-major)
    d |= static_cast<uint64_t>(1024 >> 4) << 32;     // SBO
    d |= static_cast<uint64_t>(1) << 46;             // descriptor version (Blackwell)
    d |= static_cast<uint64_t>(2) << 61;             // SWIZZLE_128B
    return d;
}
// kind::f16 instruction descriptor: fp32 accumulate, bf16 A and B, both K-major.
__host__ __device__ constexpr uint32_t make_idesc_bf16(int m, int n) {
    return (1u << 4)                                  // C format = F32
         | (1u << 7)                                  // A format = BF16
         | (1u << 10)                                 // B format = BF16
         | (static_cast<uint32_t>(n >> 3) << 17)      // N / 8
         | (static_cast<uint32_t>(m >> 4) << 24);     // M / 16
}

// 32 lanes x 32 consecutive fp32 columns: thread i of the warp reads TMEM lane (base_lane + i).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

}  // namespace hctr
