// Implicit-GEMM convolution on a CTA PAIR: tcgen05.mma.cta_group::2, M = 256 pixels x N = 256 channels per instruction.
//
// Same contraction as igemm_tcgen05.cuh (reference: nn.Conv2d 3x3/1x1 + BN + ReLU, models/handwritten_ctr_model.py:37-45,73-92)
//   D[pixel, n] = sum_{tap, c} X[b, h + dh(tap), w + dw(tap), c] * Wt[n, tap, c]
// but one tile is owned by the two SMs of a cluster: CTA r holds the 128 pixels of image row 2*h_tile + r (its half of
// M) and 128 of the 256 weight rows (its half of N) in its own shared memory; the leader CTA issues every MMA and the
// tensor cores of both SMs read both halves. Per CTA and K block that is 16 KB of activations + 16 KB of weights for
// 128 x 256 x 64 MACs - the same L2->SMEM traffic per flop as the single-CTA kernel with two sub-tiles, half the
// shared-memory operand reads per flop, and a 128-lane x 256-column accumulator per CTA, so TWO accumulator stages fit
// TMEM and the epilogue of tile i overlaps the main loop of tile i+1 (the single-CTA N=256 kernel has one stage).
//
// Barriers (same offsets in both CTAs):
//   full[s]      leader only, 1 arrival: the leader's producer (arrive.expect_tx of BOTH CTAs' bytes); both CTAs' TMA loads
//                complete_tx on the leader's barrier (.cta_group::2). The peer needs no arrival of its own: it can only
//                load into a stage after the multicast commit that freed it, i.e. after the previous phase completed,
//                and a transaction count may go negative inside a phase.
//   empty[s]     per CTA, 1 arrival: tcgen05.commit multicast from the leader frees the stage in both CTAs
//   acc_full[a]  per CTA, 1 arrival: multicast commit after the last K block
//   acc_empty[a] leader only, 2 x kEpiWarps arrivals: the epilogue warps of both CTAs (the peer's arrive remotely)
#pragma once
#include "igemm_tcgen05.cuh"

namespace hctr {

constexpr int kPairKSub = 2;        // 64-element K blocks per pipeline stage (fewer barrier round trips per flop)
constexpr int kPairAcc = 2;

// KWF (3x3 only, as in igemm_tcgen05.cuh): a stage is one (kh, 64-channel chunk): a 136-pixel activation slab that
// serves the three kw taps by a whole-row shift of the descriptor start address, plus the three weight half-tiles.
template <int BLOCK_N, int STAGES, int KWF>
struct PairSmem {
    static constexpr int kABytes = (KWF ? kSlabPix : kTileM) * kBlockK * 2;   // 16 KB (17 KB slab)
    static constexpr int kBBytes = (BLOCK_N / 2) * kBlockK * 2;        // this CTA's half of the weight rows
    static constexpr int kStageBytes = KWF ? kABytes + 3 * kBBytes : kPairKSub * (kABytes + kBBytes);
    static constexpr int kBarBytes = 1024;
    // epilogue staging: per epilogue warp 32 pixel rows of one 32-channel chunk (64 B, pitch 80 B = conflict-free for
    // both the per-pixel and the transposed access), used to turn per-thread 16-byte global accesses at a 2*N-byte stride
    // (128 L2 requests per warp and chunk) into 64-byte runs (32 requests)
    static constexpr int kEpiPitch = 80;
    static constexpr int kEpiBytes = kEpiWarps * 32 * kEpiPitch;
    // per epilogue warp: scale[] then shift[] of its BLOCK_N/2 columns of the current tile (broadcast 16-byte reads)
    static constexpr int kSclBytes = kEpiWarps * (BLOCK_N / 2) * 2 * 4;
    static constexpr int kTotal = STAGES * kStageBytes + kBarBytes + kEpiBytes + kSclBytes + 1024 /*alignment slack*/;
    static_assert(kTotal <= 227 * 1024, "shared memory of one CTA");
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory location in CTA `rank` of this cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t smem_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    // default semantics (release at CTA scope), as CUTLASS does for the 2-SM accumulator release: an explicit
    // .release.cluster costs a MEMBAR.ALL.GPU per arrive (measured: it halved the kernel when issued per K block)
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" :: "r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* m, uint32_t leader_bar, int32_t c0, int32_t c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4}], [%2];"
        :: "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(leader_bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_4d_pair(void* smem_dst, const CUtensorMap* m, uint32_t leader_bar,
                                                 int32_t c0, int32_t c1, int32_t c2, int32_t c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6}], [%2];"
        :: "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(leader_bar),
           "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                 :: "r"(smem_u32(smem_slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" :: "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive (once all MMAs issued so far have completed) on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
    const uint16_t mask = 3;
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 :: "r"(smem_u32(bar)), "h"(mask) : "memory");
}

// The i-th tile of CTA pair `pair_id`. Two orders:
//   p.col_mode == 0: tiles dealt round-robin, channel tile fastest, then the 128-px span, the row pair, the line
//   p.col_mode == 1: whole (line, 128-px span) columns dealt round-robin; inside a column the row pairs top to bottom,
//                    channel tile fastest. A pair then streams every input row of its column exactly once (the halo rows
//                    of one row pair are the rows of the next) and no two pairs - in particular no two dies - read the same
//                    rows: measured DRAM reads of a 512->512 launch were 1.4x the input with round-robin tiles.
struct PairTile { int n_tile, w_tile, h_tile, b; };
__device__ __forceinline__ int pair_tile_count(const IgemmParams& p, int pair_id, int num_pairs) {
    if (!p.col_mode) return (p.total_tiles - pair_id + num_pairs - 1) / num_pairs;
    const int ncols = p.B * p.w_tiles;
    return ((ncols - pair_id + num_pairs - 1) / num_pairs) * (p.h_tiles * p.n_tiles);
}
__device__ __forceinline__ PairTile pair_tile(const IgemmParams& p, int pair_id, int num_pairs, int i) {
    PairTile t;
    if (!p.col_mode) {
        const int tile = pair_id + i * num_pairs;
        t.n_tile = tile % p.n_tiles;
        int m = tile / p.n_tiles;
        t.w_tile = m % p.w_tiles; m /= p.w_tiles;
        t.h_tile = m % p.h_tiles;
        t.b = m / p.h_tiles;
    } else {
        const int per_col = p.h_tiles * p.n_tiles;
        const int col = pair_id + (i / per_col) * num_pairs;
        const int r = i % per_col;
        t.n_tile = r % p.n_tiles;
        t.h_tile = r / p.n_tiles;
        t.w_tile = col % p.w_tiles;
        t.b = col / p.w_tiles;
    }
    return t;
}

// EPI_CONV only, no pooling (the rows of a pool pair live in different CTAs). p.h_tiles = H/2 pair rows.
// BLOCK_N = 256 (Cout % 256 == 0) or 128 (Cout == 128: the single-CTA N=128 tile reads 128 B/clk of operands from
// shared memory, the limit; a pair reads 96).
// ADD = 1: the epilogue adds a residual tensor (p.add; dgrad, and conv2 of a residual block with p.gate): the residual of
// the whole tile row is fetched into registers BEFORE the wait for the accumulator, so that its latency hides behind the
// main loop (fetched per 32-column chunk right before use it made the epilogue longer than the main loop of the
// K = 2304 layers: 1200 instead of 1590 TFLOP/s). That variant has no room for the channel-sum butterfly (se_partial).
template <int BLOCK_N, int STAGES, int KWF, int ADD>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kIgemmThreads, 1)
igemm_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const IgemmParams p) {
    using L = PairSmem<BLOCK_N, STAGES, KWF>;
    constexpr int kPairStages = STAGES;
    constexpr int kPairBlockN = BLOCK_N;
    constexpr int kAccCols = kPairBlockN;
    constexpr int kTmemCols = kPairAcc * kAccCols;     // 512 or 256

    extern __shared__ uint8_t smem_raw[];
    // 1 KB alignment by POINTER arithmetic on the __shared__ array: rounding the address as an integer and casting it back made
    // every later access through `smem` a generic LD/ST (the compiler no longer knew the address space) - 495 generic loads and
    // 380 generic stores in the conv kernels' epilogues, long-scoreboard stalls at the staged stores (ncu source view)
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* bar_base = smem + kPairStages * L::kStageBytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(bar_base);            // [kPairStages]   (leader's are used)
    uint64_t* empty_bar = full_bar + kPairStages;                          // [kPairStages]
    uint64_t* acc_full = empty_bar + kPairStages;                          // [kPairAcc]
    uint64_t* acc_empty = acc_full + kPairAcc;                             // [kPairAcc]      (leader's are used)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + kPairAcc);
    uint8_t* epi_base = bar_base + L::kBarBytes;                           // [kEpiWarps][32][kEpiPitch]
    float* scl_base = reinterpret_cast<float*>(epi_base + L::kEpiBytes);   // [kEpiWarps][2][kPairBlockN/2]

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int i = 0; i < kPairStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        for (int i = 0; i < kPairAcc; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 2 * kEpiWarps); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_pair(tmem_slot, kTmemCols);
    tc_fence_before();
    cluster_sync_all();                       // barriers initialised and TMEM allocated in both CTAs
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // pipeline steps: KWF one per (kh, chunk); otherwise pairs of 64-wide K blocks (host guarantees divisibility)
    const int kblocks = KWF ? 3 * p.cin_chunks : p.ntaps * p.cin_chunks / kPairKSub;
    const int pair_id = blockIdx.x >> 1;
    const int num_pairs = gridDim.x >> 1;
    const int my_tiles = pair_tile_count(p, pair_id, num_pairs);

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer (both CTAs)
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0;
            for (int it = 0; it < my_tiles; ++it) {
                const PairTile tl = pair_tile(p, pair_id, num_pairs, it);
                const int n_tile = tl.n_tile, b = tl.b;
                const int h = tl.h_tile * 2 + (int)rank;           // this CTA's image row
                const int w0 = tl.w_tile * kTileM;
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    const uint32_t lbar = mapa_u32(smem_u32(&full_bar[stage]), 0);
                    if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * L::kStageBytes);
                    uint8_t* st = smem + stage * L::kStageBytes;
                    if constexpr (KWF) {
                        const int kh = kb / p.cin_chunks;
                        const int ch = kb - kh * p.cin_chunks;
                        tma_load_4d_pair(st, &tmA, lbar, ch * kBlockK, w0 - 1, h + p.tap_dh[kh * 3], b);
#pragma unroll
                        for (int kw = 0; kw < 3; ++kw)
                            tma_load_2d_pair(st + L::kABytes + kw * L::kBBytes, &tmB, lbar, ((kh * 3 + kw) * p.cin_chunks + ch) * kBlockK,
                                             n_tile * kPairBlockN + (int)rank * (kPairBlockN / 2));
                    } else {
#pragma unroll
                        for (int q = 0; q < kPairKSub; ++q) {
                            const int k64 = kb * kPairKSub + q;
                            const int tap = k64 / p.cin_chunks;
                            const int ch = k64 - tap * p.cin_chunks;
                            tma_load_4d_pair(st + q * (L::kABytes + L::kBBytes), &tmA, lbar, ch * kBlockK, w0 + p.tap_dw[tap], h + p.tap_dh[tap], b);
                            tma_load_2d_pair(st + q * (L::kABytes + L::kBBytes) + L::kABytes, &tmB, lbar, k64 * kBlockK,
                                             n_tile * kPairBlockN + (int)rank * (kPairBlockN / 2));
                        }
                    }
                    if (++stage == kPairStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer (leader CTA only)
        if (leader) {
            constexpr uint32_t idesc = make_idesc_bf16(2 * kTileM, kPairBlockN);
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int it = 0; it < my_tiles; ++it) {
                mbar_wait(&acc_empty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_base = tmem_base + acc * kAccCols;
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    if (elect_one()) {
                        if constexpr (KWF) {
                            const int kh = kb / p.cin_chunks;
                            const uint32_t a_addr = smem_u32(smem + stage * L::kStageBytes);
                            const uint32_t b_addr = a_addr + L::kABytes;
#pragma unroll
                            for (int kw = 0; kw < 3; ++kw) {
                                const uint32_t shift = static_cast<uint32_t>(p.tap_dw[kh * 3 + kw] + 1);      // rows into the slab
#pragma unroll
                                for (int k = 0; k < kBlockK / kUmmaK; ++k) {
                                    const uint64_t da = make_sw128_kmajor_desc(a_addr + shift * 128 + k * kUmmaK * 2);
                                    const uint64_t db = make_sw128_kmajor_desc(b_addr + kw * L::kBBytes + k * kUmmaK * 2);
                                    umma_bf16_pair(d_base, da, db, idesc, (kb | kw | k) != 0 ? 1u : 0u);
                                }
                            }
                        } else {
#pragma unroll
                            for (int q = 0; q < kPairKSub; ++q) {
                                const uint32_t a_addr = smem_u32(smem + stage * L::kStageBytes + q * (L::kABytes + L::kBBytes));
                                const uint32_t b_addr = a_addr + L::kABytes;
#pragma unroll
                                for (int k = 0; k < kBlockK / kUmmaK; ++k) {
                                    const uint64_t da = make_sw128_kmajor_desc(a_addr + k * kUmmaK * 2);
                                    const uint64_t db = make_sw128_kmajor_desc(b_addr + k * kUmmaK * 2);
                                    umma_bf16_pair(d_base, da, db, idesc, (kb | q | k) != 0 ? 1u : 0u);
                                }
                            }
                        }
                        umma_commit_pair(&empty_bar[stage]);             // stage free in both CTAs once these MMAs retire
                        if (kb == kblocks - 1) umma_commit_pair(&acc_full[acc]);
                    }
                    __syncwarp();
                    if (++stage == kPairStages) { stage = 0; phase ^= 1; }
                }
                if (++acc == kPairAcc) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ------------------------------------------------------------ epilogue (warps 2..9, both CTAs)
        const int quad = warp & 3;                      // TMEM lane quarter this warp may read
        const int half = (warp - 2) >> 2;               // which half of the accumulator columns this warp drains
        const int pix = quad * 32 + lane;               // pixel within the 128-px span
        const uint32_t lacc_empty0 = mapa_u32(smem_u32(&acc_empty[0]), 0);
        int acc = 0; uint32_t acc_phase = 0;
        for (int it = 0; it < my_tiles; ++it) {
            const PairTile tl = pair_tile(p, pair_id, num_pairs, it);
            const int n_tile = tl.n_tile, b = tl.b, w_tile = tl.w_tile;
            const int h = tl.h_tile * 2 + (int)rank;
            const int w = w_tile * kTileM + pix;

            constexpr int kChunks = kPairBlockN / 2 / 32;
            const bool ok = (w < p.W) && (h < p.H);
            // element offset of (this warp's pixel 0, first column of this warp's half); pixel r of the warp is r*N further
            const int wq = w_tile * kTileM + quad * 32;
            // pooled layers (p.pool): both CTAs of the pair own one of the two rows of a (2,1) max-pool window and combine
            // them in the pooled row h_tile of the [B][H/2][W][N] output with a vector bf16 max reduction in L2 (below)
            const int ho = p.pool ? tl.h_tile : h;
            const size_t warp_off = (p.out_line_pitch
                ? static_cast<size_t>(b) * p.out_line_pitch + (static_cast<size_t>(ho) * p.W + wq) * p.N
                : ((static_cast<size_t>(b) * p.out_H + ho) * p.W + wq) * p.N) + n_tile * kPairBlockN + half * (kPairBlockN / 2);
            // transposed role of this lane: 16-byte piece tq of pixel rows tr, tr+8, tr+16, tr+24
            const int tr = lane >> 2, tq = lane & 3;
            uint8_t* ebuf = epi_base + (warp - 2) * (32 * L::kEpiPitch);
            uint4 rb[ADD ? kChunks * 4 : 1];
            if constexpr (ADD) {
                // the residual of the whole tile row, fetched as 64-byte runs before the accumulator is awaited
                if (h < p.H) {
#pragma unroll
                    for (int ck = 0; ck < kChunks; ++ck)
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int r = tr + 8 * i;
                            rb[ck * 4 + i] = make_uint4(0u, 0u, 0u, 0u);
                            if (wq + r < p.W)
                                rb[ck * 4 + i] = ld_nc_v4(static_cast<const __nv_bfloat16*>(p.add) + warp_off + static_cast<size_t>(r) * p.N + ck * 32 + tq * 8);
                        }
                }
            }
            // per-channel epilogue factors of this tile, one column per lane, fetched before the accumulator is awaited and
            // parked in the warp's shared-memory strip (read back as broadcast 16-byte loads in the chunk loop, see
            // igemm_tcgen05.cuh): y = acc*scale + shift (conv bias and eval-mode BN folded, fp32), times the SE gate when it is
            // folded into the producing conv (out = relu(bn(conv) * gate[b, c] + residual))
            float* wsc = scl_base + (warp - 2) * (kPairBlockN / 2) * 2;
#pragma unroll
            for (int ck = 0; ck < kChunks; ++ck) {
                const int n = n_tile * kPairBlockN + half * (kPairBlockN / 2) + ck * 32 + lane;
                float sc_v = __ldg(p.scale + n), sh_v = __ldg(p.shift + n);
                if (p.gate) {
                    const float gt = __ldg(p.gate + static_cast<size_t>(b) * p.N + n);
                    sc_v *= gt; sh_v *= gt;
                }
                wsc[ck * 32 + lane] = sc_v;
                wsc[kPairBlockN / 2 + ck * 32 + lane] = sh_v;
            }
            __syncwarp();
            mbar_wait(&acc_full[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_base = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * kAccCols;
#pragma unroll
            for (int ck = 0; ck < kChunks; ++ck) {
                const int c0 = half * (kPairBlockN / 2) + ck * 32;
                const int n0 = n_tile * kPairBlockN + c0;
                float v[32];
                tmem_ld_32x32(t_base + c0, v);
                {
                    const float4* sc4 = reinterpret_cast<const float4*>(wsc + ck * 32);
                    const float4* sh4 = reinterpret_cast<const float4*>(wsc + kPairBlockN / 2 + ck * 32);
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float4 a = sc4[q], c = sh4[q];
                        v[4 * q + 0] = fmaf(v[4 * q + 0], a.x, c.x);
                        v[4 * q + 1] = fmaf(v[4 * q + 1], a.y, c.y);
                        v[4 * q + 2] = fmaf(v[4 * q + 2], a.z, c.z);
                        v[4 * q + 3] = fmaf(v[4 * q + 3], a.w, c.w);
                    }
                }
                if constexpr (ADD) {
                    // residual: transposed registers -> staging rows -> this lane's own pixel
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        *reinterpret_cast<uint4*>(ebuf + (tr + 8 * i) * L::kEpiPitch + tq * 16) = rb[ck * 4 + i];
                    __syncwarp();
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const uint4 a = *reinterpret_cast<const uint4*>(ebuf + lane * L::kEpiPitch + q * 16);
                        v[8 * q + 0] += bf16_lo(a.x); v[8 * q + 1] += bf16_hi(a.x);
                        v[8 * q + 2] += bf16_lo(a.y); v[8 * q + 3] += bf16_hi(a.y);
                        v[8 * q + 4] += bf16_lo(a.z); v[8 * q + 5] += bf16_hi(a.z);
                        v[8 * q + 6] += bf16_lo(a.w); v[8 * q + 7] += bf16_hi(a.w);
                    }
                    __syncwarp();
                }
                if constexpr (!ADD) {
                    if (p.se_partial && !p.sum_stored) {
                        // per-channel sum of the fp32 BN output (SELayer squeeze of this conv, models/handwritten_ctr_model.py:
                        // 27-28) over this warp's 32 pixels by a transpose-reduce butterfly; afterwards lane L holds column n0+L.
                        // One slot per (line, row, 128-px span, warp): fixed-order final sum in the consumer. (Sums of the values
                        // AS STORED - p.sum_stored, the NEXT conv's squeeze / train-mode BN statistics - come from the staged bf16
                        // chunk in the store loop below.)
                        float tsum[32];
#pragma unroll
                        for (int j = 0; j < 32; ++j) tsum[j] = ok ? v[j] : 0.f;
#define HCTR_BFLY(O)                                                                          \
                        {                                                                     \
                            const bool upper = (lane & (O)) != 0;                             \
                            _Pragma("unroll") for (int i = 0; i < (O); ++i) {                 \
                                const float send = upper ? tsum[i] : tsum[i + (O)];           \
                                const float keep = upper ? tsum[i + (O)] : tsum[i];           \
                                tsum[i] = keep + __shfl_xor_sync(0xffffffffu, send, (O));     \
                            }                                                                 \
                        }
                        HCTR_BFLY(16) HCTR_BFLY(8) HCTR_BFLY(4) HCTR_BFLY(2) HCTR_BFLY(1)
#undef HCTR_BFLY
                        const size_t slot = ((static_cast<size_t>(b) * p.H + h) * p.w_tiles + w_tile) * 4 + quad;
                        if (h < p.H) p.se_partial[slot * p.N + n0 + lane] = tsum[0];
                    }
                }
                // pack to bf16, stage this lane's pixel row, write the chunk out as 64-byte runs
                if (p.relu) stage_chunk_row<true>(v, ebuf + lane * L::kEpiPitch);
                else        stage_chunk_row<false>(v, ebuf + lane * L::kEpiPitch);
                __syncwarp();
                // (ADD == 0, p.sum_stored) channel sums / sums of squares of the chunk as stored, from the pieces read back for the
                // stores; the plain loop is kept apart so that layers without sums pay nothing for it
                if (ADD || !(p.se_partial && p.sum_stored)) {
                    if (h < p.H) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int r = tr + 8 * i;
                            if (wq + r < p.W) {
                                __nv_bfloat16* dst = static_cast<__nv_bfloat16*>(p.out) + warp_off + static_cast<size_t>(r) * p.N + ck * 32 + tq * 8;
                                const uint4 val = *reinterpret_cast<const uint4*>(ebuf + r * L::kEpiPitch + tq * 16);
                                if (p.pool) {
                                    // max(relu(a), relu(b)) with the output zero-filled by the host: every operand is >= 0, so
                                    // the order of the two CTAs' reductions does not matter and bf16 max is exact
                                    asm volatile("red.relaxed.gpu.global.max.noftz.v4.bf16x2 [%0], {%1,%2,%3,%4};"
                                                 :: "l"(dst), "r"(val.x), "r"(val.y), "r"(val.z), "r"(val.w) : "memory");
                                } else {
                                    *reinterpret_cast<uint4*>(dst) = val;
                                }
                            }
                        }
                    }
                } else {
                    float cs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, cq[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                    if (h < p.H) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int r = tr + 8 * i;
                            if (wq + r < p.W) {
                                __nv_bfloat16* dst = static_cast<__nv_bfloat16*>(p.out) + warp_off + static_cast<size_t>(r) * p.N + ck * 32 + tq * 8;
                                const uint4 val = *reinterpret_cast<const uint4*>(ebuf + r * L::kEpiPitch + tq * 16);
                                *reinterpret_cast<uint4*>(dst) = val;                     // (never pooled: the sums are of the stored tensor)
                                piece_add(val, cs);
                                if (p.sq_partial) piece_add_sq(val, cq);
                            }
                        }
                    }
                    const size_t slot = ((static_cast<size_t>(b) * p.H + h) * p.w_tiles + w_tile) * 4 + quad;
                    piece_rows_reduce(cs);
                    if (lane < 4 && h < p.H) piece_store(p.se_partial + slot * p.N + n0 + lane * 8, cs);
                    if (p.sq_partial) {
                        piece_rows_reduce(cq);
                        if (lane < 4 && h < p.H) piece_store(p.sq_partial + slot * p.N + n0 + lane * 8, cq);
                    }
                }
                __syncwarp();
            }
            // release this accumulator stage back to the leader's MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(lacc_empty0 + acc * 8);
            if (++acc == kPairAcc) { acc = 0; acc_phase ^= 1; }
        }
    }

    tc_fence_before();
    cluster_sync_all();                       // nobody may still be reading the peer's shared memory / TMEM
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, kTmemCols);
    }
}

}  // namespace hctr
