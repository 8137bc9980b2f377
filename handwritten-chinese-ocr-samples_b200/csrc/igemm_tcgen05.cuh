// Implicit-GEMM on tcgen05 / TMEM, fed by TMA, for the HCTR backbone and classifier.
//
//   D[pixel, n] = sum_{tap, c} X[b, h + dh(tap), w + dw(tap), c] * Wt[n, tap, c]
//
// X is an NHWC bf16 activation tensor seen through a 4-D TMA tensor map (C, W, H, B); a box of
// (64 ch, 128 px, 1 row, 1 line) lands in shared memory as a K-major SWIZZLE_128B operand tile
// (128 rows x 128 B), and out-of-bounds coordinates are zero-filled by the TMA unit, which is
// exactly the conv's zero padding (reference: nn.Conv2d(.., 3, 1, 1),
// models/handwritten_ctr_model.py:37-41,73-92). Wt is the packed weight matrix [N][taps*Cin]
// (K-major) seen through a 2-D map. One CTA owns NUM_SUB accumulators of 128 pixels x BLOCK_N
// channels in TMEM; the two sub-tiles are either the two rows of a (2,1) max-pool pair
// (:123,129,136,143,150) or two adjacent 128-pixel spans of a line (classifier, :172-176).
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer,
// warps 2..9 = epilogue (TMEM -> registers -> BN/ReLU/pool or bias -> global); a warp may only read the TMEM lane
// quarter (warp % 4), so the eight epilogue warps form two groups that split the accumulator columns.
#pragma once
#include "common.cuh"

namespace hctr {

constexpr int kTileM = 128;          // pixels per accumulator (UMMA M)
constexpr int kBlockK = 64;          // bf16 elements per K block = one 128-byte swizzle row
constexpr int kUmmaK = 16;
constexpr int kEpiWarps = 8;           // two warps per TMEM lane quarter, each draining half of the columns
constexpr int kIgemmThreads = 64 + kEpiWarps * 32;
constexpr int kMaxTaps = 9;

enum : int { EPI_CONV = 0, EPI_LINEAR = 1 };

struct IgemmParams {
    // problem
    int B, H, W;              // input activation dims (NHWC)
    int cin_chunks;           // Cin / 64
    int ntaps;
    int8_t tap_dh[kMaxTaps];
    int8_t tap_dw[kMaxTaps];
    int sub_dh, sub_dw;       // offset of sub-tile s relative to sub-tile 0: (s*sub_dh rows, s*sub_dw*128 px)
    int N;                    // output channels / classes
    // tiling (derived on host)
    int w_tiles, h_tiles, n_tiles, total_tiles;
    int col_mode;             // CTA-pair kernel: 1 = deal whole (line, span) columns to the pairs (see igemm2_tcgen05.cuh)
    // epilogue
    const float* scale;       // [N] (EPI_CONV) or nullptr
    const float* shift;       // [N] BN shift (EPI_CONV) / bias (EPI_LINEAR)
    void* out;
    const void* add;          // EPI_CONV: optional bf16 tensor (same layout as out) added before the store (dgrad + residual)
    int relu, pool;
    int out_H;                // output rows (H or H/2 when pooled)
    long long out_line_pitch; // EPI_CONV: elements between consecutive lines b of the output (0 = out_H*W*N)
    float* se_partial;        // EPI_CONV: optional [B][h_tiles*w_tiles][4][N] per-(tile, warp) channel sums of the fp32
                              // epilogue output (the SE squeeze folded into the producing conv; deterministic)
    int sum_stored;           // 1: the sums are taken over the values as STORED (after gate/add/ReLU, rounded to bf16) -
                              // what the next convolution will read - instead of the fp32 BN output; slots are then per
                              // (line, row, span, warp): [B][2*h_tiles*w_tiles][4][N] (hctr_conv_sum_slices)
    float* sq_partial;        // EPI_CONV: optional, same slots as se_partial: sums of the SQUARES of the same values (train-mode
                              // BatchNorm statistics of z = conv(x)+bias taken in the epilogue: no separate pass over z)
    const float* gate;        // EPI_CONV: optional [B][N] per-(line, channel) factor applied after BN, before `add`
                              // (the SE gate folded into the producing conv: out = relu(bn(conv)*gate + add))
    int out_dtype;            // EPI_LINEAR: HCTR_F32 | HCTR_BF16
    long long out_pitch;      // EPI_LINEAR: elements between consecutive (b,w) rows
    float2* lse_partial;      // EPI_LINEAR: optional [B*W][n_tiles*2] (max, sum exp(x-max)) per (row, column half-tile) of the
                              // logits as stored (log_softmax fused into the classifier epilogue; combined by lse_combine)
    int2* argmax_partial;     // EPI_LINEAR: optional [B*W][n_tiles*2] (bits of the largest stored value, its class) per (row,
                              //   column half-tile): greedy decoding without ever reading - or, with out == nullptr, writing -
                              //   the logits (numpy argmax semantics: first maximum, the first NaN beats everything)
};

// KWF ("kw-fused", 3x3 convs only): one A box of 136 pixels (w0-1 .. w0+134) per (kh, 64-channel chunk) serves the three
// kw taps - the MMA for tap kw starts (dw+1) rows = (dw+1)*128 B further into the box. The 128-byte swizzle XOR is a
// function of the absolute shared-memory address (measured: results are exact with base_offset = 0 and wrong with the
// row phase in base_offset), so a whole-row shift of the start address needs nothing else. The activation tile then
// crosses L2->SMEM 3 instead of 9 times; a stage holds 3 weight tiles.
constexpr int kSlabPix = 136;

template <int BLOCK_N, int NUM_SUB, int STAGES, int KWF = 0>
struct IgemmSmem {
    static constexpr int kABytes = (KWF ? kSlabPix : kTileM) * kBlockK * 2;      // 16 KB (17 KB slab) per sub-tile
    static constexpr int kBBytes = BLOCK_N * kBlockK * 2;
    static constexpr int kStageBytes = NUM_SUB * kABytes + (KWF ? 3 : 1) * kBBytes;
    static constexpr int kBarBytes = 1024;
    // epilogue staging: per epilogue warp 32 pixel rows of one 32-channel chunk (64 B, pitch 80 B = conflict-free for the
    // per-pixel and the transposed access): per-thread 16-byte global stores at a 2*N-byte stride become 64-byte runs
    static constexpr int kEpiPitch = 80;
    static constexpr int kEpiBytes = kEpiWarps * 32 * kEpiPitch;
    // per epilogue warp: scale[] then shift[] of the warp's BLOCK_N/2 columns of the current tile (read back as broadcast
    // 16-byte loads in the chunk loop)
    static constexpr int kSclBytes = kEpiWarps * (BLOCK_N / 2) * 2 * 4;
    static constexpr int kTotal = STAGES * kStageBytes + kBarBytes + kEpiBytes + kSclBytes + 1024 /*alignment slack*/;
    static_assert(kTotal <= 227 * 1024, "shared memory of one CTA");
};

// The i-th tile of persistent CTA `cta`. Two orders (as in the CTA-pair kernel, igemm2_tcgen05.cuh):
//   p.col_mode == 0: tiles dealt round-robin, channel tile fastest, then the 128-px span, the row pair, the line
//   p.col_mode == 1: whole (line, span) columns dealt round-robin; inside a column the row pairs top to bottom, channel tile
//                    fastest: a CTA streams every input row of its column once (the halo rows of one tile are the rows of the
//                    next) and no two CTAs - in particular no two dies - read the same rows. Round-robin tiles read the
//                    input of a 128->128 launch 1.37x from DRAM (profiles/r1_ncu_thin128.txt).
struct CtaTile { int n_tile, w_tile, h_tile, b; };
__device__ __forceinline__ int cta_tile_count(const IgemmParams& p, int cta, int nctas) {
    if (!p.col_mode) return (p.total_tiles - cta + nctas - 1) / nctas;
    const int ncols = p.B * p.w_tiles;
    return ((ncols - cta + nctas - 1) / nctas) * (p.h_tiles * p.n_tiles);
}
__device__ __forceinline__ CtaTile cta_tile(const IgemmParams& p, int cta, int nctas, int i) {
    CtaTile t;
    if (!p.col_mode) {
        const int tile = cta + i * nctas;
        t.n_tile = tile % p.n_tiles;
        int m = tile / p.n_tiles;
        t.w_tile = m % p.w_tiles; m /= p.w_tiles;
        t.h_tile = m % p.h_tiles;
        t.b = m / p.h_tiles;
    } else {
        const int per_col = p.h_tiles * p.n_tiles;
        const int col = cta + (i / per_col) * nctas;
        const int r = i % per_col;
        t.n_tile = r % p.n_tiles;
        t.h_tile = r / p.n_tiles;
        t.w_tile = col % p.w_tiles;
        t.b = col / p.w_tiles;
    }
    return t;
}

template <int BLOCK_N, int NUM_SUB, int STAGES, int ACC_STAGES, int EPI, int KWF = 0>
__global__ void __launch_bounds__(kIgemmThreads, 1)
igemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA,
                     const __grid_constant__ CUtensorMap tmB,
                     const IgemmParams p) {
    using L = IgemmSmem<BLOCK_N, NUM_SUB, STAGES, KWF>;
    constexpr int kAccCols = NUM_SUB * BLOCK_N;
    constexpr int kTmemCols = ACC_STAGES * kAccCols;
    static_assert(kTmemCols <= 512 && (kTmemCols & (kTmemCols - 1)) == 0 && kTmemCols >= 32,
                  "TMEM allocation must be a power of two in [32, 512]");
    static_assert(BLOCK_N % 32 == 0 && BLOCK_N <= 256, "BLOCK_N");

    extern __shared__ uint8_t smem_raw[];
    // 1 KB alignment by POINTER arithmetic on the __shared__ array: rounding the address as an integer and casting it back made
    // every later access through `smem` a generic LD/ST (the compiler no longer knew the address space) - 495 generic loads and
    // 380 generic stores in the conv kernels' epilogues, long-scoreboard stalls at the staged stores (ncu source view)
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* bar_base = smem + STAGES * L::kStageBytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(bar_base);            // [STAGES]
    uint64_t* empty_bar = full_bar + STAGES;                               // [STAGES]
    uint64_t* acc_full = empty_bar + STAGES;                               // [ACC_STAGES]
    uint64_t* acc_empty = acc_full + ACC_STAGES;                           // [ACC_STAGES]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + ACC_STAGES);
    uint8_t* epi_base = bar_base + L::kBarBytes;                           // [kEpiWarps][32][kEpiPitch]
    float* scl_base = reinterpret_cast<float*>(epi_base + L::kEpiBytes);   // [kEpiWarps][2][BLOCK_N/2]

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
        for (int i = 0; i < ACC_STAGES; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], kEpiWarps); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int kblocks = (KWF ? 3 : p.ntaps) * p.cin_chunks;      // KWF: one K block = (kh, chunk) = three taps
    const int sub_rows = p.sub_dh ? NUM_SUB : 1;     // input rows covered by one tile
    const int sub_cols = p.sub_dw ? NUM_SUB : 1;     // 128-px spans covered by one tile
    const int my_tiles = cta_tile_count(p, blockIdx.x, gridDim.x);

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0;
            for (int ti = 0; ti < my_tiles; ++ti) {
                const CtaTile tl = cta_tile(p, blockIdx.x, gridDim.x, ti);
                const int n_tile = tl.n_tile, w_tile = tl.w_tile, h_tile = tl.h_tile, b = tl.b;
                const int h0 = h_tile * sub_rows;
                const int w0 = w_tile * sub_cols * kTileM;
                for (int kb = 0; kb < kblocks; ++kb) {
                    const int tap = kb / p.cin_chunks;                 // KWF: this is kh
                    const int ch = kb - tap * p.cin_chunks;
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    mbar_arrive_expect_tx(&full_bar[stage], L::kStageBytes);
                    uint8_t* st = smem + stage * L::kStageBytes;
                    if constexpr (KWF) {
#pragma unroll
                        for (int s = 0; s < NUM_SUB; ++s)
                            tma_load_4d(st + s * L::kABytes, &tmA, &full_bar[stage], ch * kBlockK, w0 - 1,
                                        h0 + s * p.sub_dh + p.tap_dh[tap * 3], b);
#pragma unroll
                        for (int kw = 0; kw < 3; ++kw)
                            tma_load_2d(st + NUM_SUB * L::kABytes + kw * L::kBBytes, &tmB, &full_bar[stage],
                                        ((tap * 3 + kw) * p.cin_chunks + ch) * kBlockK, n_tile * BLOCK_N);
                    } else {
#pragma unroll
                        for (int s = 0; s < NUM_SUB; ++s) {
                            tma_load_4d(st + s * L::kABytes, &tmA, &full_bar[stage],
                                        ch * kBlockK,
                                        w0 + s * p.sub_dw * kTileM + p.tap_dw[tap],
                                        h0 + s * p.sub_dh + p.tap_dh[tap],
                                        b);
                        }
                        tma_load_2d(st + NUM_SUB * L::kABytes, &tmB, &full_bar[stage],
                                    kb * kBlockK, n_tile * BLOCK_N);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer
        constexpr uint32_t idesc = make_idesc_bf16(kTileM, BLOCK_N);
        int stage = 0; uint32_t phase = 0;
        int acc = 0; uint32_t acc_phase = 0;
        for (int ti = 0; ti < my_tiles; ++ti) {
            mbar_wait(&acc_empty[acc], acc_phase ^ 1);
            tc_fence_after();
            const uint32_t d_base = tmem_base + acc * kAccCols;
            for (int kb = 0; kb < kblocks; ++kb) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t a_addr = smem_u32(smem + stage * L::kStageBytes);
                    const uint32_t b_addr = a_addr + NUM_SUB * L::kABytes;
                    if constexpr (KWF) {
                        const int kh = kb / p.cin_chunks;
#pragma unroll
                        for (int s = 0; s < NUM_SUB; ++s) {
#pragma unroll
                            for (int kw = 0; kw < 3; ++kw) {
                                const uint32_t shift = static_cast<uint32_t>(p.tap_dw[kh * 3 + kw] + 1);   // rows into the slab
#pragma unroll
                                for (int k = 0; k < kBlockK / kUmmaK; ++k) {
                                    uint64_t da = make_sw128_kmajor_desc(a_addr + s * L::kABytes + shift * 128 + k * kUmmaK * 2);
                                    const uint64_t db = make_sw128_kmajor_desc(b_addr + kw * L::kBBytes + k * kUmmaK * 2);
                                    umma_bf16(d_base + s * BLOCK_N, da, db, idesc, (kb | kw | k) != 0 ? 1u : 0u);
                                }
                            }
                        }
                    } else {
#pragma unroll
                        for (int s = 0; s < NUM_SUB; ++s) {
#pragma unroll
                            for (int k = 0; k < kBlockK / kUmmaK; ++k) {
                                const uint64_t da = make_sw128_kmajor_desc(a_addr + s * L::kABytes + k * kUmmaK * 2);
                                const uint64_t db = make_sw128_kmajor_desc(b_addr + k * kUmmaK * 2);
                                umma_bf16(d_base + s * BLOCK_N, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
                            }
                        }
                    }
                    umma_commit(&empty_bar[stage]);                 // smem slot free once these MMAs retire
                    if (kb == kblocks - 1) umma_commit(&acc_full[acc]);
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
        }
    } else {
        // ------------------------------------------------------------ epilogue (warps 2..9)
        const int quad = warp & 3;                      // TMEM lane quarter this warp may read
        const int half = (warp - 2) >> 2;               // which half of the accumulator columns this warp drains
        const int pix = quad * 32 + lane;               // pixel within the 128-px sub-tile
        int acc = 0; uint32_t acc_phase = 0;
        for (int ti = 0; ti < my_tiles; ++ti) {
            const CtaTile tl = cta_tile(p, blockIdx.x, gridDim.x, ti);
            const int n_tile = tl.n_tile, w_tile = tl.w_tile, h_tile = tl.h_tile, b = tl.b;
            const int h0 = h_tile * sub_rows;
            const int w0 = w_tile * sub_cols * kTileM;

            if constexpr (EPI == EPI_CONV) {
                if (p.add && !p.pool) {
                    // pull this tile's residual rows into L2 while its main loop is still running
#pragma unroll
                    for (int s = 0; s < NUM_SUB; ++s) {
                        const int w = w0 + s * p.sub_dw * kTileM + pix;
                        const int h = h0 + s * p.sub_dh;
                        if (w < p.W && h < p.out_H) {
                            const size_t off = p.out_line_pitch
                                ? static_cast<size_t>(b) * p.out_line_pitch + (static_cast<size_t>(h) * p.W + w) * p.N
                                : ((static_cast<size_t>(b) * p.out_H + h) * p.W + w) * p.N;
                            const __nv_bfloat16* r = static_cast<const __nv_bfloat16*>(p.add) + off + n_tile * BLOCK_N + half * (BLOCK_N / 2);
#pragma unroll
                            for (int q = 0; q < BLOCK_N / 2; q += 64) asm volatile("prefetch.global.L2 [%0];" :: "l"(r + q));
                        }
                    }
                }
            }
            // per-channel epilogue factors of this warp's columns: y = acc*scale + shift, times the SE gate when it is folded in.
            // One column per lane and chunk, fetched before the wait on the accumulator and parked in the warp's shared-memory
            // strip; the chunk loop reads them back as broadcast 16-byte loads (16 LDS per 32 columns). Vector __ldg's per chunk
            // sat on the epilogue's critical path; a shuffle broadcast per column (round 1) cost 64 of the ~300 instructions of
            // a chunk in an epilogue that is latency-bound on its own instruction chain (profiles/r2_ncu_conv1x1_64_128.txt).
            constexpr int kWarpChunks = BLOCK_N / 2 / 32;
            float* wsc = scl_base + (warp - 2) * (BLOCK_N / 2) * 2;
            if constexpr (EPI == EPI_CONV) {
#pragma unroll
                for (int ck = 0; ck < kWarpChunks; ++ck) {
                    const int n = n_tile * BLOCK_N + half * (BLOCK_N / 2) + ck * 32 + lane;
                    float sc_v = __ldg(p.scale + n), sh_v = __ldg(p.shift + n);
                    if (p.gate) {
                        const float gt = __ldg(p.gate + static_cast<size_t>(b) * p.N + n);
                        sc_v *= gt; sh_v *= gt;
                    }
                    wsc[ck * 32 + lane] = sc_v;
                    wsc[BLOCK_N / 2 + ck * 32 + lane] = sh_v;
                }
                __syncwarp();
            }
            const uint32_t wsc_u32 = smem_u32(wsc);
            uint8_t* ebuf = epi_base + (warp - 2) * (32 * L::kEpiPitch);
            const int tr = lane >> 2, tq = lane & 3;     // transposed role: 16-byte piece tq of pixel rows tr, tr+8, tr+16, tr+24
            mbar_wait(&acc_full[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_base = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * kAccCols;

            float run_m[NUM_SUB], run_s[NUM_SUB];
            float best_v[NUM_SUB];
            int best_i[NUM_SUB];
#pragma unroll
            for (int s = 0; s < NUM_SUB; ++s) { run_m[s] = -3.0e38f; run_s[s] = 0.f; best_v[s] = -INFINITY; best_i[s] = -1; }
#pragma unroll 1
            for (int ck = 0; ck < kWarpChunks; ++ck) {
                const int c0 = half * (BLOCK_N / 2) + ck * 32;
                const int n0 = n_tile * BLOCK_N + c0;
                if (n0 >= p.N) break;                    // warp-uniform
                float v[NUM_SUB][32];
#pragma unroll
                for (int s = 0; s < NUM_SUB; ++s) tmem_ld_32x32(t_base + s * BLOCK_N + c0, v[s]);

                if constexpr (EPI == EPI_CONV) {
                    // y = acc*scale + shift  (conv bias and eval-mode BN folded, fp32; SE gate folded into both), ReLU, H-pair max
                    {
                        // (32-bit shared-memory addresses: the generic pointer pair cost the N = 128 variants their last registers)
                        const uint32_t sc_u32 = wsc_u32 + ck * 128;
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            float4 a, c;
                            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w) : "r"(sc_u32 + q * 16));
                            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(c.x), "=f"(c.y), "=f"(c.z), "=f"(c.w) : "r"(sc_u32 + (BLOCK_N / 2) * 4 + q * 16));
#pragma unroll
                            for (int s = 0; s < NUM_SUB; ++s) {
                                v[s][4 * q + 0] = fmaf(v[s][4 * q + 0], a.x, c.x);
                                v[s][4 * q + 1] = fmaf(v[s][4 * q + 1], a.y, c.y);
                                v[s][4 * q + 2] = fmaf(v[s][4 * q + 2], a.z, c.z);
                                v[s][4 * q + 3] = fmaf(v[s][4 * q + 3], a.w, c.w);
                            }
                        }
                    }
                    if (p.se_partial && !p.sum_stored) {
                        // SELayer squeeze (models/handwritten_ctr_model.py:27-28) folded in: per-channel sum of the fp32 BN output
                        // over this warp's 32 pixels x NUM_SUB rows by a transpose-reduce butterfly (31 shuffles per 32 columns);
                        // afterwards lane L holds column n0+L. Slots are per (tile, warp): fixed-order final sum.
                        // (Sums of the values AS STORED - p.sum_stored - come from the staged bf16 chunk in the store loop below.)
                        float tsum[32];
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            float a = 0.f;
#pragma unroll
                            for (int s = 0; s < NUM_SUB; ++s) {
                                const bool ok = (w0 + s * p.sub_dw * kTileM + pix < p.W) && (h0 + s * p.sub_dh < p.out_H);
                                a += ok ? v[s][j] : 0.f;
                            }
                            tsum[j] = a;
                        }
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
                        const size_t slot = ((static_cast<size_t>(b) * p.h_tiles * p.w_tiles + static_cast<size_t>(h_tile) * p.w_tiles + w_tile) * 4 + quad);
                        p.se_partial[slot * p.N + n0 + lane] = tsum[0];
                    }
                    __nv_bfloat16* out = static_cast<__nv_bfloat16*>(p.out);
                    if (p.pool) {
                        const int ho = h_tile;
                        const int wq = w0 + quad * 32;                   // this warp's first pixel
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[0][j] = fmaxf(v[0][j], v[NUM_SUB - 1][j]);
                        if (p.relu) stage_chunk_row<true>(v[0], ebuf + lane * L::kEpiPitch);
                        else        stage_chunk_row<false>(v[0], ebuf + lane * L::kEpiPitch);
                        __syncwarp();
                        if (ho < p.out_H) {
                            __nv_bfloat16* obase = out + ((static_cast<size_t>(b) * p.out_H + ho) * p.W + wq) * p.N + n0;
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const int r = tr + 8 * i;
                                if (wq + r < p.W)
                                    *reinterpret_cast<uint4*>(obase + static_cast<size_t>(r) * p.N + tq * 8) =
                                        *reinterpret_cast<const uint4*>(ebuf + r * L::kEpiPitch + tq * 16);
                            }
                        }
                        __syncwarp();
                    } else {
#pragma unroll
                        for (int s = 0; s < NUM_SUB; ++s) {
                            const int w = w0 + s * p.sub_dw * kTileM + pix;
                            const int h = h0 + s * p.sub_dh;
                            if (w < p.W && h < p.out_H) {
                                const size_t off = p.out_line_pitch
                                    ? static_cast<size_t>(b) * p.out_line_pitch + (static_cast<size_t>(h) * p.W + w) * p.N + n0
                                    : ((static_cast<size_t>(b) * p.out_H + h) * p.W + w) * p.N + n0;
                                if (p.add) {
                                    const uint4* src = reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(p.add) + off);
#pragma unroll
                                    for (int q = 0; q < 4; ++q) {
                                        const uint4 a = ld_nc_v4(src + q);
                                        v[s][8 * q + 0] += bf16_lo(a.x); v[s][8 * q + 1] += bf16_hi(a.x);
                                        v[s][8 * q + 2] += bf16_lo(a.y); v[s][8 * q + 3] += bf16_hi(a.y);
                                        v[s][8 * q + 4] += bf16_lo(a.z); v[s][8 * q + 5] += bf16_hi(a.z);
                                        v[s][8 * q + 6] += bf16_lo(a.w); v[s][8 * q + 7] += bf16_hi(a.w);
                                    }
                                }
                            }
                            // pack, stage this lane's pixel row, write the chunk out as 64-byte runs
                            if (p.relu) stage_chunk_row<true>(v[s], ebuf + lane * L::kEpiPitch);
                            else        stage_chunk_row<false>(v[s], ebuf + lane * L::kEpiPitch);
                            __syncwarp();
                            // With p.sum_stored: channel sums (and sums of squares) of the chunk as stored, from the pieces read
                            // back for the stores - one slot per (line, ROW, 128-px span, warp quarter); rows beyond the tensor
                            // write zeros. The plain loop is kept apart so that layers without sums pay nothing for it.
                            const int wq = w0 + s * p.sub_dw * kTileM + quad * 32;           // this warp's first pixel
                            const size_t woff = p.out_line_pitch
                                ? static_cast<size_t>(b) * p.out_line_pitch + (static_cast<size_t>(h) * p.W + wq) * p.N + n0
                                : ((static_cast<size_t>(b) * p.out_H + h) * p.W + wq) * p.N + n0;
                            if (!(p.se_partial && p.sum_stored)) {
                                if (h < p.out_H) {
#pragma unroll
                                    for (int i = 0; i < 4; ++i) {
                                        const int r = tr + 8 * i;
                                        if (wq + r < p.W)
                                            *reinterpret_cast<uint4*>(out + woff + static_cast<size_t>(r) * p.N + tq * 8) =
                                                *reinterpret_cast<const uint4*>(ebuf + r * L::kEpiPitch + tq * 16);
                                    }
                                }
                            } else {
                                float cs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, cq[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                                if (h < p.out_H) {
#pragma unroll
                                    for (int i = 0; i < 4; ++i) {
                                        const int r = tr + 8 * i;
                                        if (wq + r < p.W) {
                                            const uint4 val = *reinterpret_cast<const uint4*>(ebuf + r * L::kEpiPitch + tq * 16);
                                            *reinterpret_cast<uint4*>(out + woff + static_cast<size_t>(r) * p.N + tq * 8) = val;
                                            piece_add(val, cs);
                                            if (p.sq_partial) piece_add_sq(val, cq);
                                        }
                                    }
                                }
                                const size_t slot = ((static_cast<size_t>(b) * (2 * p.h_tiles) + h) * p.w_tiles + w_tile) * 4 + quad;
                                piece_rows_reduce(cs);
                                if (lane < 4) piece_store(p.se_partial + slot * p.N + n0 + lane * 8, cs);
                                if (p.sq_partial) {
                                    piece_rows_reduce(cq);
                                    if (lane < 4) piece_store(p.sq_partial + slot * p.N + n0 + lane * 8, cq);
                                }
                            }
                            __syncwarp();
                        }
                    }
                } else {
                    // logits[b, w, n] = acc + bias[n]   (reference: nn.Linear, handwritten_ctr_model.py:169,175)
#pragma unroll
                    for (int s = 0; s < NUM_SUB; ++s) {
                        const int w = w0 + s * p.sub_dw * kTileM + pix;
                        if (w >= p.W) continue;
                        const size_t row = static_cast<size_t>(b) * p.W + w;
                        if (p.argmax_partial) {
                            // first maximum of the values as they are (or would be) stored; a NaN, once met, stays
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                float x = v[s][j] + __ldg(p.shift + min(n0 + j, p.N - 1));
                                if (p.out_dtype != HCTR_F32) x = __bfloat162float(__float2bfloat16_rn(x));
                                const bool take = (n0 + j < p.N) && (best_i[s] < 0 || x > best_v[s] || (x != x && best_v[s] == best_v[s]));
                                if (take) { best_v[s] = x; best_i[s] = n0 + j; }
                            }
                        }
                        if (p.out == nullptr) continue;                  // arg-max only: the logits are never written
                        if (p.lse_partial) {
                            // online (max, sum exp) over this thread's row: the row of D lives in one TMEM lane, so the
                            // softmax statistics need no cross-thread traffic. Values are taken as they will be stored.
                            float cm = run_m[s];
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                float x = v[s][j] + __ldg(p.shift + min(n0 + j, p.N - 1));
                                if (p.out_dtype != HCTR_F32) x = __bfloat162float(__float2bfloat16_rn(x));
                                if (n0 + j < p.N) cm = fmaxf(cm, x);
                            }
                            float acc = run_s[s] * __expf(run_m[s] - cm);
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                float x = v[s][j] + __ldg(p.shift + min(n0 + j, p.N - 1));
                                if (p.out_dtype != HCTR_F32) x = __bfloat162float(__float2bfloat16_rn(x));
                                if (n0 + j < p.N) acc += __expf(x - cm);
                            }
                            run_m[s] = cm; run_s[s] = acc;
                        }
                        if (p.out_dtype == HCTR_F32) {
                            float* dst = static_cast<float*>(p.out) + row * p.out_pitch + n0;
                            if (n0 + 32 <= p.N && (p.out_pitch & 3) == 0) {
#pragma unroll
                                for (int j = 0; j < 32; j += 4) {
                                    const float4 bs = __ldg(reinterpret_cast<const float4*>(p.shift + n0 + j));
                                    *reinterpret_cast<float4*>(dst + j) =
                                        make_float4(v[s][j] + bs.x, v[s][j + 1] + bs.y, v[s][j + 2] + bs.z, v[s][j + 3] + bs.w);
                                }
                            } else {
#pragma unroll
                                for (int j = 0; j < 32; ++j)
                                    if (n0 + j < p.N) dst[j] = v[s][j] + __ldg(p.shift + n0 + j);
                            }
                        } else {
                            __nv_bfloat16* dst = static_cast<__nv_bfloat16*>(p.out) + row * p.out_pitch + n0;
                            if (n0 + 32 <= p.N && (p.out_pitch & 7) == 0) {
                                uint32_t pk[16];
#pragma unroll
                                for (int j = 0; j < 32; j += 2)
                                    pk[j >> 1] = pack_bf16x2(v[s][j] + __ldg(p.shift + n0 + j),
                                                             v[s][j + 1] + __ldg(p.shift + n0 + j + 1));
                                uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
                                for (int q = 0; q < 4; ++q)
                                    d4[q] = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
                            } else {
#pragma unroll
                                for (int j = 0; j < 32; ++j)
                                    if (n0 + j < p.N) dst[j] = __float2bfloat16_rn(v[s][j] + __ldg(p.shift + n0 + j));
                            }
                        }
                    }
                }
            }
            if constexpr (EPI == EPI_LINEAR) {
                if (p.argmax_partial) {
#pragma unroll
                    for (int s = 0; s < NUM_SUB; ++s) {
                        const int w = w0 + s * p.sub_dw * kTileM + pix;
                        if (w < p.W)
                            p.argmax_partial[(static_cast<size_t>(b) * p.W + w) * (p.n_tiles * 2) + n_tile * 2 + half] =
                                make_int2(__float_as_int(best_v[s]), best_i[s]);
                    }
                }
                if (p.lse_partial) {
#pragma unroll
                    for (int s = 0; s < NUM_SUB; ++s) {
                        const int w = w0 + s * p.sub_dw * kTileM + pix;
                        if (w < p.W)
                            p.lse_partial[(static_cast<size_t>(b) * p.W + w) * (p.n_tiles * 2) + n_tile * 2 + half] =
                                make_float2(run_m[s], run_s[s]);
                    }
                }
            }
            // release this accumulator stage back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[acc]);
            if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

}  // namespace hctr
