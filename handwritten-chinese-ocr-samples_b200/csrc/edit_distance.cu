// Character error rate on device: Levenshtein distance between every decoded label sequence and its
// ground truth, so that evaluation needs no per-sample Python loop. Reference: editdistance.eval(pre, tru) and
// CER = total / nchars in main.py:506-517 and test.py:275-286.
// One CTA per (prediction, target) pair. Rows = target characters; a row is updated in two parallel steps:
//   tmp[j] = min(prev[j] + 1, prev[j-1] + (a_i != b_j))                 (deletion / substitution, elementwise)
//   cur[j] = min(tmp[j], cur[j-1] + 1) = j + min_{k<=j}(tmp[k] - k)     (insertions as a block-wide prefix-min)
#include "common.cuh"
#include "../../include/hctr_b200.h"

namespace hctr {

constexpr int kEdThreads = 256;

__global__ void __launch_bounds__(kEdThreads)
edit_distance_kernel(const int32_t* __restrict__ pred, const int32_t* __restrict__ pred_len, int pred_pitch,
                     const int32_t* __restrict__ tgt, const int32_t* __restrict__ tgt_len, int32_t* __restrict__ dist) {
    extern __shared__ int rows[];                        // prev[n+1], cur[n+1]
    __shared__ int warp_min[kEdThreads / 32];
    __shared__ int s_off;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        int off = 0;
        for (int i = 0; i < b; ++i) off += tgt_len[i];
        s_off = off;
    }
    __syncthreads();
    const int n = pred_len[b], m = tgt_len[b];
    const int32_t* a = tgt + s_off;                      // target (rows)
    const int32_t* p = pred + (long long)b * pred_pitch; // prediction (columns)
    int* prev = rows;
    int* cur = rows + (n + 1);
    for (int j = tid; j <= n; j += kEdThreads) prev[j] = j;
    __syncthreads();
    const int chunk = (n + 1 + kEdThreads - 1) / kEdThreads;
    const int j0 = tid * chunk, j1 = min(n + 1, j0 + chunk);
    for (int i = 1; i <= m; ++i) {
        const int ai = a[i - 1];
        // local pass: v[j] = tmp[j] - j, running min inside the chunk
        int run = 0x3fffffff;
        for (int j = j0; j < j1; ++j) {
            int tmp;
            if (j == 0) tmp = i;
            else tmp = min(prev[j] + 1, prev[j - 1] + (ai != p[j - 1] ? 1 : 0));
            run = min(run, tmp - j);
            cur[j] = run;                                // chunk-local prefix min of v
        }
        // exclusive prefix-min of the chunk totals across threads
        int incl = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t2 = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl = min(incl, t2);
        }
        if (lane == 31) warp_min[warp] = incl;
        __syncthreads();
        int carry = 0x3fffffff;
        for (int w2 = 0; w2 < warp; ++w2) carry = min(carry, warp_min[w2]);
        const int excl_lane = __shfl_up_sync(0xffffffffu, incl, 1);
        if (lane > 0) carry = min(carry, excl_lane);
        for (int j = j0; j < j1; ++j) cur[j] = min(cur[j], carry) + j;
        __syncthreads();
        int* t3 = prev; prev = cur; cur = t3;
    }
    if (tid == 0) dist[b] = prev[n];
}

}  // namespace hctr

using namespace hctr;

extern "C" int hctr_edit_distance(const int32_t* pred_idx, const int32_t* pred_len, int pred_pitch, int max_pred_len,
                                  const int32_t* targets, const int32_t* target_lengths, int B, int32_t* dist,
                                  void* stream) {
    HCTR_CHECK(pred_idx && pred_len && target_lengths && dist, HCTR_ERR_INVALID, "edit_distance: null pointer");
    HCTR_CHECK(B >= 0 && max_pred_len >= 0 && max_pred_len <= pred_pitch, HCTR_ERR_INVALID, "edit_distance: bad shape");
    if (B == 0) return HCTR_OK;
    const size_t smem = 2 * (size_t)(max_pred_len + 1) * sizeof(int);
    HCTR_CHECK(smem <= 200 * 1024, HCTR_ERR_INVALID, "edit_distance: prediction too long (%d)", max_pred_len);
    static PerDeviceOnce once;
    int dev;
    if (once.need(dev)) {
        HCTR_CUDA(cudaFuncSetAttribute(edit_distance_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        once.mark(dev);
    }
    edit_distance_kernel<<<B, kEdThreads, smem, static_cast<cudaStream_t>(stream)>>>(pred_idx, pred_len, pred_pitch, targets,
                                                                                  target_lengths, dist);
    HCTR_CUDA(cudaGetLastError());
    return HCTR_OK;
}
