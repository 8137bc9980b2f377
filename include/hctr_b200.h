/*
 * hctr_b200 — C ABI of the B200-native (sm_100a) HCTR recognition hot path.
 *
 * The reference (AndrewCullacino/handwritten-chinese-ocr-samples) has no FFI layer: its boundary is
 * the Python surface models/handwritten_ctr_model.py (hctr_model), utils/ctc_codec.py (ctc_codec) and
 * the CTCLoss call in main.py. Each entry point below names the reference call it stands in for
 * (file:line under /root/reference). The host-side Python mirrors of those objects live in
 * handwritten-chinese-ocr-samples_b200/{models,utils}/ and bind this library through ctypes;
 * INTEGRATION.md shows the stub.
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error; hctr_last_error() gives a thread-local message.
 *   - all pointers are DEVICE pointers unless the name says host_; the library borrows them for the
 *     duration of the call and allocates nothing persistent.
 *   - `stream` is a cudaStream_t (pass torch.cuda.current_stream().cuda_stream); launches are async.
 *   - activations are NHWC bf16 ([B][H][W][C], C innermost); parameters are fp32 unless packed.
 *   - there is no CPU fallback: without an sm_100a device the calls fail.
 */
#ifndef HCTR_B200_H_
#define HCTR_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HCTR_ABI_VERSION 1

/* error codes */
#define HCTR_OK 0
#define HCTR_ERR_INVALID (-1)      /* bad argument: maps to ValueError / RuntimeError(shape) */
#define HCTR_ERR_CUDA (-2)         /* CUDA failure: RuntimeError */
#define HCTR_ERR_UNSUPPORTED (-3)  /* not an sm_100a device */
#define HCTR_ERR_INDEX (-4)        /* reference raises IndexError (utils/ctc_codec.py:139,198) */

/* element types of logit tensors */
#define HCTR_F32 0
#define HCTR_BF16 1

const char* hctr_last_error(void);
int hctr_abi_version(void);
/* 0 iff `device` can run these kernels (compute capability 10.x). */
int hctr_device_supported(int device);

/* ---- backbone ------------------------------------------------------------------------------- */

/* cnn.conv0_1 + bn0_1 + relu (models/handwritten_ctr_model.py:116-118).
 * x: fp32 [B][1][H][W] (the reference input, [-1,1]); w: fp32 [64][9] (OIHW flattened);
 * scale/shift: fp32 [64], y = relu(conv(x)*scale + shift) with the conv bias and eval-mode BN folded in
 * (relu=0, scale=1, shift=bias gives the raw conv output the train-mode BN needs); y: bf16 NHWC [B][H][W][64]. */
int hctr_stem_conv_fwd(const float* x, const float* w, const float* scale, const float* shift, void* y, int B, int H,
                       int W, int relu, void* stream);

/* 3x3 (pad 1) or 1x1 convolution + per-channel fp32 scale/shift (+ReLU) (+(2,1) max-pool over H pairs):
 * conv0_2/bn0_2 (:119-123), BasicBlock conv1/bn1/relu and conv2/bn2 (:49-53), downsample (:55-57, :104-107),
 * cnn.convS/bnS/relu/max_pool2d (:126-129,133-136,140-143,147-150).
 * x: bf16 NHWC [B][H][W][Cin]; w_packed: bf16 [Cout][ksize*ksize][Cin]; y: bf16 NHWC [B][H or H/2][W][Cout].
 * tcgen05 implicit GEMM, TMA-fed; Cin % 64 == 0; Cout in {64,128} or a multiple of 256. */
int hctr_conv_bn_act_fwd(const void* x, const void* w_packed, const float* scale, const float* shift, void* y, int B,
                         int H, int W, int Cin, int Cout, int ksize, int relu, int pool, void* stream);

/* BasicBlock conv2 + bn2 (:52-53) with the SELayer squeeze (:27-28) folded into the epilogue: besides y (bf16 NHWC,
 * no ReLU) the kernel writes per-(tile, warp) channel sums of the fp32 BN output to se_partial
 * [B][hctr_conv_se_slices(H,W,Cout)][Cout]; hctr_se_excite(se_partial, slices = hctr_conv_se_slices(H,W,Cout), ...) finishes
 * the mean in a fixed order (deterministic). Saves one full read of the activation per residual block. */
int hctr_conv_se_slices(int H, int W, int Cout);
int hctr_conv_bn_se_fwd(const void* x, const void* w_packed, const float* scale, const float* shift, void* y,
                        float* se_partial, int B, int H, int W, int Cin, int Cout, int ksize, void* stream);

/* The residual block without the intermediate tensor (BasicBlock.forward, models/handwritten_ctr_model.py:47-58):
 * the SE gate needs mean_hw(bn2(conv2(t))), which is linear in t - sum_hw z[co] = scale[co] * sum_{tap,ci} W[co,tap,ci] *
 * S_tap[ci] + HW*shift[co] with S_tap = the sum of t over the pixels tap (dh,dw) reads - so it is computed BEFORE conv2:
 *   hctr_conv_bn_act_sum_fwd   conv1 + bn1 + relu, plus per-(tile, warp) channel sums of the STORED bf16 output t:
 *                              partial fp32 [B][hctr_conv_sum_slices(H,W,Cin,Cout,ksize)][Cout]
 *   hctr_se_gate_from_input    totals + border rows/columns of t + a [C x 9C] mat-vec with conv2's packed weights ->
 *                              mean z -> SELayer.fc (:19-24) -> gate fp32 [B][C]
 *   hctr_conv_bn_gate_res_fwd  conv2 + bn2, times gate[b,c], plus the residual, ReLU - one write, no re-read. */
int hctr_conv_sum_slices(int H, int W, int Cin, int Cout, int ksize);
int hctr_conv_bn_act_sum_fwd(const void* x, const void* w_packed, const float* scale, const float* shift, void* y,
                             float* partial, int B, int H, int W, int Cin, int Cout, int ksize, int relu, void* stream);
/* Train-mode convolution with the BatchNorm batch statistics taken in its epilogue (nn.BatchNorm2d in train(),
 * models/handwritten_ctr_model.py:38,40,74-92 - F.batch_norm's mean / biased variance over (B,H,W)): z = conv(x)*scale + shift
 * (bf16 NHWC, no ReLU) plus per-(tile row, 128-px span, warp quarter) sums of z and z*z AS STORED (bf16-rounded), both
 * fp32 [B][hctr_conv_sum_slices(H,W,Cin,Cout,ksize)][Cout] - the layout hctr_chan_stats produces with its own slice count, so
 * hctr_bn_finalize_train(psum, psq, B, slices = hctr_conv_sum_slices(...), ...) finishes them. Replaces one full read of z. */
int hctr_conv_stats_fwd(const void* x, const void* w_packed, const float* scale, const float* shift, void* y, float* psum,
                        float* psq, int B, int H, int W, int Cin, int Cout, int ksize, void* stream);
long long hctr_se_gate_workspace_bytes(int B, int C);
int hctr_se_gate_from_input(const void* t, const float* partial, int slices, const void* conv_w_packed, const float* scale,
                            const float* shift, const float* w1, const float* w2, float* gate, int B, int H, int W, int C,
                            int Cr, void* workspace, long long workspace_bytes, void* stream);
int hctr_conv_bn_gate_res_fwd(const void* x, const void* w_packed, const float* scale, const float* shift, const float* gate,
                              const void* residual, void* y, int B, int H, int W, int Cin, int Cout, int ksize, int relu,
                              void* stream);

/* SELayer squeeze (:27-28): deterministic two-stage mean over (H,W) incl. padded columns.
 * x: bf16 NHWC; partial: fp32 workspace [B][slices][C]; the second stage runs inside hctr_se_excite.
 * `slices` must equal hctr_se_slices(H, W). */
int hctr_se_slices(int H, int W);
int hctr_se_squeeze(const void* x, float* partial, int B, int H, int W, int C, void* stream);
/* SELayer excite (:19-24,29): gate = sigmoid(W2 . relu(W1 . mean)); w1: fp32 [C/r][C]; w2: fp32 [C][C/r];
 * gate: fp32 [B][C]. */
int hctr_se_excite(const float* partial, int slices, const float* w1, const float* w2, float* gate, int B, int C,
                   int Cr, int HW, void* stream);
/* BasicBlock tail (:30,54-58): y = relu(x * gate[b,c] + residual); all bf16 NHWC, same shape. */
int hctr_se_scale_residual_relu(const void* x, const float* gate, const void* residual, void* y, int B, int H, int W,
                                int C, void* stream);

/* Column classifier, hctr_model.forward (:172-176): feat: bf16 NHWC [B][Hf][W][Cf] (Hf*Cf = 2048);
 * w_packed: bf16 [num_classes][Hf*Cf] with k = h*Cf + c (the reference's flatten gives d = c*Hf + h);
 * logits: [B][W][out_pitch] of out_dtype, only the first num_classes columns of a row are written. The
 * reference's [W,B,C] result is the (1,0,2) permuted view of it. */
int hctr_classifier_fwd(const void* feat, const void* w_packed, const float* bias, void* logits, int out_dtype,
                        long long out_pitch, int B, int Hf, int W, int Cf, int num_classes, void* stream);

/* The same classifier GEMM fused with log_softmax: the epilogue keeps an online (max, sum exp) per logits row and
 * column half-tile (a row of the accumulator lives in one TMEM lane, so no cross-thread traffic), a fix-up kernel combines
 * them into row_lse fp32 [B][W]; log_softmax(logits)[b,w,c] = logits[b,w,c] - row_lse[b,w] (reference: main.py:406
 * `preds.log_softmax(2)`). workspace: hctr_classifier_lse_workspace_bytes(B, W, num_classes), 16-byte aligned. */
int hctr_classifier_lse_fwd(const void* feat, const void* w_packed, const float* bias, void* logits, int out_dtype,
                            long long out_pitch, int B, int Hf, int W, int Cf, int num_classes, float* row_lse,
                            void* workspace, long long workspace_bytes, void* stream);
long long hctr_classifier_lse_workspace_bytes(int B, int W, int num_classes);

/* The same classifier GEMM fused with greedy decoding (hctr_model.forward :172-176 followed by
 * ctc_codec.decode -> __greedy_search__, utils/ctc_codec.py:63-99): the epilogue keeps the first maximum per logits row
 * and column half-tile of the values AS THEY ARE STORED in out_dtype (numpy.argmax semantics: ties -> lowest index, the first
 * NaN beats everything), a fix-up kernel combines them into argmax_bt int32 [B][W], then blank / unknown / repeats are
 * dropped exactly as hctr_ctc_greedy_decode does: out_idx int32 [B][W], out_len int32 [B]. logits may be NULL: the 1.93 GB
 * of bf16 logits of a 64-line batch are then never written (nor read back by the arg-max pass); out_dtype still says how a
 * value would have been rounded, so the result equals hctr_classifier_fwd + hctr_ctc_greedy_decode bit for bit.
 * workspace: hctr_classifier_greedy_workspace_bytes(B, W, num_classes), 16-byte aligned. */
int hctr_classifier_greedy_fwd(const void* feat, const void* w_packed, const float* bias, void* logits, int out_dtype,
                               long long out_pitch, int B, int Hf, int W, int Cf, int num_classes, int32_t* argmax_bt,
                               int32_t* out_idx, int32_t* out_len, void* workspace, long long workspace_bytes, void* stream);
long long hctr_classifier_greedy_workspace_bytes(int B, int W, int num_classes);

/* ---- CTC codec ------------------------------------------------------------------------------ */

/* ctc_codec.__greedy_search__ (utils/ctc_codec.py:70-99): per (t,b) argmax over C (ties -> lowest index,
 * NaN -> first NaN, as numpy.argmax), then drop blank (0), unknown (C-1) and repeats of the raw previous index.
 * logits element (t,b,c) is at logits[t*stride_t + b*stride_b + c] (elements of `dtype`).
 * argmax_out: int32 [B][T] raw per-step argmax (required scratch/result); out_idx: int32 [B][T] compacted label indices;
 * out_len: int32 [B]. */
int hctr_ctc_greedy_decode(const void* logits, int dtype, int T, int B, int C, long long stride_t, long long stride_b,
                           int32_t* argmax_out, int32_t* out_idx, int32_t* out_len, void* stream);

/* The second half of hctr_ctc_greedy_decode alone: argmax_bt int32 [B][T] raw per-step classes -> out_idx / out_len. */
int hctr_ctc_collapse(const int32_t* argmax_bt, int T, int B, int C, int32_t* out_idx, int32_t* out_len, void* stream);

/* log_softmax over C (scipy.special.log_softmax, utils/ctc_codec.py:65) fused with the per-step top-k
 * (np.argsort flip, :186). topk_idx: int32 [T][B][k] descending by log-prob (ties -> lower index first);
 * topk_logp: fp32 [T][B][k]; lse: fp32 [T][B] (logp = logit - lse). k <= 16. */
int hctr_ctc_topk_logsoftmax(const void* logits, int dtype, int T, int B, int C, long long stride_t,
                             long long stride_b, int k, int32_t* topk_idx, float* topk_logp, float* lse, void* stream);

/* ctc_codec.__cbs_full__ + __context_beam_search__ + Beam (utils/ctc_codec.py:183-285,288-307) without a
 * transformer; the language model is a per-class unigram table (lm_table[c], fp64, may be NULL = zero LM),
 * scored over prefix + look-ahead suffix exactly as `ngram.score(' '.join(prefix+suffix))` would for a
 * unigram model. float64 accumulators, stable ordering by insertion. One CTA per sequence.
 * Inputs are the outputs of hctr_ctc_topk_logsoftmax (k <= 16, beam_size <= 16). out_idx: int32 [B][T]; out_len: int32 [B];
 * status: int32 [B], 0 ok, HCTR_ERR_INDEX if the greedy path is empty (the reference raises IndexError). */
int hctr_ctc_prefix_beam_search(const int32_t* topk_idx, const float* topk_logp, int T, int B, int C, int k,
                                int beam_size, double lm_penalty, double len_bonus, const double* lm_table,
                                int32_t* out_idx, int32_t* out_len, int32_t* status, void* workspace,
                                long long workspace_bytes, void* stream);
/* bytes of caller-owned device scratch (greedy look-ahead list + prefix trie) for the call above */
long long hctr_ctc_beam_workspace_bytes(int T, int B, int beam_size);

/* Back-off n-gram language model for the beam search (reference: kenlm.Model(ngram_path).score(' '.join(prefix +
 * suffix), eos=False), utils/ctc_codec.py:120-122,276-279; a 5-gram from lmplz, third-party/README.md:28-42). Words are
 * single characters = class indices; ids: classes 0..C-1, then <s>, </s>, <unk>. One open-addressing hash table over all
 * orders (linear probing, capacity a power of two): entry i = uint4 {key_lo low, key_lo high, key_hi, float bits of
 * log10 p}, key_lo = r0 | r1<<16 | r2<<32 | r3<<48 and key_hi = r4 | order<<16 with r0 the most recent word; key_hi == 0
 * marks an empty slot; slot of a key = splitmix64(key_lo ^ key_hi * 0x9E3779B97F4A7C15) & mask. All device pointers.
 * hctr_b200/ngram_lm.py builds it from an ARPA file. */
typedef struct hctr_ngram_lm {
    const void* entries;          /* uint4 [mask + 1] */
    const float* backoff;         /* [mask + 1] log10 back-off weight of the n-gram in the same slot */
    const int32_t* vocab;         /* [num_ids] id -> LM word id: itself, or unk_id for a word without a unigram */
    unsigned long long mask;      /* capacity - 1 */
    int order;                    /* 1..5 */
    int bos_id, unk_id, num_ids;
} hctr_ngram_lm;
/* kenlm.Model.score(sentence, bos=True, eos=False) for nseq sequences of class indices ids[offsets[q] .. offsets[q+1]):
 * float32 accumulation as in KenLM; out: fp32 [nseq]. One thread per sequence. */
int hctr_ngram_score(const hctr_ngram_lm* lm, const int32_t* ids, const int32_t* offsets, int nseq, float* out, void* stream);
/* hctr_ctc_prefix_beam_search with the n-gram model above as the language model (lm_table must be NULL then). */
int hctr_ctc_prefix_beam_search_lm(const int32_t* topk_idx, const float* topk_logp, int T, int B, int C, int k,
                                   int beam_size, double lm_penalty, double len_bonus, const double* lm_table,
                                   const hctr_ngram_lm* ngram, int32_t* out_idx, int32_t* out_len, int32_t* status,
                                   void* workspace, long long workspace_bytes, void* stream);

/* ctc_codec.__cbs_skip__ (utils/ctc_codec.py:124-181): candidates per step are the classes with log-prob > log(0.001)
 * in index order; a single candidate takes the reference's in-place fast path (quirks included), otherwise a
 * duplicate-aware context beam search runs over the candidates. Fused log-softmax/prune pre-pass + one CTA per sequence.
 * status[b]: 0 ok; HCTR_ERR_INDEX where the reference raises IndexError (empty greedy path :139, or a step without any
 * candidate :179); HCTR_ERR_UNSUPPORTED if a step has more than 128 candidates (see hctr_ctc_skip_beam_search_ex below). */
int hctr_ctc_skip_beam_search(const void* logits, int dtype, int T, int B, int C, long long stride_t, long long stride_b,
                              int beam_size, double lm_penalty, double len_bonus, const double* lm_table, int32_t* out_idx,
                              int32_t* out_len, int32_t* status, void* workspace, long long workspace_bytes, void* stream);
/* the same with the back-off n-gram model as the language model (lm_table must be NULL then) */
int hctr_ctc_skip_beam_search_lm(const void* logits, int dtype, int T, int B, int C, long long stride_t, long long stride_b,
                                 int beam_size, double lm_penalty, double len_bonus, const double* lm_table,
                                 const hctr_ngram_lm* ngram, int32_t* out_idx, int32_t* out_len, int32_t* status,
                                 void* workspace, long long workspace_bytes, void* stream);
long long hctr_ctc_skip_workspace_bytes(int T, int B);
/* Candidates per step: the reference takes every class with p > 0.001 (utils/ctc_codec.py:144), at most 999. The entry points
 * above hold up to 128 per step (small candidate tables, a step's dict entries in shared memory) and report
 * HCTR_ERR_UNSUPPORTED in status[b] for a sequence that needs more; the caller then repeats the call through
 * hctr_ctc_skip_beam_search_ex with max_candidates = hctr_ctc_skip_max_candidates() (= 1024: 8 KB of candidate table per
 * row, dict entries in the workspace) and a workspace of hctr_ctc_skip_workspace_bytes_ex(T, B, max_candidates).
 * max_candidates must be 128 or 1024; ngram may be NULL. */
int hctr_ctc_skip_max_candidates(void);
long long hctr_ctc_skip_workspace_bytes_ex(int T, int B, int max_candidates);
int hctr_ctc_skip_beam_search_ex(const void* logits, int dtype, int T, int B, int C, long long stride_t, long long stride_b,
                                 int beam_size, double lm_penalty, double len_bonus, const double* lm_table,
                                 const hctr_ngram_lm* ngram, int max_candidates, int32_t* out_idx, int32_t* out_len,
                                 int32_t* status, void* workspace, long long workspace_bytes, void* stream);

/* CTCLoss(blank=0, reduction='mean', zero_infinity=True) on log_softmax(logits) and its gradient wrt the
 * logits (main.py:205,406-409,426). logits element (t,b,c) at logits[t*stride_t + b*stride_b + c];
 * targets: int32 concatenated [sum L]; target_lengths/input_lengths: int32 [B].
 * nll: fp32 [B] per-sequence negative log-likelihood (inf -> 0 when zero_infinity); loss: fp32 [1] =
 * mean_b(nll_b / max(L_b,1)); grad (may be NULL): same dtype/strides as logits, d loss / d logits scaled by
 * grad_scale. row_lse (may be NULL): fp32 [B][T] log-sum-exp of every logits row, e.g. from hctr_classifier_lse_fwd -
 * then the first pass over the logits only gathers the label columns. */
int hctr_ctc_loss_fwd_bwd(const void* logits, int dtype, int T, int B, int C, long long stride_t, long long stride_b,
                          const int32_t* targets, const int32_t* target_lengths, const int32_t* input_lengths,
                          int max_target_len, const float* row_lse, float* nll, float* loss, void* grad, float grad_scale,
                          void* workspace, long long workspace_bytes, void* stream);
long long hctr_ctc_loss_workspace_bytes(int T, int B, int max_target_len);
/* Byte offset inside that workspace of an int32 [B] array that, after the call, holds 1 for every sequence that was
 * recomputed by the log-space recursion instead of the scaled linear one (diagnostics / tests). */
long long hctr_ctc_loss_flag_offset(int T, int B, int max_target_len);

/* ---- training (train()-mode forward and the backward pass; reference: main.py:367,383-438) -------------------- */

/* Per-(line, slice, channel) sum and sum of squares of an NHWC bf16 tensor (fixed order, deterministic): the batch
 * statistics of nn.BatchNorm2d in train() (models/handwritten_ctr_model.py:38,40,74-92) and the SE squeeze (:27-28)
 * come from the same pass. psum/psq: fp32 [B][hctr_stat_slices(B,H,W)][C]; psq may be NULL. The slice size adapts to B
 * so that even 2 lines per GPU fill the SMs. */
int hctr_stat_slices(int B, int H, int W);
int hctr_chan_stats(const void* x, float* psum, float* psq, int B, int H, int W, int C, void* stream);
/* Batch mean / biased variance -> invstd, scale = gamma*invstd, shift = beta - mean*scale; updates running stats with
 * `momentum` and the unbiased variance (pass NULL to skip); line_sum: fp32 [B][C] = sum over (h,w) (may be NULL).
 * The partials are consumed: with more than 64 slices per line (the conv epilogue's slots, hctr_conv_stats_fwd) each line's
 * partials are first added into its slice 0 in place. */
int hctr_bn_finalize_train(const float* psum, const float* psq, int B, int slices, int C, int HW, const float* gamma,
                           const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                           float* mean, float* invstd, float* scale, float* shift, float* line_sum, void* stream);
/* SE gate from BN-folded line means (train mode): m = scale*mean_hw(z)+shift, hidden = relu(W1 m), gate = sigmoid(W2 hidden);
 * se_mean [B][C], hidden [B][Cr], gate [B][C] are kept for the backward. */
int hctr_se_excite_train(const float* line_sum, const float* scale, const float* shift, const float* w1, const float* w2,
                         float* se_mean, float* hidden, float* gate, int B, int C, int Cr, int HW, void* stream);
/* out = dropout(pool(relu((z*scale+shift) [*gate] [+res]))) - BN apply, SE scale, residual, ReLU, (2,1) max-pool and
 * counter-based dropout (keep-scale 1/(1-p)) in one pass. z/res: bf16 NHWC [B][H][W][C]; out: [B][H or H/2][W][C];
 * mask (may be NULL when no backward follows): uint8 [B][H][W][C/8], one keep-bit per INPUT element with the ReLU sign,
 * the pool winner and the dropout keep folded together - all the backward needs of this pass. */
int hctr_train_apply_fwd(const void* z, const float* scale, const float* shift, const float* gate, const void* res,
                         void* out, void* mask, int B, int H, int W, int C, int relu, int pool, float drop_p, unsigned seed,
                         void* stream);
/* Backward of the pass above + BatchNorm (+SE) backward in three steps; with d_pre = mask ? dout/(1-p) : 0:
 * reduce -> per-(b,slice,c) sums of d_pre and d_pre*z; finalize -> dgamma, dbeta, conv-bias gradient, SE FC gradients and
 * the coefficients P[B][C], Q[B][C], R[C]; apply -> dz = P*d_pre + Q + R*z, dres = d_pre. */
int hctr_train_bwd_reduce(const void* dout, const void* z, const void* mask, float* pA2, float* pA3, int B, int H, int W,
                          int C, int pool, float drop_p, void* stream);
/* (finalize consumes the per-slice partials: for layers with an SE gate slice 0 of pA2/pA3 is overwritten with the totals;
 *  C must be a multiple of 32) */
int hctr_train_bwd_finalize(const float* pA2, const float* pA3, int slices, int B, int C, int HW, const float* gamma,
                            const float* mean, const float* invstd, const float* scale, const float* shift,
                            const float* line_sum, const float* gate, const float* se_hidden, const float* se_mean,
                            const float* w1, const float* w2, int Cr, float* dw1, float* dw2, float* dgamma,
                            float* dbeta, float* dbias, float* P, float* Q, float* R, void* stream);
int hctr_train_bwd_apply(const void* dout, const void* z, const void* mask, const float* P, const float* Q, const float* R,
                         void* dz, void* dres, int B, int H, int W, int C, int pool, float drop_p, void* stream);

/* Data gradient of a 3x3/1x1 convolution: the same tcgen05 implicit GEMM with mirrored taps. dz: bf16 NHWC
 * [B][H][W][Cout]; w_packed_t: bf16 [Cin][k*k][Cout]; ones/zeros: fp32 [Cin]; add: optional bf16 [B][H][W][Cin]
 * accumulated into the result (residual branch); dx: bf16 [B][H][W][Cin]. */
int hctr_conv_dgrad(const void* dz, const void* w_packed_t, const float* ones, const float* zeros, const void* add,
                    void* dx, int B, int H, int W, int Cout, int Cin, int ksize, void* stream);
/* Weight gradient on tcgen05 (K = pixels, MN-major operands straight from the NHWC tensors, split-K with a
 * fixed-order reduction). dw: fp32 [Cout][Cin][k][k] (the reference's OIHW parameter layout). */
int hctr_conv_wgrad(const void* dz, const void* x, float* dw, int B, int H, int W, int Cout, int Cin, int ksize,
                    void* workspace, long long workspace_bytes, void* stream);
long long hctr_wgrad_workspace_bytes(int B, int H, int W, int M, int N, int ntaps);
/* Classifier backward: dfeat[b,h,w,c] = sum_n dlogits[b,w,n]*W[n, c*Hf+h] (w_t: bf16 [Hf*Cf][pitch], row k = h*Cf+c,
 * zero padded) and dW[n][c*Hf+h] (fp32, reference layout); dlogits: bf16 [B][W][pitch]. */
int hctr_classifier_dgrad(const void* dlogits, long long pitch, const void* w_t, const float* ones, const float* zeros,
                          void* dfeat, int B, int Hf, int W, int Cf, int num_classes, void* stream);
int hctr_linear_wgrad(const void* dlogits, long long pitch, const void* feat, float* dw, int B, int Hf, int W, int Cf,
                      int num_classes, void* workspace, long long workspace_bytes, void* stream);
long long hctr_linear_wgrad_workspace_bytes(int B, int Hf, int W, int Cf, int num_classes);
/* Column sums of a bf16 [rows][pitch] matrix (classifier bias gradient), two-stage fixed order. */
int hctr_colsum_bf16(const void* x, long long rows, int cols, long long pitch, float* out, float* workspace,
                     long long workspace_bytes, void* stream);
long long hctr_colsum_workspace_bytes(long long rows, int cols);
/* Stem (Cin=1) weight gradient: dw fp32 [64][9] from dz bf16 [B][H][W][64] and the fp32 input image. */
int hctr_stem_wgrad(const void* dz, const float* x, float* dw, int B, int H, int W, float* workspace,
                    long long workspace_bytes, void* stream);
long long hctr_stem_wgrad_workspace_bytes(int B, int H, int W);
/* Optimizer tail over flat fp32 buffers (main.py:210-213,430-438): total_norm = ||grad*grad_scale||_2,
 * coef = min(1, max_norm/(total_norm+1e-6)) (clip_grad_norm_), then SGD with momentum and weight decay as
 * torch.optim.SGD. norm_out: fp32 [4] = {total_norm, coef, skipped, unused}: a non-finite total_norm sets skipped = 1 and
 * leaves parameters and momentum untouched (main.py:413 `if not torch.isfinite(loss): continue`, GradScaler.step);
 * workspace: hctr_sgd_workspace_bytes(). */
int hctr_sgd_clip_step(float* params, const float* grads, float* momentum_buf, long long n, float grad_scale,
                       float max_norm, float lr, float momentum, float weight_decay, int first_step, float* norm_out,
                       float* workspace, void* stream);
long long hctr_sgd_workspace_bytes(void);

/* All weight operands of a training step in one launch (the optimizer rewrites every parameter each step): fp32 parameters
 * in the reference layout - conv OIHW [Cout][Cin][kh*kw], nn.Linear [N][Cf*Hf] read as Cout = N, Cin = Cf, taps = Hf
 * (models/handwritten_ctr_model.py:37-40,169) - to the bf16 layouts hctr_conv_bn_act_fwd / hctr_classifier_fwd (dst_fwd:
 * [Cout][taps][Cin]) and hctr_conv_dgrad / hctr_classifier_dgrad (dst_bwd, may be NULL: bwd_mode 0 = [Cin][taps][Cout],
 * bwd_mode 1 = [taps][Cin][bwd_pitch], columns >= Cout untouched) read. taps is 1, 4 or 9. The descriptor array lives in DEVICE
 * memory; tile_start = running sum of ceil(cout/32)*ceil(cin/32) over the preceding descriptors, total_tiles = that sum. */
typedef struct hctr_pack_desc {
    const float* src;
    void* dst_fwd;
    void* dst_bwd;
    int cout, cin, taps, bwd_mode;
    long long bwd_pitch;
    long long tile_start;
} hctr_pack_desc;
int hctr_pack_weights(const hctr_pack_desc* descs_device, int ndesc, long long total_tiles, void* stream);

/* ---- input pipeline (next to the path: utils/dataset.py:78-93 NormalizePAD, test.py:170-186) ----------------- */

/* pixels: uint8 grayscale lines of height H, line b stored row-major [H][widths[b]] at pixels + offsets[b] (device);
 * out: fp32 [B][1][H][Wb] = ((p/255) - 0.5) / 0.5, columns >= widths[b] replicate the last real column. Bit-exact with
 * torchvision ToTensor + sub_(0.5).div_(0.5). Requires widths[b] <= Wb. */
int hctr_normalize_pad(const void* pixels, const long long* offsets, const int32_t* widths, float* out, int B, int H,
                       int Wb, void* stream);

/* cv2.resize(line, (dst_w, dst_h), interpolation=cv2.INTER_AREA) for one uint8 grayscale line (device pointers, row pitches in
 * bytes): the reference's resize to height 128 in front of NormalizePAD (utils/dataset.py:53-57, test.py:206-214; the caller
 * computes dst_w = int(src_w * dst_h / src_h) as the reference does). Bit-identical to OpenCV's C++ paths: area-weighted
 * mean (float32, table order) when both scale factors are >= 1, integer-sum fast path for integer factors, otherwise the
 * bilinear fixed-point path with INTER_AREA's coefficient rule. */
int hctr_resize_area_u8(const void* src, int src_h, int src_w, long long src_pitch, void* dst, int dst_h, int dst_w,
                        long long dst_pitch, void* stream);

/* ---- evaluation (the step after decode: main.py:497-517, test.py:266-286) -------------------------------------- */

/* Levenshtein distance of each decoded label sequence to its ground truth (editdistance.eval(pre, tru)); CER =
 * sum(dist) / sum(target_lengths). pred_idx: int32 [B][pred_pitch] (the decode output), pred_len: int32 [B]
 * (all <= max_pred_len), targets: int32 concatenated labels (ctc_codec.encode), target_lengths: int32 [B]; dist: int32 [B]. */
int hctr_edit_distance(const int32_t* pred_idx, const int32_t* pred_len, int pred_pitch, int max_pred_len,
                       const int32_t* targets, const int32_t* target_lengths, int B, int32_t* dist, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HCTR_B200_H_ */
