/*
 * hctr_b200 - test hooks. NOT part of the product ABI (include/hctr_b200.h): nothing in the drop-in Python surface calls
 * these. They exist so that tests/ can run the same convolution on every kernel variant inside one process.
 * Each hook refuses (HCTR_ERR_UNSUPPORTED) unless HCTR_TEST_HOOKS=1 is set in the environment.
 */
#ifndef HCTR_B200_TESTING_H_
#define HCTR_B200_TESTING_H_

#ifdef __cplusplus
extern "C" {
#endif

/* Kernel variant of the following hctr_conv_* launches of this process (reference op: nn.Conv2d,
 * models/handwritten_ctr_model.py:37-40): kw_fused_slab 1 = one 136-pixel activation box per (kh, 64-channel chunk) serves the
 * three kw taps (default), 0 = one TMA box per tap; cta_pairs 1 = wide layers on tcgen05.mma.cta_group::2 CTA pairs
 * (default), 0 = single-CTA kernel. The same switches exist as HCTR_IGEMM_KWF=0 / HCTR_IGEMM_PAIR=0 at process start. */
int hctr_testing_set_conv_variant(int kw_fused_slab, int cta_pairs);

#ifdef __cplusplus
}
#endif
#endif /* HCTR_B200_TESTING_H_ */
