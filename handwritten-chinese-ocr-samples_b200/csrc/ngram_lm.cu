// hctr_ngram_score: kenlm.Model.score(sentence, bos=True, eos=False) for sequences of class indices, on the device.
#include "ngram_lm.cuh"

namespace hctr {

__global__ void ngram_score_kernel(hctr_ngram_lm lm, const int32_t* __restrict__ ids, const int32_t* __restrict__ offsets,
                                   int nseq, float* __restrict__ out) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nseq) return;
    int ctx[kNgramMaxOrder - 1];
    int m = 0;
    const int cap = lm.order - 1;
#pragma unroll
    for (int i = 0; i < kNgramMaxOrder - 1; ++i) ctx[i] = 0;
    if (cap > 0) { ctx[0] = lm.bos_id; m = 1; }
    float total = 0.f;
    for (int i = offsets[q]; i < offsets[q + 1]; ++i) {
        const int c = ids[i];
        const int w = (c >= 0 && c < lm.num_ids) ? __ldg(lm.vocab + c) : lm.unk_id;
        total = __fadd_rn(total, ngram_word_score(lm, ctx, m, w));
        ngram_push(ctx, m, cap, w);
    }
    out[q] = total;
}

int check_ngram(const hctr_ngram_lm* lm, const char* who) {
    HCTR_CHECK(lm && lm->entries && lm->backoff && lm->vocab, HCTR_ERR_INVALID, "%s: null n-gram table", who);
    HCTR_CHECK(lm->order >= 1 && lm->order <= kNgramMaxOrder, HCTR_ERR_INVALID, "%s: n-gram order must be in [1,%d] (got %d)", who, kNgramMaxOrder, lm->order);
    HCTR_CHECK((lm->mask & (lm->mask + 1)) == 0 && lm->mask > 0, HCTR_ERR_INVALID, "%s: table capacity must be a power of two", who);
    HCTR_CHECK(lm->num_ids > 0 && lm->num_ids <= 65536 && lm->bos_id >= 0 && lm->bos_id < lm->num_ids && lm->unk_id >= 0 && lm->unk_id < lm->num_ids,
               HCTR_ERR_INVALID, "%s: word ids must fit 16 bits", who);
    HCTR_CHECK((reinterpret_cast<uintptr_t>(lm->entries) & 15) == 0, HCTR_ERR_INVALID, "%s: table must be 16-byte aligned", who);
    return HCTR_OK;
}

}  // namespace hctr

using namespace hctr;

extern "C" {

int hctr_ngram_score(const hctr_ngram_lm* lm, const int32_t* ids, const int32_t* offsets, int nseq, float* out, void* stream) {
    int rc = check_ngram(lm, "ngram_score");
    if (rc) return rc;
    HCTR_CHECK(nseq >= 0 && (nseq == 0 || (ids && offsets && out)), HCTR_ERR_INVALID, "ngram_score: null pointer");
    if (nseq == 0) return HCTR_OK;
    ngram_score_kernel<<<(nseq + 63) / 64, 64, 0, static_cast<cudaStream_t>(stream)>>>(*lm, ids, offsets, nseq, out);
    HCTR_CUDA(cudaGetLastError());
    return HCTR_OK;
}

}  // extern "C"
