/*
 * oracle/ctc_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the reference's CTC codec / loss algorithms
 * (AndrewCullacino/handwritten-chinese-ocr-samples). It exists only so that tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs can check and time against it; nothing under
 * handwritten-chinese-ocr-samples_b200/ may link, import or call it.
 *
 * Parity pin: the reference ships no tests or golden vectors, so this file is pinned
 * against outputs of the reference itself, generated in the build container by
 * tests/golden/make_golden.py (which imports /root/reference) and committed under tests/golden/.
 * tests/test_oracle_golden.py replays them.
 *
 * Each function cites the reference lines it follows.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define NEG_INF (-INFINITY)

/* ------------------------------------------------------------------------------------------------
 * numpy.argmax over the class axis: first maximum wins; a NaN compares as the maximum and the first
 * NaN wins (numpy semantics relied on by utils/ctc_codec.py:75).
 * ---------------------------------------------------------------------------------------------- */
static int argmax_row(const float* row, int C) {
    int best = 0;
    float bv = row[0];
    if (bv != bv) return 0;
    for (int c = 1; c < C; ++c) {
        float v = row[c];
        if (v != v) return c;
        if (v > bv) { bv = v; best = c; }
    }
    return best;
}

/* ctc_codec.__greedy_search__ (utils/ctc_codec.py:70-99).
 * preds: [T][B][C] contiguous fp32. raw_argmax (may be NULL): [B][T]. out_idx: [B][T], out_len: [B]. */
void oracle_ctc_greedy(const float* preds, int T, int B, int C, int32_t* raw_argmax, int32_t* out_idx,
                       int32_t* out_len) {
    const int unknown = C - 1;                                       /* len(self.characters) - 1   :92 */
    for (int b = 0; b < B; ++b) {
        int n = 0, prev = -1;
        for (int t = 0; t < T; ++t) {
            const int cur = argmax_row(preds + ((size_t)t * B + b) * C, C);   /* np.argmax(preds, 2) :75 */
            if (raw_argmax) raw_argmax[(size_t)b * T + t] = cur;
            /* keep iff not blank, not unknown, and not equal to the RAW previous index          :91-93 */
            if (cur != 0 && cur != unknown && !(t > 0 && prev == cur)) out_idx[(size_t)b * T + n++] = cur;
            prev = cur;
        }
        out_len[b] = n;
    }
}

/* scipy.special.log_softmax along the class axis (utils/ctc_codec.py:65):
 *   tmp = x - max(x); out = tmp - log(sum(exp(tmp)))   — evaluated per row in the input dtype (fp32);
 * the sum is accumulated in double here (scipy/numpy use pairwise fp32 summation; the difference is
 * below 1e-6 and covered by the test tolerance). */
void oracle_log_softmax(const float* x, int rows, int C, float* out) {
    for (int r = 0; r < rows; ++r) {
        const float* p = x + (size_t)r * C;
        float* o = out + (size_t)r * C;
        float m = p[0];
        for (int c = 1; c < C; ++c) if (p[c] > m) m = p[c];
        double s = 0.0;
        for (int c = 0; c < C; ++c) s += (double)expf(p[c] - m);
        const float ls = logf((float)s);
        for (int c = 0; c < C; ++c) o[c] = (p[c] - m) - ls;
    }
}

/* Top-k of one row, descending, ties toward the lower index (np.flip(np.argsort(..)) leaves tie order
 * unspecified, utils/ctc_codec.py:186; fixtures avoid ties inside the top k+1). */
void oracle_topk_row(const float* row, int C, int k, int32_t* idx) {
    for (int j = 0; j < k; ++j) {
        int best = -1;
        for (int c = 0; c < C; ++c) {
            int taken = 0;
            for (int q = 0; q < j; ++q) if (idx[q] == c) { taken = 1; break; }
            if (taken) continue;
            if (best < 0 || row[c] > row[best]) best = c;
        }
        idx[j] = best;
    }
}

/* np.logaddexp for doubles (numpy/core/src/npymath: npy_logaddexp). */
static double logaddexp_d(double x, double y) {
    if (x == y) return x + 0.693147180559945309417232121458176568; /* log(2) */
    const double tmp = x - y;
    if (tmp > 0) return x + log1p(exp(-tmp));
    if (tmp <= 0) return y + log1p(exp(tmp));
    return tmp; /* NaN */
}

typedef struct {
    int32_t* prefix;   /* label indices */
    int len;
    double pb, pnb, pt; /* Beam fields, utils/ctc_codec.py:288-297 */
} Beam;

static double beam_prob(const Beam* b) { return logaddexp_d(b->pb, b->pnb); }           /* Beam.prob  :299-300 */
static double beam_total(const Beam* b) { return logaddexp_d(b->pb, b->pnb) + b->pt; }  /* Beam.total :302-303 */

static int find_beam(Beam* g, int n, const int32_t* pre, int len, int extra /* <0: none */) {
    const int L = len + (extra >= 0 ? 1 : 0);
    for (int i = 0; i < n; ++i) {
        if (g[i].len != L) continue;
        if (len > 0 && memcmp(g[i].prefix, pre, sizeof(int32_t) * len) != 0) continue;
        if (extra >= 0 && g[i].prefix[len] != extra) continue;
        return i;
    }
    return -1;
}

static int new_beam(Beam* g, int n, const int32_t* pre, int len, int extra, int cap) {
    Beam* b = &g[n];
    b->len = len + (extra >= 0 ? 1 : 0);
    b->prefix = (int32_t*)malloc(sizeof(int32_t) * (size_t)(cap + 1));
    if (len > 0) memcpy(b->prefix, pre, sizeof(int32_t) * len);
    if (extra >= 0) b->prefix[len] = extra;
    b->pb = NEG_INF; b->pnb = NEG_INF; b->pt = 0.0;          /* Beam(prefix, pb=NEG_INF, pnb=NEG_INF) :243-244 */
    return n;
}

/* ctc_codec.__cbs_full__ + __context_beam_search__ (utils/ctc_codec.py:183-285), use_tfm_pred=False,
 * use_tfm_score=False, with a unigram-table language model: ngram.score(' '.join(prefix+suffix)) is
 * restated as the left-to-right sum of lm_table[c] over prefix+suffix (lm_table NULL = zero LM).
 * logp: [T][B][C] fp32 log-softmax output; topk: [T][B][k] candidate indices.
 * Returns 0, or -4 for sequence(s) whose greedy path is empty (the reference raises IndexError, :198);
 * status[b] carries the per-sequence code. */
int oracle_ctc_beam_search(const float* logp, const int32_t* topk, int T, int B, int C, int k, int beam_size,
                           double lm_penalty, double len_bonus, const double* lm_table, int32_t* out_idx,
                           int32_t* out_len, int32_t* status) {
    const int unknown = C - 1;
    int rc = 0;
    const int max_gen = beam_size * (k + 1) + 1;
    int32_t* g_char = (int32_t*)malloc(sizeof(int32_t) * (size_t)(T + 1));
    int32_t* g_time = (int32_t*)malloc(sizeof(int32_t) * (size_t)(T + 1));
    Beam* kept = (Beam*)calloc((size_t)max_gen, sizeof(Beam));
    Beam* gen = (Beam*)calloc((size_t)max_gen, sizeof(Beam));
    int* order = (int*)malloc(sizeof(int) * (size_t)max_gen);

    for (int b = 0; b < B; ++b) {
        /* top_line: greedy (char, t) list from the top-1 candidates                               :188-195 */
        int ng = 0, prev = -1;
        for (int t = 0; t < T; ++t) {
            const int cur = topk[((size_t)t * B + b) * k];
            if (cur != 0 && cur != unknown && !(t > 0 && prev == cur)) { g_char[ng] = cur; g_time[ng] = t; ++ng; }
            prev = cur;
        }
        if (ng == 0) { status[b] = -4; out_len[b] = 0; rc = -4; continue; }   /* top_line[-1] -> IndexError :198 */
        status[b] = 0;
        int end_step = g_time[ng - 1] + 4;                                    /* :198-199 */
        if (end_step >= T) end_step = T;

        int nkept = 1;
        new_beam(kept, 0, NULL, 0, -1, T);
        kept[0].pb = 0.0; kept[0].pnb = NEG_INF; kept[0].pt = 0.0;            /* kept_beams = [Beam()] :200, :289 */
        int gptr = 0;
        for (int t = 0; t < end_step; ++t) {
            /* suffix: first <=4 greedy chars with time > t                                        :202-203 */
            while (gptr < ng && g_time[gptr] <= t) ++gptr;
            int nsuf = ng - gptr; if (nsuf > 4) nsuf = 4;
            const float* lp = logp + ((size_t)t * B + b) * C;
            const int32_t* cand = topk + ((size_t)t * B + b) * k;
            int ngen = 0;
            for (int j = 0; j < nkept; ++j) {                                   /* :235 */
                const Beam* in = &kept[j];
                const double P = beam_prob(in);
                const int tail = in->len ? in->prefix[in->len - 1] : -1;       /* :252 */
                for (int q = 0; q < k; ++q) {                                   /* :236 */
                    const int idx = cand[q];
                    if (idx >= unknown) continue;                               /* :238-239 */
                    const double p = (double)lp[idx];                           /* fp32 promoted on add */
                    int self = find_beam(gen, ngen, in->prefix, in->len, -1);
                    if (self < 0) { self = new_beam(gen, ngen, in->prefix, in->len, -1, T); ++ngen; }   /* :243-244 */
                    if (idx == 0) {                                             /* :246-249 */
                        gen[self].pb = logaddexp_d(gen[self].pb, P + p);
                        continue;
                    }
                    int ext = find_beam(gen, ngen, in->prefix, in->len, idx);
                    if (ext < 0) { ext = new_beam(gen, ngen, in->prefix, in->len, idx, T); ++ngen; }    /* :253-255 */
                    if (idx != tail) {
                        gen[ext].pnb = logaddexp_d(gen[ext].pnb, P + p);        /* :256-258 */
                    } else {
                        gen[ext].pnb = logaddexp_d(gen[ext].pnb, in->pb + p);   /* :261-262 */
                        gen[self].pnb = logaddexp_d(gen[self].pnb, in->pnb + p);/* :264-265 */
                    }
                }
            }
            /* LM score + length bonus                                                              :277-281 */
            for (int i = 0; i < ngen; ++i) {
                double lm = 0.0;
                if (lm_table) {
                    for (int c = 0; c < gen[i].len; ++c) lm += lm_table[gen[i].prefix[c]];
                    for (int c = 0; c < nsuf; ++c) lm += lm_table[g_char[gptr + c]];
                }
                gen[i].pt = lm * lm_penalty + (double)gen[i].len * len_bonus;
            }
            /* stable sort by total(), descending, keep beam_size                                   :283-285 */
            for (int i = 0; i < ngen; ++i) order[i] = i;
            for (int i = 1; i < ngen; ++i) {
                const int cur = order[i];
                const double tv = beam_total(&gen[cur]);
                int j2 = i - 1;
                while (j2 >= 0 && beam_total(&gen[order[j2]]) < tv) { order[j2 + 1] = order[j2]; --j2; }
                order[j2 + 1] = cur;
            }
            for (int j = 0; j < nkept; ++j) { free(kept[j].prefix); kept[j].prefix = NULL; }
            const int keep = ngen < beam_size ? ngen : beam_size;
            for (int j = 0; j < keep; ++j) { kept[j] = gen[order[j]]; gen[order[j]].prefix = NULL; }
            for (int i = 0; i < ngen; ++i) if (gen[i].prefix) { free(gen[i].prefix); gen[i].prefix = NULL; }
            nkept = keep;
        }
        out_len[b] = kept[0].len;                                              /* texts.append(kept_beams[0].prefix) :208 */
        for (int c = 0; c < kept[0].len; ++c) out_idx[(size_t)b * T + c] = kept[0].prefix[c];
        for (int j = 0; j < nkept; ++j) { free(kept[j].prefix); kept[j].prefix = NULL; }
    }
    free(g_char); free(g_time); free(kept); free(gen); free(order);
    return rc;
}

/* One __context_beam_search__ step (utils/ctc_codec.py:212-285) over an explicit candidate list (shared by the full and
 * the skip search). kept/nkept are replaced by the surviving beams. */
static void context_step(Beam* kept, int* nkept_io, Beam* gen, int* order, const float* lp, const int32_t* cand, int ncand,
                         int unknown, int beam_size, double lm_penalty, double len_bonus, const double* lm_table,
                         const int32_t* suffix, int nsuf, int cap) {
    int nkept = *nkept_io, ngen = 0;
    for (int j = 0; j < nkept; ++j) {
        const Beam* in = &kept[j];
        const double P = beam_prob(in);
        const int tail = in->len ? in->prefix[in->len - 1] : -1;
        for (int q = 0; q < ncand; ++q) {
            const int idx = cand[q];
            if (idx >= unknown) continue;
            const double p = (double)lp[idx];
            int self = find_beam(gen, ngen, in->prefix, in->len, -1);
            if (self < 0) { self = new_beam(gen, ngen, in->prefix, in->len, -1, cap); ++ngen; }
            if (idx == 0) { gen[self].pb = logaddexp_d(gen[self].pb, P + p); continue; }
            int ext = find_beam(gen, ngen, in->prefix, in->len, idx);
            if (ext < 0) { ext = new_beam(gen, ngen, in->prefix, in->len, idx, cap); ++ngen; }
            if (idx != tail) {
                gen[ext].pnb = logaddexp_d(gen[ext].pnb, P + p);
            } else {
                gen[ext].pnb = logaddexp_d(gen[ext].pnb, in->pb + p);
                gen[self].pnb = logaddexp_d(gen[self].pnb, in->pnb + p);
            }
        }
    }
    for (int i = 0; i < ngen; ++i) {
        double lm = 0.0;
        if (lm_table) {
            for (int c = 0; c < gen[i].len; ++c) lm += lm_table[gen[i].prefix[c]];
            for (int c = 0; c < nsuf; ++c) lm += lm_table[suffix[c]];
        }
        gen[i].pt = lm * lm_penalty + (double)gen[i].len * len_bonus;
    }
    for (int i = 0; i < ngen; ++i) order[i] = i;
    for (int i = 1; i < ngen; ++i) {
        const int cur = order[i];
        const double tv = beam_total(&gen[cur]);
        int j2 = i - 1;
        while (j2 >= 0 && beam_total(&gen[order[j2]]) < tv) { order[j2 + 1] = order[j2]; --j2; }
        order[j2 + 1] = cur;
    }
    for (int j = 0; j < nkept; ++j) { free(kept[j].prefix); kept[j].prefix = NULL; }
    const int keep = ngen < beam_size ? ngen : beam_size;
    for (int j = 0; j < keep; ++j) { kept[j] = gen[order[j]]; gen[order[j]].prefix = NULL; }
    for (int i = 0; i < ngen; ++i) if (gen[i].prefix) { free(gen[i].prefix); gen[i].prefix = NULL; }
    *nkept_io = keep;
}

/* ctc_codec.__cbs_skip__ (utils/ctc_codec.py:124-181): candidates at step t are ALL classes with log-prob > log(0.001)
 * in index order (:144); exactly one candidate -> the in-place fast path (:147-171, with its quirks: a blank step leaves
 * pnb untouched, no merging of equal prefixes); otherwise the normal context beam search over those candidates. A step
 * without any candidate empties the beam list and the reference ends in IndexError (:179) -> status -4.
 * top1: [T][B] arg-max class per step (the reference takes it from the arg-sort, :128-134). */
int oracle_ctc_beam_search_skip(const float* logp, const int32_t* top1, int T, int B, int C, int beam_size,
                                double lm_penalty, double len_bonus, const double* lm_table, int32_t* out_idx,
                                int32_t* out_len, int32_t* status) {
    const int unknown = C - 1;
    const double thresh = log(0.001);                         /* prune_thresh = np.log(0.001)  :129 */
    int rc = 0;
    const int max_gen = beam_size * (C + 1) + 1 < 200000 ? beam_size * (C + 1) + 1 : 200000;
    int32_t* g_char = (int32_t*)malloc(sizeof(int32_t) * (size_t)(T + 1));
    int32_t* g_time = (int32_t*)malloc(sizeof(int32_t) * (size_t)(T + 1));
    int32_t* cand = (int32_t*)malloc(sizeof(int32_t) * (size_t)C);
    Beam* kept = (Beam*)calloc((size_t)max_gen, sizeof(Beam));
    Beam* gen = (Beam*)calloc((size_t)max_gen, sizeof(Beam));
    int* order = (int*)malloc(sizeof(int) * (size_t)max_gen);
    for (int b = 0; b < B; ++b) {
        int ng = 0, prev = -1;
        for (int t = 0; t < T; ++t) {
            const int cur = top1[(size_t)t * B + b];
            if (cur != 0 && cur != unknown && !(t > 0 && prev == cur)) { g_char[ng] = cur; g_time[ng] = t; ++ng; }
            prev = cur;
        }
        if (ng == 0) { status[b] = -4; out_len[b] = 0; rc = -4; continue; }
        status[b] = 0;
        int end_step = g_time[ng - 1] + 4;
        if (end_step >= T) end_step = T;
        int nkept = 1;
        new_beam(kept, 0, NULL, 0, -1, T);
        kept[0].pb = 0.0; kept[0].pnb = NEG_INF; kept[0].pt = 0.0;
        int gptr = 0;
        for (int t = 0; t < end_step; ++t) {
            const float* lp = logp + ((size_t)t * B + b) * C;
            int nc = 0;
            for (int c = 0; c < C; ++c) if ((double)lp[c] > thresh) cand[nc++] = c;          /* np.where(...)  :144 */
            if (nc == 1) {                                                                    /* :147 */
                const int pidx = cand[0];
                if (pidx >= unknown) continue;                                                /* :150-151 */
                for (int j = 0; j < nkept; ++j) {
                    Beam* kb = &kept[j];
                    const int tail = kb->len ? kb->prefix[kb->len - 1] : -1;
                    if (pidx == 0) {
                        kb->pb = beam_prob(kb) + (double)lp[0];                               /* :155-157 */
                    } else if (pidx != tail) {
                        const double pr = beam_prob(kb);
                        kb->prefix[kb->len++] = pidx;
                        kb->pnb = pr + (double)lp[pidx];
                        kb->pb = NEG_INF;                                                     /* :158-161 */
                    } else if (kb->pb != NEG_INF) {
                        kb->prefix[kb->len++] = pidx;
                        kb->pnb = kb->pb + (double)lp[pidx];
                        kb->pb = NEG_INF;                                                     /* :163-167 */
                    } else {
                        kb->pb = beam_prob(kb) + (double)lp[0];
                        kb->pnb = kb->pnb + (double)lp[pidx];                                 /* :168-171 */
                    }
                }
            } else {
                while (gptr < ng && g_time[gptr] <= t) ++gptr;
                int nsuf = ng - gptr; if (nsuf > 4) nsuf = 4;
                context_step(kept, &nkept, gen, order, lp, cand, nc, unknown, beam_size, lm_penalty, len_bonus, lm_table,
                             g_char + gptr, nsuf, T);
            }
        }
        if (nkept == 0) { status[b] = -4; out_len[b] = 0; rc = -4; continue; }              /* kept_beams[0] -> IndexError :179 */
        out_len[b] = kept[0].len;
        for (int c = 0; c < kept[0].len; ++c) out_idx[(size_t)b * T + c] = kept[0].prefix[c];
        for (int j = 0; j < nkept; ++j) { free(kept[j].prefix); kept[j].prefix = NULL; }
    }
    free(g_char); free(g_time); free(cand); free(kept); free(gen); free(order);
    return rc;
}

/* ------------------------------------------------------------------------------------------------
 * CTC loss forward/backward as called by the reference training loop (main.py:205,406-409):
 *   criterion = CTCLoss(blank=0, reduction='mean', zero_infinity=True)
 *   loss = criterion(preds.log_softmax(2), targets, input_lengths, target_lengths)
 * Restates the alpha/beta recursion of torch's native ctc_loss (ATen LossCTC.cpp, not vendored in the
 * reference; torch pinned only as >=1.8.0 in requirements.txt:1) in double precision, and the gradient with
 * respect to the *logits* (SURVEY.md §8 a13):
 *   g[t,b,c] = (softmax[t,b,c] - sum_{s: l'_s = c} exp(alpha_t(s)+beta_t(s) - ll_b - lp[t,b,c])) / (max(L_b,1) * B)
 * logits: [T][B][C] fp32 contiguous. nll: [B]. grad (may be NULL): [T][B][C]. Returns the mean loss.
 * ---------------------------------------------------------------------------------------------- */
double oracle_ctc_loss(const float* logits, int T, int B, int C, const int32_t* targets,
                       const int32_t* target_lengths, const int32_t* input_lengths, double* nll, double* grad) {
    double loss = 0.0;
    size_t toff = 0;
    if (grad) memset(grad, 0, sizeof(double) * (size_t)T * B * C);
    for (int b = 0; b < B; ++b) {
        const int L = target_lengths[b];
        const int Tb = input_lengths[b];
        const int S = 2 * L + 1;
        const int32_t* tg = targets + toff;
        toff += (size_t)L;
        double* lse = (double*)malloc(sizeof(double) * (size_t)(Tb > 0 ? Tb : 1));
        double* alpha = (double*)malloc(sizeof(double) * (size_t)(Tb > 0 ? Tb : 1) * S);
        double* beta = (double*)malloc(sizeof(double) * (size_t)(Tb > 0 ? Tb : 1) * S);
        for (int t = 0; t < Tb; ++t) {
            const float* row = logits + ((size_t)t * B + b) * C;
            double m = row[0];
            for (int c = 1; c < C; ++c) if (row[c] > m) m = row[c];
            double s = 0.0;
            for (int c = 0; c < C; ++c) s += exp((double)row[c] - m);
            lse[t] = m + log(s);
        }
#define LBL(s) (((s) & 1) ? tg[(s) >> 1] : 0)
#define LP(t, s) ((double)logits[((size_t)(t) * B + b) * C + LBL(s)] - lse[t])
        double ll = NEG_INF;
        if (Tb > 0) {
            for (int s = 0; s < S; ++s) alpha[s] = NEG_INF;
            alpha[0] = LP(0, 0);
            if (S > 1) alpha[1] = LP(0, 1);
            for (int t = 1; t < Tb; ++t) {
                for (int s = 0; s < S; ++s) {
                    double a = alpha[(size_t)(t - 1) * S + s];
                    if (s > 0) a = logaddexp_d(a, alpha[(size_t)(t - 1) * S + s - 1]);
                    if (s > 1 && LBL(s) != 0 && LBL(s) != LBL(s - 2)) a = logaddexp_d(a, alpha[(size_t)(t - 1) * S + s - 2]);
                    alpha[(size_t)t * S + s] = (a == NEG_INF) ? NEG_INF : a + LP(t, s);
                }
            }
            ll = alpha[(size_t)(Tb - 1) * S + S - 1];
            if (S > 1) ll = logaddexp_d(ll, alpha[(size_t)(Tb - 1) * S + S - 2]);
        } else if (L == 0) {
            ll = 0.0;
        }
        double n = -ll;
        const int infeasible = !(n < INFINITY);          /* zero_infinity=True: inf -> 0, zero grad */
        if (infeasible) n = 0.0;
        nll[b] = n;
        loss += n / (double)(L > 0 ? L : 1);
        if (grad && !infeasible && Tb > 0) {
            for (int s = 0; s < S; ++s) beta[(size_t)(Tb - 1) * S + s] = NEG_INF;
            beta[(size_t)(Tb - 1) * S + S - 1] = LP(Tb - 1, S - 1);
            if (S > 1) beta[(size_t)(Tb - 1) * S + S - 2] = LP(Tb - 1, S - 2);
            for (int t = Tb - 2; t >= 0; --t) {
                for (int s = 0; s < S; ++s) {
                    double a = beta[(size_t)(t + 1) * S + s];
                    if (s + 1 < S) a = logaddexp_d(a, beta[(size_t)(t + 1) * S + s + 1]);
                    if (s + 2 < S && LBL(s + 2) != 0 && LBL(s + 2) != LBL(s)) a = logaddexp_d(a, beta[(size_t)(t + 1) * S + s + 2]);
                    beta[(size_t)t * S + s] = (a == NEG_INF) ? NEG_INF : a + LP(t, s);
                }
            }
            const double scale = 1.0 / ((double)(L > 0 ? L : 1) * (double)B);
            for (int t = 0; t < Tb; ++t) {
                const float* row = logits + ((size_t)t * B + b) * C;
                double* g = grad + ((size_t)t * B + b) * C;
                for (int c = 0; c < C; ++c) g[c] = exp((double)row[c] - lse[t]);
                for (int s = 0; s < S; ++s) {
                    const double ab = alpha[(size_t)t * S + s] + beta[(size_t)t * S + s];
                    if (ab == NEG_INF) continue;
                    g[LBL(s)] -= exp(ab - ll - LP(t, s));
                }
                for (int c = 0; c < C; ++c) g[c] *= scale;
            }
            /* frames t >= input_length keep zero gradient (torch: grad is zeroed beyond input_length) */
        }
#undef LBL
#undef LP
        free(lse); free(alpha); free(beta);
    }
    return loss / (double)B;
}

/* editdistance.eval(a, b): plain Levenshtein distance (insert / delete / substitute, unit costs), the metric of the
 * reference's CER (main.py:506-517, test.py:275-286; the `editdistance` PyPI package, unpinned in requirements.txt:6). */
int oracle_edit_distance(const int32_t* a, int n, const int32_t* b, int m) {
    int* prev = (int*)malloc(sizeof(int) * (size_t)(m + 1));
    int* cur = (int*)malloc(sizeof(int) * (size_t)(m + 1));
    for (int j = 0; j <= m; ++j) prev[j] = j;
    for (int i = 1; i <= n; ++i) {
        cur[0] = i;
        for (int j = 1; j <= m; ++j) {
            int v = prev[j - 1] + (a[i - 1] != b[j - 1]);
            if (prev[j] + 1 < v) v = prev[j] + 1;
            if (cur[j - 1] + 1 < v) v = cur[j - 1] + 1;
            cur[j] = v;
        }
        int* t = prev; prev = cur; cur = t;
    }
    const int d = prev[m];
    free(prev); free(cur);
    return d;
}
