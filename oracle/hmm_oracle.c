/*
 * hmm_oracle.c -- CPU restatement of the pytorch_hmm hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the checker for the B200 kernels.  Nothing under pytorch_hmm_b200/ may call it;
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
 *
 * Every function restates one reference routine (citations are relative to /root/reference):
 *   orc_viterbi_f32            pytorch_hmm/hmm.py:152-178          (delta/psi recursion + traceback)
 *   orc_forward_backward_f64   pytorch_hmm/hmm.py:86-126           (log-space alpha/beta/gamma, in double)
 *   orc_forward_backward_f32   same recursion, fp32 arithmetic     (mirrors the reference's own rounding)
 *   orc_gmm_emission_f64       pytorch_hmm/mixture_gaussian.py:141-214, hmm_layer.py:300-321, hsmm.py:194-204
 *   orc_hsmm_viterbi_f32       pytorch_hmm/hsmm.py:245-354         (HSMMLayer._viterbi_decode_single)
 *   orc_hsmm_forward_f64       pytorch_hmm/semi_markov.py:308-383  (SemiMarkovHMM._unsupervised_forward)
 *   orc_hsmm_backward_f64      no reference (new); matching beta recursion, checked by brute force in tests
 *   orc_bw_stats_f64           docs/01_hmm_theory.md:196-227       (Baum-Welch sufficient statistics)
 *   orc_greedy_decode_f32      pytorch_hmm/streaming.py:292-308    (per-frame greedy argmax chain)
 *   orc_tv_viterbi_f32         pytorch_hmm/neural.py:463-511       (NeuralHMM Viterbi, time-varying transitions)
 *   orc_tv_forward_backward_f64 pytorch_hmm/neural.py:403-461      (NeuralHMM forward / backward, in double)
 *
 * Parity status: pinned.  tests/test_oracle_golden.py checks these against fixtures produced by
 * importing the real reference (oracle/make_golden.py -> the .npz fixtures under tests/golden).
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -shared -fPIC  (see oracle/Makefile).
 * -ffp-contract=off matters: the fp32 routines must round after every add exactly like ATen does.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

static inline double lse2(double a, double b) {
    if (a == -INFINITY) return b;
    if (b == -INFINITY) return a;
    double m = a > b ? a : b;
    return m + log(exp(a - m) + exp(b - m));
}

/* ------------------------------------------------------------------------------------------
 * Viterbi, fp32, first-index ties.  hmm.py:159-178:
 *   delta_0 = log_p0 + log_b_0
 *   (m, psi_t[j]) = max_i (delta_{t-1}[i] + logP[i][j])   -- torch.max returns the lowest index on ties
 *   delta_t[j] = m + log_b_t[j]                            -- second rounding
 *   s_{T-1} = argmax_j delta_{T-1}[j] (first index); s_t = psi_{t+1}[s_{t+1}]
 * psi_0 is all zero, as in the reference (hmm.py:156).
 * ------------------------------------------------------------------------------------------ */
ORC_API void orc_viterbi_f32(const float *logb, const float *logP, const float *logp0,
                             int B, int T, int K,
                             float *delta, int32_t *psi, int64_t *states, float *score) {
    for (int b = 0; b < B; ++b) {
        const float *lb = logb + (size_t)b * T * K;
        float *dl = delta + (size_t)b * T * K;
        int32_t *ps = psi + (size_t)b * T * K;
        for (int j = 0; j < K; ++j) {
            dl[j] = logp0[j] + lb[j];
            ps[j] = 0;
        }
        for (int t = 1; t < T; ++t) {
            const float *prev = dl + (size_t)(t - 1) * K;
            for (int j = 0; j < K; ++j) {
                float best = prev[0] + logP[j];
                int arg = 0;
                for (int i = 1; i < K; ++i) {
                    float c = prev[i] + logP[(size_t)i * K + j];
                    if (c > best) { best = c; arg = i; }
                }
                dl[(size_t)t * K + j] = best + lb[(size_t)t * K + j];
                ps[(size_t)t * K + j] = arg;
            }
        }
        const float *last = dl + (size_t)(T - 1) * K;
        int s = 0;
        for (int j = 1; j < K; ++j) if (last[j] > last[s]) s = j;
        if (score) score[b] = last[s];
        int64_t *st = states + (size_t)b * T;
        st[T - 1] = s;
        for (int t = T - 2; t >= 0; --t) {
            s = ps[(size_t)(t + 1) * K + s];
            st[t] = s;
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * Forward-backward in log space, double precision ("truth" for the 1e-4 gates; SURVEY finding 9).
 * hmm.py:92-126.  logb is the *effective* log-emission (after any floor).  Outputs may be NULL.
 * ------------------------------------------------------------------------------------------ */
ORC_API void orc_forward_backward_f64(const double *logb, const double *logP, const double *logp0,
                                      int B, int T, int K,
                                      double *log_alpha, double *log_beta, double *gamma, double *loglik) {
    double *la = (double *)malloc(sizeof(double) * (size_t)T * K);
    double *lbeta = (double *)malloc(sizeof(double) * (size_t)T * K);
    double *tmp = (double *)malloc(sizeof(double) * (size_t)K);
    for (int b = 0; b < B; ++b) {
        const double *lb = logb + (size_t)b * T * K;
        for (int j = 0; j < K; ++j) la[j] = logp0[j] + lb[j];
        for (int t = 1; t < T; ++t) {
            for (int j = 0; j < K; ++j) {
                double m = -INFINITY;
                for (int i = 0; i < K; ++i) {
                    tmp[i] = la[(size_t)(t - 1) * K + i] + logP[(size_t)i * K + j];
                    if (tmp[i] > m) m = tmp[i];
                }
                double s = 0.0;
                if (m == -INFINITY) { la[(size_t)t * K + j] = -INFINITY; continue; }
                for (int i = 0; i < K; ++i) s += exp(tmp[i] - m);
                la[(size_t)t * K + j] = m + log(s) + lb[(size_t)t * K + j];
            }
        }
        for (int i = 0; i < K; ++i) lbeta[(size_t)(T - 1) * K + i] = 0.0;
        for (int t = T - 2; t >= 0; --t) {
            for (int i = 0; i < K; ++i) {
                double m = -INFINITY;
                for (int j = 0; j < K; ++j) {
                    tmp[j] = logP[(size_t)i * K + j] + lb[(size_t)(t + 1) * K + j] + lbeta[(size_t)(t + 1) * K + j];
                    if (tmp[j] > m) m = tmp[j];
                }
                if (m == -INFINITY) { lbeta[(size_t)t * K + i] = -INFINITY; continue; }
                double s = 0.0;
                for (int j = 0; j < K; ++j) s += exp(tmp[j] - m);
                lbeta[(size_t)t * K + i] = m + log(s);
            }
        }
        if (loglik) {
            double acc = -INFINITY;
            for (int j = 0; j < K; ++j) acc = lse2(acc, la[(size_t)(T - 1) * K + j]);
            loglik[b] = acc;
        }
        if (gamma) {
            double *g = gamma + (size_t)b * T * K;
            for (int t = 0; t < T; ++t) {
                double m = -INFINITY;
                for (int j = 0; j < K; ++j) {
                    tmp[j] = la[(size_t)t * K + j] + lbeta[(size_t)t * K + j];
                    if (tmp[j] > m) m = tmp[j];
                }
                double s = 0.0;
                for (int j = 0; j < K; ++j) s += exp(tmp[j] - m);
                double z = m + log(s);
                for (int j = 0; j < K; ++j) g[(size_t)t * K + j] = exp(tmp[j] - z);
            }
        }
        if (log_alpha) memcpy(log_alpha + (size_t)b * T * K, la, sizeof(double) * (size_t)T * K);
        if (log_beta) memcpy(log_beta + (size_t)b * T * K, lbeta, sizeof(double) * (size_t)T * K);
    }
    free(la); free(lbeta); free(tmp);
}

/* Same recursion carried in fp32 (max-subtracted logsumexp like ATen's): reproduces the magnitude of the
 * reference's own rounding noise at long T.  Used only to report "reference-like fp32 error vs double". */
ORC_API void orc_forward_backward_f32(const float *logb, const float *logP, const float *logp0,
                                      int B, int T, int K,
                                      float *log_alpha, float *log_beta, float *gamma) {
    float *tmp = (float *)malloc(sizeof(float) * (size_t)K);
    for (int b = 0; b < B; ++b) {
        const float *lb = logb + (size_t)b * T * K;
        float *la = log_alpha + (size_t)b * T * K;
        float *lbe = log_beta + (size_t)b * T * K;
        for (int j = 0; j < K; ++j) la[j] = logp0[j] + lb[j];
        for (int t = 1; t < T; ++t)
            for (int j = 0; j < K; ++j) {
                float m = -INFINITY;
                for (int i = 0; i < K; ++i) {
                    tmp[i] = la[(size_t)(t - 1) * K + i] + logP[(size_t)i * K + j];
                    if (tmp[i] > m) m = tmp[i];
                }
                float s = 0.f;
                for (int i = 0; i < K; ++i) s += expf(tmp[i] - m);
                la[(size_t)t * K + j] = (logf(s) + m) + lb[(size_t)t * K + j];
            }
        for (int i = 0; i < K; ++i) lbe[(size_t)(T - 1) * K + i] = 0.f;
        for (int t = T - 2; t >= 0; --t)
            for (int i = 0; i < K; ++i) {
                float m = -INFINITY;
                for (int j = 0; j < K; ++j) {
                    tmp[j] = (logP[(size_t)i * K + j] + lb[(size_t)(t + 1) * K + j]) + lbe[(size_t)(t + 1) * K + j];
                    if (tmp[j] > m) m = tmp[j];
                }
                float s = 0.f;
                for (int j = 0; j < K; ++j) s += expf(tmp[j] - m);
                lbe[(size_t)t * K + i] = logf(s) + m;
            }
        if (gamma) {
            float *g = gamma + (size_t)b * T * K;
            for (int t = 0; t < T; ++t) {
                float m = -INFINITY;
                for (int j = 0; j < K; ++j) {
                    tmp[j] = la[(size_t)t * K + j] + lbe[(size_t)t * K + j];
                    if (tmp[j] > m) m = tmp[j];
                }
                float s = 0.f;
                for (int j = 0; j < K; ++j) s += expf(tmp[j] - m);
                float z = logf(s) + m;
                for (int j = 0; j < K; ++j) g[(size_t)t * K + j] = expf(tmp[j] - z);
            }
        }
    }
    free(tmp);
}

/* ------------------------------------------------------------------------------------------
 * Diagonal-Gaussian / GMM emission log-likelihood in double.
 *   comp[k][c] = -0.5 * ( sum_d (x_d - mu)^2 / var + sum_d log var + D log 2pi )
 *     mixture_gaussian.py:200-214 (var = exp(log_vars));  hmm_layer.py:300-309,321 (log var = 2*log_scales);
 *     hsmm.py:194-204.  `log_var_scale` selects the parameterisation (1 or 2).
 *   out[k] = own_lse_c( comp[k][c] + logw[k][c] )  with the reference's private logsumexp
 *     (mixture_gaussian.py:141-155: max, inf->0, log(clamp(sum,1e-8)) + max).  C == 1 and logw == NULL
 *     means "single Gaussian": out[k] = comp[k][0] with no LSE (hmm_layer.py:321).
 * ------------------------------------------------------------------------------------------ */
ORC_API void orc_gmm_emission_f64(const float *x, const float *means, const float *log_vars,
                                  double log_var_scale, const float *logw,
                                  int64_t N, int K, int C, int D, double *out) {
    const double log2pi = log(2.0 * M_PI);
    double *comp = (double *)malloc(sizeof(double) * (size_t)C);
    for (int64_t n = 0; n < N; ++n) {
        const float *xn = x + (size_t)n * D;
        for (int k = 0; k < K; ++k) {
            for (int c = 0; c < C; ++c) {
                const float *mu = means + ((size_t)k * C + c) * D;
                const float *lv = log_vars + ((size_t)k * C + c) * D;
                double q = 0.0, sl = 0.0;
                for (int d = 0; d < D; ++d) {
                    double lvd = log_var_scale * (double)lv[d];
                    double diff = (double)xn[d] - (double)mu[d];
                    q += diff * diff / exp(lvd);
                    sl += lvd;
                }
                comp[c] = -0.5 * (q + sl + D * log2pi);
                if (logw) comp[c] += (double)logw[(size_t)k * C + c];
            }
            if (C == 1 && !logw) { out[(size_t)n * K + k] = comp[0]; continue; }
            double m = comp[0];
            for (int c = 1; c < C; ++c) if (comp[c] > m) m = comp[c];
            if (isinf(m)) m = 0.0;
            double s = 0.0;
            for (int c = 0; c < C; ++c) s += exp(comp[c] - m);
            if (s < 1e-8) s = 1e-8;
            out[(size_t)n * K + k] = log(s) + m;
        }
    }
    free(comp);
}

/* ------------------------------------------------------------------------------------------
 * HSMMLayer Viterbi, fp32, exact reference operation order (hsmm.py:245-354).
 *   Dm = number of duration bins (the table is indexed d-1 for d = 1..Dm; hsmm.py:253,264).
 *   first segments:  delta[d-1][s][d-1] = seg(0,d,s) + logdur[s][d-1]                     (:262-268, no prior)
 *   later segments:  candidates (s'!=s, d') in lexicographic order, strict '>' keeps the first maximum of
 *                    ((delta[t-1][s'][d'-1] + logA[s'][s]) + seg(t,d,s)) + logdur[s][d-1]   (:271-316)
 *   seg(t,d,s) = torch.sum(obs[t:t+d, s]) on a strided fp32 view.  ATen's strided row_sum keeps four
 *   interleaved partial sums, folds the tail into partial 0 and then adds partials 1..3 in order;
 *   seg_sum4() reproduces that bit for bit (verified against torch 2.11 in oracle/make_golden.py).
 *   final: strict '>' over (s, d) lexicographic at t = T-1 (:319-329); segment backtrack (:332-352).
 * ------------------------------------------------------------------------------------------ */
static float seg_sum4(const float *col, int stride, int d) {
    float p[4] = {0.f, 0.f, 0.f, 0.f};
    int q = d / 4;
    for (int i = 0; i < q; ++i)
        for (int k = 0; k < 4; ++k) p[k] = p[k] + col[(size_t)(4 * i + k) * stride];
    for (int i = 4 * q; i < d; ++i) p[0] = p[0] + col[(size_t)i * stride];
    for (int k = 1; k < 4; ++k) p[0] = p[0] + p[k];
    return p[0];
}

ORC_API void orc_hsmm_viterbi_f32(const float *logb, const float *logdur, const float *logA,
                                  int B, int T, int K, int Dm,
                                  int64_t *states, float *score) {
    size_t tab = (size_t)T * K * Dm;
    float *delta = (float *)malloc(sizeof(float) * tab);
    int32_t *ps = (int32_t *)malloc(sizeof(int32_t) * tab);
    int32_t *pd = (int32_t *)malloc(sizeof(int32_t) * tab);
#define IDX(t, s, d) (((size_t)(t) * K + (s)) * Dm + (d))
    for (int b = 0; b < B; ++b) {
        const float *ob = logb + (size_t)b * T * K;
        for (size_t i = 0; i < tab; ++i) { delta[i] = -INFINITY; ps[i] = 0; pd[i] = 0; }
        for (int s = 0; s < K; ++s)
            for (int d = 1; d <= Dm && d <= T; ++d)
                delta[IDX(d - 1, s, d - 1)] = seg_sum4(ob + s, K, d) + logdur[(size_t)s * Dm + d - 1];
        for (int t = 1; t < T; ++t)
            for (int s = 0; s < K; ++s)
                for (int d = 1; d <= Dm && t + d - 1 < T; ++d) {
                    int te = t + d - 1;
                    float osum = seg_sum4(ob + (size_t)t * K + s, K, d);
                    float dsc = logdur[(size_t)s * Dm + d - 1];
                    float best = -INFINITY; int bs = 0, bd = 1;
                    for (int sp = 0; sp < K; ++sp) {
                        if (sp == s) continue;
                        for (int dp = 1; dp <= Dm; ++dp) {
                            if (t - 1 - dp + 1 < 0) continue;
                            float prev = delta[IDX(t - 1, sp, dp - 1)];
                            if (prev == -INFINITY) continue;
                            float tot = ((prev + logA[(size_t)sp * K + s]) + osum) + dsc;
                            if (tot > best) { best = tot; bs = sp; bd = dp; }
                        }
                    }
                    if (best != -INFINITY) {
                        delta[IDX(te, s, d - 1)] = best;
                        ps[IDX(te, s, d - 1)] = bs;
                        pd[IDX(te, s, d - 1)] = bd;
                    }
                }
        float best = -INFINITY; int cs = 0, cd = 1;
        for (int s = 0; s < K; ++s)
            for (int d = 1; d <= Dm; ++d) {
                float v = delta[IDX(T - 1, s, d - 1)];
                if (v > best) { best = v; cs = s; cd = d; }
            }
        if (score) score[b] = best;
        int64_t *st = states + (size_t)b * T;
        for (int t = 0; t < T; ++t) st[t] = 0;
        int t = T - 1;
        while (t >= 0) {
            int st0 = t - cd + 1; if (st0 < 0) st0 = 0;
            for (int u = st0; u <= t; ++u) st[u] = cs;
            if (st0 > 0) {
                int ns = ps[IDX(t, cs, cd - 1)], nd = pd[IDX(t, cs, cd - 1)];
                t = st0 - 1; cs = ns; cd = nd;
            } else break;
        }
    }
#undef IDX
    free(delta); free(ps); free(pd);
}

/* ------------------------------------------------------------------------------------------
 * HSMM forward (and a matching backward) in double.  semi_markov.py:308-383 with the crash at :353
 * repaired the obvious way (the accumulator is a value, not a Python float).
 *   seg[t][s][d]  : log-probability of frames (t-d+1 .. t) under state s as ONE segment; supplied by the
 *                   caller because SemiMarkovHMM counts the Gaussian constant once per segment
 *                   (semi_markov.py:422-424) while HSMMLayer counts it per frame (hsmm.py:285).
 *   alpha[t][s][d] = seg[t][s][d] + logdur[s][d] + ( logpi[s]                              if t-d+1 == 0
 *                                                   LSE_{s'!=s, d'} alpha[t-d][s'][d'] + logA[s'][s]  otherwise )
 *   total = LSE_{s,d} alpha[T-1][s][d]                                                  (:372-378)
 *   beta[t][s] (new) = log P(o_{t+1..T-1} | a segment of s ends at t); beta[T-1] = 0.
 * Layout: seg, alpha = [T][K][Dm]; index d-1 for duration d.  alpha/beta may be NULL.
 * ------------------------------------------------------------------------------------------ */
ORC_API void orc_hsmm_forward_f64(const double *seg, const double *logdur, const double *logA,
                                  const double *logpi, int T, int K, int Dm,
                                  double *alpha_out, double *beta_out, double *total) {
    size_t tab = (size_t)T * K * Dm;
    double *al = (double *)malloc(sizeof(double) * tab);
    double *endv = (double *)malloc(sizeof(double) * (size_t)T * K);   /* LSE_d alpha[t][s][d] */
#define IDX(t, s, d) (((size_t)(t) * K + (s)) * Dm + (d))
    for (size_t i = 0; i < tab; ++i) al[i] = -INFINITY;
    for (int t = 0; t < T; ++t) {
        for (int s = 0; s < K; ++s) {
            for (int d = 1; d <= Dm && d <= t + 1; ++d) {
                int st = t - d + 1;
                double inc;
                if (st == 0) inc = logpi[s];
                else {
                    inc = -INFINITY;
                    for (int sp = 0; sp < K; ++sp) if (sp != s)
                        inc = lse2(inc, endv[(size_t)(st - 1) * K + sp] + logA[(size_t)sp * K + s]);
                }
                if (inc == -INFINITY) continue;
                al[IDX(t, s, d - 1)] = inc + seg[IDX(t, s, d - 1)] + logdur[(size_t)s * Dm + d - 1];
            }
        }
        for (int s = 0; s < K; ++s) {
            double e = -INFINITY;
            for (int d = 0; d < Dm; ++d) e = lse2(e, al[IDX(t, s, d)]);
            endv[(size_t)t * K + s] = e;
        }
    }
    double tot = -INFINITY;
    for (int s = 0; s < K; ++s) tot = lse2(tot, endv[(size_t)(T - 1) * K + s]);
    if (total) *total = tot;
    if (alpha_out) memcpy(alpha_out, al, sizeof(double) * tab);
    if (beta_out) {
        /* beta[t][s]: state s's segment ended at t.  beta[T-1][s] = 0.
         * beta[t][s] = LSE_{s'!=s, d} logA[s][s'] + logdur[s'][d] + seg[t+d][s'][d] + beta[t+d][s'] */
        for (int s = 0; s < K; ++s) beta_out[(size_t)(T - 1) * K + s] = 0.0;
        for (int t = T - 2; t >= 0; --t)
            for (int s = 0; s < K; ++s) {
                double acc = -INFINITY;
                for (int sp = 0; sp < K; ++sp) if (sp != s)
                    for (int d = 1; d <= Dm && t + d < T; ++d)
                        acc = lse2(acc, logA[(size_t)s * K + sp] + logdur[(size_t)sp * Dm + d - 1] +
                                        seg[IDX(t + d, sp, d - 1)] + beta_out[(size_t)(t + d) * K + sp]);
                beta_out[(size_t)t * K + s] = acc;
            }
    }
#undef IDX
    free(al); free(endv);
}

/* ------------------------------------------------------------------------------------------
 * Baum-Welch sufficient statistics in double (docs/01_hmm_theory.md:196-227; the Gaussian-mixture
 * M-step statistics are the standard extension):
 *   gamma1[k]      += gamma_0[k]                               (:216)
 *   xi[i][j]       += sum_t xi_t(i,j)                          (:209,:221)
 *   occ[k][c]      += sum_t gamma_t[k] * resp_t[k][c]
 *   sx[k][c][d]    += sum_t gamma_t[k] * resp_t[k][c] * x_t[d]
 *   sxx[k][c][d]   += sum_t gamma_t[k] * resp_t[k][c] * x_t[d]^2
 *   loglik         += log p(o_1..T)
 * Inputs are the per-component log-likelihoods comp[n][k][c] (already including log w) so that the
 * caller controls the emission parameterisation.  logb_k = LSE_c comp.
 * ------------------------------------------------------------------------------------------ */
ORC_API void orc_bw_stats_f64(const float *x, const double *comp, const double *logP, const double *logp0,
                              int B, int T, int K, int C, int D,
                              double *gamma1, double *xi, double *occ, double *sx, double *sxx, double *loglik) {
    size_t n = (size_t)T * K;
    double *lb = (double *)malloc(sizeof(double) * n);
    double *la = (double *)malloc(sizeof(double) * n);
    double *lbe = (double *)malloc(sizeof(double) * n);
    double *g = (double *)malloc(sizeof(double) * n);
    for (int b = 0; b < B; ++b) {
        const double *cp = comp + (size_t)b * T * K * C;
        for (int t = 0; t < T; ++t)
            for (int k = 0; k < K; ++k) {
                double a = -INFINITY;
                for (int c = 0; c < C; ++c) a = lse2(a, cp[((size_t)t * K + k) * C + c]);
                lb[(size_t)t * K + k] = a;
            }
        double ll;
        orc_forward_backward_f64(lb, logP, logp0, 1, T, K, la, lbe, g, &ll);
        *loglik += ll;
        for (int k = 0; k < K; ++k) gamma1[k] += g[k];
        for (int t = 0; t + 1 < T; ++t)
            for (int i = 0; i < K; ++i)
                for (int j = 0; j < K; ++j)
                    xi[(size_t)i * K + j] += exp(la[(size_t)t * K + i] + logP[(size_t)i * K + j] +
                                                 lb[(size_t)(t + 1) * K + j] + lbe[(size_t)(t + 1) * K + j] - ll);
        for (int t = 0; t < T; ++t) {
            const float *xt = x + ((size_t)b * T + t) * D;
            for (int k = 0; k < K; ++k)
                for (int c = 0; c < C; ++c) {
                    double r = g[(size_t)t * K + k] * exp(cp[((size_t)t * K + k) * C + c] - lb[(size_t)t * K + k]);
                    occ[(size_t)k * C + c] += r;
                    double *px = sx + ((size_t)k * C + c) * D, *pxx = sxx + ((size_t)k * C + c) * D;
                    for (int d = 0; d < D; ++d) { px[d] += r * xt[d]; pxx[d] += r * (double)xt[d] * xt[d]; }
                }
        }
    }
    free(lb); free(la); free(lbe); free(g);
}

/* ------------------------------------------------------------------------------------------
 * Greedy streaming decode (streaming.py:292-308): s_0 = argmax(logb_0 - log K) on the first chunk, or
 * argmax(logA[prev] + logb_0) when continuing; s_t = argmax(logA[s_{t-1}] + logb_t).  First-index ties.
 * prev_state < 0 means "first chunk".  conf_t = exp(score_t)  (:320).
 * ------------------------------------------------------------------------------------------ */
ORC_API void orc_greedy_decode_f32(const float *logb, const float *logA, int T, int K, int prev_state,
                                   int64_t *states, float *scores) {
    int s = prev_state;
    for (int t = 0; t < T; ++t) {
        const float *lb = logb + (size_t)t * K;
        float best; int arg = 0;
        if (s < 0) {
            float lk = logf((float)K);
            best = lb[0] - lk;
            for (int j = 1; j < K; ++j) { float v = lb[j] - lk; if (v > best) { best = v; arg = j; } }
        } else {
            best = logA[(size_t)s * K] + lb[0];
            for (int j = 1; j < K; ++j) { float v = logA[(size_t)s * K + j] + lb[j]; if (v > best) { best = v; arg = j; } }
        }
        s = arg; states[t] = arg; if (scores) scores[t] = best;
    }
}


/* ------------------------------------------------------------------------------------------
 * NeuralHMM recursions: log-emissions used as they are, a [K,K] log-transition slice per frame.
 * neural.py:487-499: log_delta[0] = log_init + log_obs[0]; for t >= 1 the slice t-1 carries t-1 -> t:
 *   (m, psi_t[j]) = max_i(log_delta[t-1][i] + logT[t-1][i][j]) (first index), log_delta[t][j] = m + log_obs[t][j]
 * neural.py:502-506: s_{T-1} = argmax (first index); s_t = psi_{t+1}[s_{t+1}].
 * ------------------------------------------------------------------------------------------ */
ORC_API void orc_tv_viterbi_f32(const float *logb, const float *logT, const float *logp0, int B, int T, int K,
                                float *delta, int32_t *psi, int64_t *states) {
    for (int b = 0; b < B; ++b) {
        const float *lb = logb + (size_t)b * T * K;
        const float *lt = logT + (size_t)b * T * K * K;
        float *dl = delta + (size_t)b * T * K;
        int32_t *ps = psi + (size_t)b * T * K;
        for (int j = 0; j < K; ++j) { dl[j] = logp0[j] + lb[j]; ps[j] = 0; }
        for (int t = 1; t < T; ++t) {
            const float *sl = lt + (size_t)(t - 1) * K * K;
            for (int j = 0; j < K; ++j) {
                float best = -INFINITY; int arg = 0;
                for (int i = 0; i < K; ++i) {
                    float c = dl[(size_t)(t - 1) * K + i] + sl[(size_t)i * K + j];
                    if (c > best) { best = c; arg = i; }
                }
                dl[(size_t)t * K + j] = best + lb[(size_t)t * K + j];
                ps[(size_t)t * K + j] = arg;
            }
        }
        int s = 0; float bv = dl[(size_t)(T - 1) * K];
        for (int j = 1; j < K; ++j) if (dl[(size_t)(T - 1) * K + j] > bv) { bv = dl[(size_t)(T - 1) * K + j]; s = j; }
        states[(size_t)b * T + T - 1] = s;
        for (int t = T - 1; t >= 1; --t) { s = ps[(size_t)t * K + s]; states[(size_t)b * T + t - 1] = s; }
    }
}

/* neural.py:417-431 (forward: slice t-1 into frame t) and :441-459 (backward: slice t out of frame t), posterior :396-399 */
ORC_API void orc_tv_forward_backward_f64(const double *logb, const double *logT, const double *logp0, int B, int T, int K,
                                         double *log_alpha, double *log_beta, double *gamma, double *loglik) {
    double *tmp = (double *)malloc(sizeof(double) * (size_t)K);
    for (int b = 0; b < B; ++b) {
        const double *lb = logb + (size_t)b * T * K;
        const double *lt = logT + (size_t)b * T * K * K;
        double *la = log_alpha + (size_t)b * T * K, *lbe = log_beta + (size_t)b * T * K, *g = gamma + (size_t)b * T * K;
        for (int j = 0; j < K; ++j) la[j] = logp0[j] + lb[j];
        for (int t = 1; t < T; ++t) {
            const double *sl = lt + (size_t)(t - 1) * K * K;
            for (int j = 0; j < K; ++j) {
                double acc = -INFINITY;
                for (int i = 0; i < K; ++i) acc = lse2(acc, la[(size_t)(t - 1) * K + i] + sl[(size_t)i * K + j]);
                la[(size_t)t * K + j] = acc + lb[(size_t)t * K + j];
            }
        }
        for (int i = 0; i < K; ++i) lbe[(size_t)(T - 1) * K + i] = 0.0;
        for (int t = T - 2; t >= 0; --t) {
            const double *sl = lt + (size_t)t * K * K;
            for (int i = 0; i < K; ++i) {
                double acc = -INFINITY;
                for (int j = 0; j < K; ++j) acc = lse2(acc, sl[(size_t)i * K + j] + lb[(size_t)(t + 1) * K + j] + lbe[(size_t)(t + 1) * K + j]);
                lbe[(size_t)t * K + i] = acc;
            }
        }
        for (int t = 0; t < T; ++t) {
            double z = -INFINITY;
            for (int k = 0; k < K; ++k) { tmp[k] = la[(size_t)t * K + k] + lbe[(size_t)t * K + k]; z = lse2(z, tmp[k]); }
            for (int k = 0; k < K; ++k) g[(size_t)t * K + k] = exp(tmp[k] - z);
        }
        if (loglik) {
            double z = -INFINITY;
            for (int k = 0; k < K; ++k) z = lse2(z, la[(size_t)(T - 1) * K + k]);
            loglik[b] = z;
        }
    }
    free(tmp);
}
