/*
 * hmm_b200.h -- C ABI of the B200-native HMM inference engine (libhmm_b200.so).
 *
 * This is the drop-in boundary for pytorch_hmm's hot path.  The reference has no FFI (it is pure Python on
 * torch); each entry point below names the reference routine whose arithmetic it replaces (paths relative to
 * the reference repository root), and pytorch_hmm_b200/*.py binds them with ctypes behind the reference's
 * class API.  INTEGRATION.md shows the stub a maintainer of the reference would add.
 *
 * Conventions (all entry points):
 *   - plain C types only: device pointers, sizes, a CUDA stream passed as void* (cudaStream_t), no torch types;
 *   - every pointer is a DEVICE pointer unless the name ends in _host; tensors are dense row-major fp32;
 *   - no allocation, no ownership transfer, no host synchronisation, no global mutable state: safe to call
 *     from several host threads on different streams; work is enqueued on `stream` and the call returns;
 *   - return 0 on success, <0 on error (HMMB200_E*); hmmb200_last_error() gives the thread's last message;
 *   - there is NO CPU fallback: on a machine without an sm_100 device every compute call returns
 *     HMMB200_ENODEVICE.
 */
#ifndef HMM_B200_H
#define HMM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HMMB200_ABI_VERSION 1

#define HMMB200_OK            0
#define HMMB200_EINVAL       -1   /* bad argument (null pointer, non-positive size, ...)            */
#define HMMB200_EUNSUPPORTED -2   /* shape outside what this entry point covers (e.g. K > 32)        */
#define HMMB200_ELAUNCH      -3   /* CUDA reported an error at launch                                */
#define HMMB200_EWORKSPACE   -4   /* workspace too small (see the *_workspace_bytes query)           */
#define HMMB200_ENODEVICE    -5   /* no CUDA device of compute capability 10.x                       */

/* How an emission tensor `emis[B,T,K]` is to be read by the recursions. */
#define HMMB200_EMIS_LOG            0  /* log b_t(k) as is (NeuralHMM-style log emissions; mixture_gaussian.py:312-324)      */
#define HMMB200_EMIS_PROB_FLOOR     1  /* probabilities: log b = log(p + eps)             (hmm.py:86, :152)                   */
#define HMMB200_EMIS_LOG_NORM_FLOOR 2  /* log-lik l: log b = log(exp(l - max_k l) + eps)  (per-frame      */
                                       /*   max-normalised probabilities fed to HMMPyTorch)                                   */
#define HMMB200_EMIS_LOG_EXP_FLOOR  3  /* log-lik l: log b = log(exp(l) + eps)            (hmm_layer.py:336-337 then hmm.py:86) */

int         hmmb200_abi_version(void);
const char *hmmb200_last_error(void);
/* 0 if device `ordinal` (or the current device when ordinal < 0) is compute capability 10.x. */
int         hmmb200_device_check(int ordinal);

/* ---------------------------------------------------------------------------------------------------------
 * Emission: diagonal-Gaussian / GMM log-likelihood.
 *   replaces  MixtureGaussianHMMLayer.get_observation_log_probs   pytorch_hmm/mixture_gaussian.py:157-214
 *             GaussianHMMLayer._compute_gaussian_log_probs        pytorch_hmm/hmm_layer.py:270-323 ('diag')
 *             HSMMLayer.get_observation_log_probs                 pytorch_hmm/hsmm.py:181-206
 *
 * hmmb200_gmm_pack_f32 turns the layer parameters into the kernel's packed form (once per parameter update):
 *   means, log_vars : [K, C, D]   variance = exp(log_var_scale * log_vars)  (1 for log_vars, 2 for log_scales)
 *   log_weights     : [K, C] log mixture weights, or NULL for a single Gaussian per state (C must be 1)
 *   packed          : hmmb200_gmm_packed_floats(K, C, D) floats
 * hmmb200_gmm_emission_f32:
 *   x [n_frames, D] -> logb [n_frames, K];   logb_k = own_lse_c( log w_kc + log N(x | mu_kc, var_kc) ),
 *   where own_lse is the reference's private logsumexp (max, log(clamp(sum, 1e-8)) + max).
 * --------------------------------------------------------------------------------------------------------- */
size_t hmmb200_gmm_packed_floats(int K, int C, int D);
int    hmmb200_gmm_pack_f32(const float *means, const float *log_vars, float log_var_scale,
                            const float *log_weights, int K, int C, int D, float *packed, void *stream);
int    hmmb200_gmm_emission_f32(const float *x, const float *packed, int64_t n_frames, int K, int C, int D,
                                float *logb, void *stream);
/* Setup-time query (the ONE entry point that synchronises `stream`): 1 if `packed` runs on the tcgen05 kernel, 0 if it needs the
 * fp32 kernel (shape outside the TMEM budget, or weights outside the fp16 range).  When it returns 1, hmmb200_gmm_emission_tc_f32
 * may be used per batch: same result as hmmb200_gmm_emission_f32, one kernel launch instead of two. */
int    hmmb200_gmm_pack_on_tensor_cores(const float *packed, int K, int C, int D, void *stream);
int    hmmb200_gmm_emission_tc_f32(const float *x, const float *packed, int64_t n_frames, int K, int C, int D,
                                   float *logb, void *stream);

/* Full-covariance GMM emission.
 *   replaces  MixtureGaussianHMMLayer._full_gaussian_log_probs   pytorch_hmm/mixture_gaussian.py:216-240
 * The triangular solve L y = x - mu of the reference is folded into the parameters by the host (once per parameter update):
 *   W [K*C, D, DP] = L^-1 per component, lower triangular, rows zero-padded to DP = (D + 3) & ~3 floats, 16-byte aligned;
 *   cvec [K*C, D] = -W mu;   cst [K*C] = log w - 0.5 (log det + D log 2 pi).
 *   x [n_frames, D] -> logb [n_frames, K] (the reference's private log-sum-exp over components), comp [n_frames, K*C] or NULL. */
int    hmmb200_gmm_emission_full_f32(const float *x, const float *W, const float *cvec, const float *cst, int64_t n_frames,
                                     int K, int C, int D, float *logb, float *comp, void *stream);

/* ---------------------------------------------------------------------------------------------------------
 * Forward-backward (K <= 32: warp-per-sequence sweeps; 32 < K <= 512: cluster kernels, BASELINE config 5; 512 < K <= 2048: one launch per frame).
 *   replaces  HMMPyTorch.forward_backward / compute_likelihood    pytorch_hmm/hmm.py:66-130, :186-211
 *
 *   emis [B,T,K] read according to emis_mode / floor_eps (see HMMB200_EMIS_*).
 *   trans_prob [K,K], init_prob [K]: the EFFECTIVE probabilities exp(log_P), exp(log_p0), i.e. P + 1e-8 and
 *     p0 + 1e-8 for HMMPyTorch (hmm.py:42,55).  The recursion runs in scaled-probability space.
 *   outputs (any may be NULL): gamma [B,T,K] posterior (hmm.py:120-126); fwd_prob, bwd_prob [B,T,K] =
 *     exp(log alpha), exp(log beta) as the reference returns them (hmm.py:127-128; they underflow to 0);
 *     log_alpha, log_beta [B,T,K]; loglik [B] = logsumexp_k log alpha_{T-1} (the true value; the reference's
 *     saturating compute_likelihood is derived from fwd_prob on the host).
 *   add_rowmax: for EMIS_LOG_NORM_FLOOR, also add sum_t max_k l_t back into loglik (0 = reference behaviour).
 *   workspace: hmmb200_fb_workspace_bytes(B, T, K) bytes of device scratch.
 * --------------------------------------------------------------------------------------------------------- */
size_t hmmb200_fb_workspace_bytes(int B, int T, int K);
int    hmmb200_forward_backward_f32(const float *emis, int emis_mode, float floor_eps, int add_rowmax,
                                    const float *trans_prob, const float *init_prob, int B, int T, int K,
                                    float *gamma, float *fwd_prob, float *bwd_prob,
                                    float *log_alpha, float *log_beta, float *loglik,
                                    void *workspace, size_t workspace_bytes, void *stream);

/* Time-parallel variant for LONG sequences at SMALL batch (K <= 32): same arguments, same results (1e-4), different schedule.
 * The step alpha_t = alpha_{t-1} P diag(b_t) is associative: time is cut into segments, the K x K segment products are computed
 * in parallel (K x the sequential arithmetic), chained by one warp per sequence, and every segment is then filled from its true
 * boundary vectors.  Pays off when B is far below the ~300 sequences the sequential sweeps need to fill a B200.
 *   workspace: hmmb200_fb_scan_workspace_bytes(B, T, K) bytes. */
size_t hmmb200_fb_scan_workspace_bytes(int B, int T, int K);
int    hmmb200_forward_backward_scan_f32(const float *emis, int emis_mode, float floor_eps, int add_rowmax,
                                         const float *trans_prob, const float *init_prob, int B, int T, int K,
                                         float *gamma, float *fwd_prob, float *bwd_prob,
                                         float *log_alpha, float *log_beta, float *loglik,
                                         void *workspace, size_t workspace_bytes, void *stream);

/* ---------------------------------------------------------------------------------------------------------
 * Viterbi.  K <= 32: packed uint8 backpointers in shared memory, chunk-parallel on-device traceback.
 *           32 < K <= 512: cluster kernel; the traceback recomputes backpointers on the path, the full table is written only
 *           when psi is given.
 *   replaces  HMMPyTorch.viterbi_decode                           pytorch_hmm/hmm.py:132-184
 *             MixtureGaussianHMMLayer._viterbi_decode             pytorch_hmm/mixture_gaussian.py:290-338
 *
 *   delta_0 = log_init + log b_0;  (m, psi_t[j]) = max_i(delta_{t-1}[i] + log_trans[i][j]) with the LOWEST i on
 *   ties;  delta_t = m + log b_t (two fp32 roundings in that order);  s_{T-1} = first argmax; s_t = psi_{t+1}[s_{t+1}].
 *   With emis_mode == HMMB200_EMIS_LOG the result is bit-identical to the reference given the same fp32 inputs.
 *   outputs (NULL allowed except states): delta [B,T,K]; psi [B,T,K] packed backpointers, uint8 for K <= 256 and uint16 above
 *     (psi_0 = 0); states [B,T] int64;
 *     score [B] = max_k delta_{T-1}.
 *   workspace: hmmb200_viterbi_workspace_bytes(B, T, K) bytes (0 when the backpointers fit in shared memory).
 * --------------------------------------------------------------------------------------------------------- */
size_t hmmb200_viterbi_workspace_bytes(int B, int T, int K);
int    hmmb200_viterbi_f32(const float *emis, int emis_mode, float floor_eps,
                           const float *log_trans, const float *log_init, int B, int T, int K,
                           float *delta, void *psi, int64_t *states, float *score,
                           void *workspace, size_t workspace_bytes, void *stream);

/* ---------------------------------------------------------------------------------------------------------
 * Forward-backward AND Viterbi on the same emissions in one pass (K <= 32: ONE launch hosting the forward sweep, the backward
 * sweep and the Viterbi recursion of a group of sequences in one CTA, each chain on its own SM sub-partition; other shapes: the two
 * entry points above, back to back).
 *   replaces  the whole-batch pass the reference times: hmm.forward_backward(obs) then hmm.viterbi_decode(obs)
 *             pytorch_hmm/examples/benchmark.py:120-196; HMMLayer.forward train / eval, pytorch_hmm/hmm_layer.py:119-131
 *   The two recursions may read `emis` differently (fb_mode / vit_mode): HMMPyTorch semantics on max-normalised probabilities for
 *   the posteriors and MixtureGaussianHMMLayer._viterbi_decode on the raw log-emissions is the BASELINE configs[1] step.
 *   Arguments and outputs as for hmmb200_forward_backward_f32 and hmmb200_viterbi_f32 (same NULL rules).
 *   flags: HMMB200_FUSED_PDL -- launch with programmatic stream serialisation: the kernel's set-up overlaps the tail of the
 *     preceding kernel on `stream` (normally the emission kernel writing `emis`); the caller promises that this preceding kernel
 *     writes none of trans_prob / init_prob / log_trans / log_init.
 *   workspace: hmmb200_fb_viterbi_workspace_bytes(B, T, K) bytes.
 * --------------------------------------------------------------------------------------------------------- */
#define HMMB200_FUSED_PDL 1
#define HMMB200_FUSED_BF16_OUT 2   /* gamma / fwd_prob / bwd_prob point to bfloat16 arrays (K <= 32; log_alpha / log_beta must be NULL):   */
                                   /* the posterior kernel is a pure streaming write -- half the bytes (north star: bf16/fp32 outputs)     */
size_t hmmb200_fb_viterbi_workspace_bytes(int B, int T, int K);
int    hmmb200_fb_viterbi_f32(const float *emis, int fb_mode, int vit_mode, float floor_eps, int add_rowmax,
                              const float *trans_prob, const float *init_prob,
                              const float *log_trans, const float *log_init, int B, int T, int K,
                              float *gamma, float *fwd_prob, float *bwd_prob, float *log_alpha, float *log_beta,
                              float *loglik, float *delta, void *psi, int64_t *states, float *score,
                              void *workspace, size_t workspace_bytes, int flags, void *stream);

/* ---------------------------------------------------------------------------------------------------------
 * Recursions with TIME-VARYING transitions (K <= 32): the NeuralHMM form of the path.
 *   replaces  NeuralHMM._forward_algorithm / _backward_algorithm    pytorch_hmm/neural.py:403-461
 *             NeuralHMM.viterbi_decode (recursion + backtrack)       pytorch_hmm/neural.py:463-511
 *   log_emis [B,T,K] log-emissions, used as they are (no floor).  Slice t of the [B,T,K,K] transition tensor carries frame t to
 *   frame t+1 (neural.py:424, :448, :490); slice T-1 is never read.  Forward-backward takes PROBABILITIES trans_prob =
 *   exp(log_transition_probs) and init_prob [K] (the recursion runs in scaled-probability space); Viterbi takes the logs.
 *   Outputs and NULL rules as for the fixed-transition entry points; forward-backward workspace: hmmb200_fb_workspace_bytes(B,T,K);
 *   Viterbi workspace: hmmb200_tv_viterbi_workspace_bytes (0 when the backpointers fit in shared memory).
 * --------------------------------------------------------------------------------------------------------- */
int    hmmb200_tv_forward_backward_f32(const float *log_emis, const float *trans_prob, const float *init_prob,
                                       int B, int T, int K, float *gamma, float *fwd_prob, float *bwd_prob,
                                       float *log_alpha, float *log_beta, float *loglik,
                                       void *workspace, size_t workspace_bytes, void *stream);
size_t hmmb200_tv_viterbi_workspace_bytes(int B, int T, int K);
int    hmmb200_tv_viterbi_f32(const float *log_emis, const float *log_trans, const float *log_init,
                              int B, int T, int K, float *delta, void *psi, int64_t *states, float *score,
                              void *workspace, size_t workspace_bytes, void *stream);

/* ---------------------------------------------------------------------------------------------------------
 * Explicit-duration (semi-Markov) recursions.
 *   hmmb200_hsmm_viterbi_f32 replaces  HSMMLayer._viterbi_decode_single      pytorch_hmm/hsmm.py:245-354
 *                                      SemiMarkovHMM.viterbi_decode          pytorch_hmm/semi_markov.py:455-570
 *   hmmb200_hsmm_forward_f32 replaces  SemiMarkovHMM._unsupervised_forward   pytorch_hmm/semi_markov.py:308-383
 *
 *   frame_logp [B,T,K] per-frame log-emission term; seg_const [K] (or NULL) is added ONCE per segment (SemiMarkovHMM counts the
 *   Gaussian constant per segment, semi_markov.py:422-424; HSMMLayer per frame, hsmm.py:285 -> NULL); log_dur [K,Dm] with column
 *   d-1 for duration d; log_trans [K,K] (self transitions are never taken); log_init [K] or NULL (HSMMLayer: no prior).
 *   Viterbi: delta = ((delta_prev + log_trans) + seg) + log_dur maximised over (s' != s, d') in lexicographic order with a strict
 *   '>' exactly like the reference; sum_order 0 sums seg(t,d,s) in ATen's strided-sum order (bit-identical scores), 1 sequentially.
 *   outputs: states [B,T] int64, score [B];  forward: alpha [B,T,K,Dm] (NULL ok), end_scores [B,T,K] (NULL ok), total [B].
 * --------------------------------------------------------------------------------------------------------- */
size_t hmmb200_hsmm_viterbi_workspace_bytes(int B, int T, int K, int Dm);
int    hmmb200_hsmm_viterbi_f32(const float *frame_logp, const float *seg_const, const float *log_dur,
                                const float *log_trans, const float *log_init, int B, int T, int K, int Dm,
                                int sum_order, int64_t *states, float *score,
                                void *workspace, size_t workspace_bytes, void *stream);
int    hmmb200_hsmm_forward_f32(const float *frame_logp, const float *seg_const, const float *log_dur,
                                const float *log_trans, const float *log_init, int B, int T, int K, int Dm,
                                float *alpha, float *end_scores, float *total, void *stream);
/* Duration-augmented forward-BACKWARD (new: the reference has no HSMM backward pass; BASELINE config 4).  Same model arguments.
 *   gamma [B,T,K] = P(state_t = s | o) summed over all segmentations, total [B] = log p(o);
 *   beta_begin / beta_end [B,T,K] (NULL ok): log p(o_{t..} | a segment of s begins at t) / log p(o_{t+1..} | a segment of s ends at t).
 *   workspace: hmmb200_hsmm_fb_workspace_bytes(B, T, K) bytes. */
size_t hmmb200_hsmm_fb_workspace_bytes(int B, int T, int K);
int    hmmb200_hsmm_forward_backward_f32(const float *frame_logp, const float *seg_const, const float *log_dur,
                                         const float *log_trans, const float *log_init, int B, int T, int K, int Dm,
                                         float *gamma, float *total, float *beta_begin, float *beta_end,
                                         void *workspace, size_t workspace_bytes, void *stream);

/* ---------------------------------------------------------------------------------------------------------
 * Streaming: per-chunk kernels with state carried between calls (one stream per batch row, K <= 32).
 *   hmmb200_greedy_decode_f32 replaces  StreamingHMMProcessor._greedy_decode   pytorch_hmm/streaming.py:267-320
 *     s_t = argmax_j(log_trans[s_{t-1}][j] + logb_t[j]); state_io[b] < 0 marks the first chunk (argmax_j(logb_0[j] - log K)),
 *     and is updated to the chunk's last state.  states [B,T] int64, scores [B,T] (NULL ok) = the winning value.
 *   hmmb200_forward_chunk_f32 is new (the reference's processor has no forward algorithm): the forward recursion over one
 *     chunk; state_alpha [B,K] (filtered distribution), state_loglik [B] (double) and started [B] carry the stream state;
 *     filtered [B,T,K] (NULL ok) receives p(state_t | o_1..t).  Chunked calls equal one unchunked pass.
 * --------------------------------------------------------------------------------------------------------- */
int    hmmb200_greedy_decode_f32(const float *logb, const float *log_trans, int B, int T, int K,
                                 int32_t *state_io, int64_t *states, float *scores, void *stream);
int    hmmb200_forward_chunk_f32(const float *emis, int emis_mode, float floor_eps, const float *trans_prob,
                                 const float *init_prob, int B, int T, int K, float *state_alpha,
                                 double *state_loglik, int32_t *started, float *filtered, void *stream);

/* ---------------------------------------------------------------------------------------------------------
 * Baum-Welch E-step statistics for a GMM-HMM (new functionality; formulas docs/01_hmm_theory.md:196-227 + the
 * standard Gaussian-mixture extension).  Multi-GPU: each rank accumulates its shard of utterances, the host all-reduces
 * the stats vector (NCCL) once per EM iteration.
 *   hmmb200_gmm_components_f32: comp [n_frames, K*C] = log w_kc + log N(x | mu_kc, var_kc)  (per-component, no log-sum-exp)
 *   hmmb200_bw_accumulate_f32 : ADDS into stats (double, hmmb200_bw_stats_doubles(K,C,D) values, zeroed by the caller):
 *       gamma1[K] | xi[K,K] | occ[K,C] | sx[K,C,D] | sxx[K,C,D]
 *     inputs: x [B,T,D]; comp; logb [B,T,K] (= LSE_c comp); gamma [B,T,K]; emis/emis_mode/floor_eps/trans_prob exactly as passed
 *     to hmmb200_forward_backward_f32, and that call's workspace (it still holds the scaled forward/backward vectors).
 * --------------------------------------------------------------------------------------------------------- */
int    hmmb200_gmm_components_f32(const float *x, const float *packed, int64_t n_frames, int K, int C, int D,
                                  float *comp, void *stream);
/* logb [n_frames, K] and comp [n_frames, K*C] in ONE pass over x (the tcgen05 emission kernel writes both from its epilogue). */
int    hmmb200_gmm_emission_components_f32(const float *x, const float *packed, int64_t n_frames, int K, int C, int D,
                                           float *logb, float *comp, void *stream);
size_t hmmb200_bw_stats_doubles(int K, int C, int D);
int    hmmb200_bw_accumulate_f32(const float *x, const float *comp, const float *logb, const float *gamma,
                                 const float *emis, int emis_mode, float floor_eps, const float *trans_prob,
                                 const void *fb_workspace, int B, int T, int K, int C, int D,
                                 double *stats, void *stream);

/* Transition / initial-state statistics alone, optionally weighted per sequence (seq_weights [B] or NULL = 1):
 *   xi[K,K] += sum_b w_b sum_t xi_t(i,j);  gamma1[K] += sum_b w_b gamma_0(k)   (gamma1 may be NULL).
 * Arguments as for hmmb200_bw_accumulate_f32.  With w_b = d loss / d loglik_b these are d loss / d log P and d loss / d log p0:
 * the backward pass of the training callers (HMMLayer.compute_loss, pytorch_hmm/hmm_layer.py:144-173). */
int    hmmb200_xi_sum_f32(const float *emis, int emis_mode, float floor_eps, const float *trans_prob,
                          const void *fb_workspace, const float *seq_weights, int B, int T, int K,
                          double *xi, double *gamma1, void *stream);

/* Backward pass of a loss on the POSTERIORS (HMMLayer.forward in training mode feeding a loss; the supervised cross-entropy of
 * HMMLayer.compute_loss, pytorch_hmm/hmm_layer.py:161-167): with grad_gamma = dL/dgamma [B,T,K] and the forward pass's gamma and
 * workspace, writes dL/d log b [B,T,K] and ADDS dL/d log P [K,K], dL/d log p0 [K] (double; either may be NULL).  K <= 32. */
int    hmmb200_posterior_backward_f32(const float *emis, int emis_mode, float floor_eps, const float *trans_prob,
                                      const void *fb_workspace, const float *gamma, const float *grad_gamma,
                                      int B, int T, int K, float *grad_logb, double *grad_logP, double *grad_logp0, void *stream);

/* Weighted Gaussian-mixture statistics alone: with weight[n,K] per (frame, state) and r = exp(comp - logb) the component
 * responsibilities,  occ[K,C] += sum_n weight r,  sx[K,C,D] += sum_n weight r x,  sxx[K,C,D] += sum_n weight r x^2  (double).
 * With weight = dL/d log b these are the sufficient statistics of the emission kernel's backward pass (d/d mu, d/d log var, d/d log w). */
int    hmmb200_gmm_stats_f32(const float *x, const float *comp, const float *logb, const float *weight, int64_t n_frames,
                             int K, int C, int D, double *occ, double *sx, double *sxx, void *stream);

/* ---------------------------------------------------------------------------------------------------------
 * Alignment utilities: CTC trellises and dynamic time warping.
 *   replaces  ctc_forward_algorithm    pytorch_hmm/alignment/ctc.py:32-121   (direction 0: log alpha, log-likelihood)
 *             ctc_backward_algorithm   pytorch_hmm/alignment/ctc.py:124-199  (direction 1: log beta)
 *             compute_dtw_path         pytorch_hmm/alignment/dtw.py:47-153
 * hmmb200_ctc_trellis_f32:
 *   log_probs [T,B,C], targets [B,L] int64, input_lengths / target_lengths [B] int64 (all device pointers);
 *   table [B,T,2L+1] or NULL (every element is written: -inf where the reference leaves its initial value);
 *   loglik [B] or NULL (forward only).
 * hmmb200_dtw_f32:
 *   dist [n_pairs,N,M] -> cost [n_pairs,N,M]; dir_ws [n_pairs*N*M] bytes of scratch; path_i / path_j [n_pairs, N+M-1] int64
 *   (the first path_len[p] entries of a row are the path from (0,0) to (N-1,M-1)); step_pattern 0 symmetric, 1 asymmetric,
 *   2 rabiner_juang.  Costs, and therefore paths, are bit-identical to the reference's (fp32 adds and minima only).
 * --------------------------------------------------------------------------------------------------------- */
int    hmmb200_ctc_trellis_f32(int direction, const float *log_probs, const int64_t *targets, const int64_t *input_lengths,
                               const int64_t *target_lengths, int blank, int T, int B, int C, int L,
                               float *table, float *loglik, void *stream);
int    hmmb200_dtw_f32(const float *dist, int n_pairs, int N, int M, int step_pattern, float *cost, void *dir_ws,
                       int64_t *path_i, int64_t *path_j, int *path_len, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* HMM_B200_H */
