// common.cuh -- shared helpers for libhmm_b200.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/hmm_b200.h"

#define HMMB200_EXPORT extern "C" __attribute__((visibility("default")))

namespace hmmb200 {

// Thread-local last-error string (no global mutable state shared between host threads).
char *last_error_buf();
int set_error(int code, const char *fmt, ...);

// Called after every launch: converts a launch error into HMMB200_ELAUNCH without synchronising.
int check_launch(const char *what);

// 0 when the current device is compute capability 10.x (cached per device ordinal).
int require_sm100();

constexpr unsigned FULL_MASK = 0xffffffffu;

// Smallest power of two >= k, clamped to [4, 32]: lanes per sequence in the small-K recursions.
inline int group_lanes(int K) { int g = 4; while (g < K) g <<= 1; return g; }
// Number of broadcast slots (multiple of 4, >= K).
inline int pad4(int K) { return (K + 3) & ~3; }

// emission_tc.cu (tcgen05 path)
bool tc_shape_ok(int K, int C, int D);
size_t tc_floats(int K, int C, int D);
int launch_pack_tc(const float *means, const float *log_vars, float scale, const float *logw, int K, int C, int D,
                   float *tc, cudaStream_t s);
// returns 0 launched, >0 not applicable (caller uses the fp32 kernel only), <0 error
int launch_emission_tc(const float *x, const float *tc, const float *packed32, int64_t n_frames, int K, int C, int D,
                       float *logb, cudaStream_t s, float *comp = nullptr);
// emission.cu: log b (and, with comp != null, the per-component values from the tcgen05 kernel).  *tc_flag_out receives the device
// flag that is 1 when the tensor-core kernel did the work (null when it was not launched).
int gmm_emission_dispatch(const float *x, const float *packed, int64_t n_frames, int K, int C, int D, float *logb, float *comp,
                          cudaStream_t s, int tc_known, const float **tc_flag_out);

// recursion_largek.cu (32 < K <= 512: cluster kernels)
bool largek_shape_ok(int K);
size_t largek_fb_workspace_bytes(int B, int T, int K);
size_t largek_viterbi_workspace_bytes(int B, int T, int K);
int largek_forward_backward(const float *emis, int emis_mode, float floor_eps, int add_rowmax, const float *trans_prob,
                            const float *init_prob, int B, int T, int K, float *gamma, float *fwd_prob, float *bwd_prob,
                            float *log_alpha, float *log_beta, float *loglik, void *workspace, cudaStream_t s);
int largek_viterbi(const float *emis, int emis_mode, float floor_eps, const float *log_trans, const float *log_init,
                   int B, int T, int K, float *delta, void *psi, int64_t *states, float *score, void *workspace, cudaStream_t s);

int largek_fb_viterbi(const float *emis, int fb_mode, int vit_mode, float floor_eps, int add_rowmax, const float *trans_prob,
                      const float *init_prob, const float *log_trans, const float *log_init, int B, int T, int K,
                      float *gamma, float *fwd_prob, float *bwd_prob, float *log_alpha, float *log_beta, float *loglik,
                      float *delta, void *psi, int64_t *states, float *score, void *fb_workspace, void *vit_workspace, cudaStream_t s);

}  // namespace hmmb200
