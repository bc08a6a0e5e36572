// streaming.cu -- per-chunk kernels of the streaming path (sm_100a).
//
//   greedy_decode_kernel   s_t = argmax_j( logA[s_{t-1}][j] + log b_t[j] ), continuing from the previous chunk's last state
//                          (pytorch_hmm/streaming.py:292-308; first chunk: argmax_j(log b_0[j] - log K)).  First-index ties.
//   forward_chunk_kernel   the forward recursion over one chunk with the filtered state vector carried between calls
//                          (new: the reference's processor has no forward algorithm; chunked == unchunked by construction).
// One warp per stream, lane = state (K <= 32).  Chunks are short (tens to hundreds of frames), so these are plain
// warp-shuffle recursions rather than the pipelined kernels of recursion_smallk.cu.
#include "common.cuh"

namespace hmmb200 {

__global__ void __launch_bounds__(32) greedy_decode_kernel(const float *logb, const float *logA, int B, int T, int K,
                                                           int32_t *state_io, int64_t *states, float *scores) {
    extern __shared__ float A_s[];
    const int b = blockIdx.x, j = threadIdx.x;
    for (int i = j; i < K * K; i += 32) A_s[i] = logA[i];
    __syncwarp();
    int prev = state_io[b];                       // < 0: first chunk of the stream
    const float logK = logf((float)K);
    const float *lb = logb + (size_t)b * T * K;
    for (int t = 0; t < T; ++t) {
        float v = -INFINITY;
        if (j < K) {
            const float e = lb[(size_t)t * K + j];
            v = (prev < 0) ? __fadd_rn(e, -logK) : __fadd_rn(A_s[prev * K + j], e);
        }
        int idx = j;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(FULL_MASK, v, o);
            const int oi = __shfl_xor_sync(FULL_MASK, idx, o);
            if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
        }
        prev = idx;
        if (j == 0) {
            states[(size_t)b * T + t] = idx;
            if (scores) scores[(size_t)b * T + t] = v;
        }
    }
    if (j == 0) state_io[b] = prev;
}

// state_alpha [B,K]: filtered state distribution after the previous chunk (sums to 1), or all zeros before the first
// chunk, in which case init_prob [K] starts the recursion.  state_loglik [B] accumulates log p(o_1..t) in double.
__global__ void __launch_bounds__(32) forward_chunk_kernel(const float *emis, int mode, float eps, const float *trans,
                                                           const float *init, int B, int T, int K,
                                                           float *state_alpha, double *state_loglik, int32_t *started,
                                                           float *filtered) {
    const int b = blockIdx.x, j = threadIdx.x;
    const bool ok = j < K;
    float col[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) col[i] = (ok && i < K) ? trans[i * K + j] : 0.f;
    float a = ok ? state_alpha[(size_t)b * K + j] : 0.f;
    double ll = state_loglik[b];
    bool first = started[b] == 0;
    const float *e = emis + (size_t)b * T * K;
    for (int t = 0; t < T; ++t) {
        float x = ok ? e[(size_t)t * K + j] : 0.f;
        float m = 0.f, bt;
        if (mode == HMMB200_EMIS_PROB_FLOOR) bt = ok ? x + eps : 0.f;
        else if (mode == HMMB200_EMIS_LOG_EXP_FLOOR) bt = ok ? expf(x) + eps : 0.f;
        else {
            float mx = ok ? x : -INFINITY;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(FULL_MASK, mx, o));
            if (!(mx > -INFINITY)) mx = 0.f;
            bt = ok ? expf(x - mx) + ((mode == HMMB200_EMIS_LOG_NORM_FLOOR) ? eps : 0.f) : 0.f;
            m = (mode == HMMB200_EMIS_LOG) ? mx : 0.f;
        }
        float pred;
        if (first) { pred = ok ? init[j] : 0.f; first = false; }
        else {
            pred = 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i) pred = fmaf(__shfl_sync(FULL_MASK, a, i), col[i], pred);
        }
        float v = pred * bt, s = v;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL_MASK, s, o);
        a = (s > 0.f) ? v / s : 0.f;
        ll += (double)logf(s) + (double)m;
        if (filtered && ok) filtered[((size_t)b * T + t) * K + j] = a;
    }
    if (ok) state_alpha[(size_t)b * K + j] = a;
    if (j == 0) { state_loglik[b] = ll; started[b] = 1; }
}

}  // namespace hmmb200

using namespace hmmb200;

HMMB200_EXPORT int hmmb200_greedy_decode_f32(const float *logb, const float *log_trans, int B, int T, int K,
                                             int32_t *state_io, int64_t *states, float *scores, void *stream) {
    if (B < 0 || T < 0 || K <= 0) return set_error(HMMB200_EINVAL, "greedy_decode: bad shape");
    if (B == 0 || T == 0) return HMMB200_OK;
    if (K > 32) return set_error(HMMB200_EUNSUPPORTED, "greedy_decode: K <= 32 (got %d)", K);
    if (!logb || !log_trans || !state_io || !states) return set_error(HMMB200_EINVAL, "greedy_decode: null argument");
    if (int rc = require_sm100()) return rc;
    greedy_decode_kernel<<<B, 32, (size_t)K * K * sizeof(float), (cudaStream_t)stream>>>(logb, log_trans, B, T, K, state_io, states, scores);
    return check_launch("greedy_decode_kernel");
}

HMMB200_EXPORT int hmmb200_forward_chunk_f32(const float *emis, int emis_mode, float floor_eps, const float *trans_prob,
                                             const float *init_prob, int B, int T, int K, float *state_alpha,
                                             double *state_loglik, int32_t *started, float *filtered, void *stream) {
    if (B < 0 || T < 0 || K <= 0) return set_error(HMMB200_EINVAL, "forward_chunk: bad shape");
    if (B == 0 || T == 0) return HMMB200_OK;
    if (K > 32) return set_error(HMMB200_EUNSUPPORTED, "forward_chunk: K <= 32 (got %d)", K);
    if (!emis || !trans_prob || !init_prob || !state_alpha || !state_loglik || !started)
        return set_error(HMMB200_EINVAL, "forward_chunk: null argument");
    if (emis_mode < 0 || emis_mode > 3) return set_error(HMMB200_EINVAL, "forward_chunk: bad emis_mode %d", emis_mode);
    if (int rc = require_sm100()) return rc;
    forward_chunk_kernel<<<B, 32, 0, (cudaStream_t)stream>>>(emis, emis_mode, floor_eps, trans_prob, init_prob, B, T, K,
                                                            state_alpha, state_loglik, started, filtered);
    return check_launch("forward_chunk_kernel");
}
