// autograd_kernels.cu -- backward passes of the training callers, sm_100a.
//
// The reference gets its gradients from torch autograd through the T per-step ATen ops of hmm.py:95-126.  With the recursion in one
// launch the backward pass is written out.  For a loss L(gamma) with G = dL/dgamma and
//     gamma_t = softmax_k(log alpha_t(k) + log beta_t(k))                                   (pytorch_hmm/hmm.py:120-126)
// let h_t(k) = gamma_t(k) (G_t(k) - sum_j gamma_t(j) G_t(j)) be dL/d(log alpha_t(k) + log beta_t(k)).  Reverse mode through
//     log alpha_t(j) = LSE_i(log alpha_{t-1}(i) + log P(i,j)) + log b_t(j)                   (hmm.py:98-101)
//     log beta_t(i)  = LSE_j(log P(i,j) + log b_{t+1}(j) + log beta_{t+1}(j))                (hmm.py:113-117)
// gives two more sweeps over time with K x K work per step:
//     abar_{T-1} = h_{T-1};   abar_t(i) = h_t(i) + sum_j abar_{t+1}(j) w_t(i,j),   w_t(i,j) = alpha_t(i) P(i,j) / sum_i' alpha_t(i') P(i',j)
//     bbar_0 = h_0;           c_{t+1}(j) = sum_i bbar_t(i) v_t(i,j),  bbar_{t+1} = h_{t+1} + c_{t+1},
//                             v_t(i,j) = P(i,j) u_{t+1}(j) / sum_j' P(i,j') u_{t+1}(j'),  u = b .* beta
//     dL/d log b_t(j) = abar_t(j) + c_t(j);   dL/d log p0 = abar_0;   dL/d log P(i,j) = sum_t abar_{t+1}(j) w_t(i,j) + bbar_t(i) v_t(i,j)
// w and v are invariant to the per-frame scaling of alpha / beta, so the scaled vectors the forward pass left in its workspace are
// used as they are; the "bar" vectors sum to zero and are pushed through column- / row-stochastic maps, so fp32 needs no scaling.
// One warp per sequence, lane = state (K <= 32): this is the training caller's path, not the inference hot path.
#include "recursion_smallk.cuh"

namespace hmmb200 {

template <int KP>
__global__ void __launch_bounds__(32) posterior_backward_kernel(const float *emis, int mode, float eps, const float *trans,
                                                                const float *ws_a, const float *ws_b, const float *gamma,
                                                                const float *G, int B, int T, int K,
                                                                float *grad_logb, double *grad_logP, double *grad_logp0) {
    __shared__ double gP_s[32 * 32];
    const int b = blockIdx.x, lane = threadIdx.x;
    const bool ok = lane < K;
    for (int i = lane; i < K * K; i += 32) gP_s[i] = 0.0;
    __syncwarp();
    float col[KP], row[KP];                                   // column `lane` and row `lane` of P
#pragma unroll
    for (int i = 0; i < KP; ++i) {
        col[i] = (ok && i < K) ? trans[i * K + lane] : 0.f;
        row[i] = (ok && i < K) ? trans[lane * K + i] : 0.f;
    }
    const size_t base = (size_t)b * T * K;
    auto h_at = [&](int t) {
        const float g = ok ? gamma[base + (size_t)t * K + lane] : 0.f;
        const float Gv = ok ? G[base + (size_t)t * K + lane] : 0.f;
        float dot = g * Gv;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(FULL_MASK, dot, o);
        return g * (Gv - dot);
    };
    auto bprob = [&](int t) {                                  // emission of frame t in probability form, up to a per-frame scale
        const float e = ok ? emis[base + (size_t)t * K + lane] : 0.f;
        if (mode == HMMB200_EMIS_PROB_FLOOR) return ok ? e + eps : 0.f;
        if (mode == HMMB200_EMIS_LOG_EXP_FLOOR) return ok ? expf(e) + eps : 0.f;
        float mx = ok ? e : -INFINITY;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(FULL_MASK, mx, o));
        if (!(mx > -INFINITY)) mx = 0.f;
        return ok ? expf(e - mx) + ((mode == HMMB200_EMIS_LOG_NORM_FLOOR) ? eps : 0.f) : 0.f;
    };
    float accP[KP];                                            // lane j: sum_t abar_{t+1}(j) w_t(i,j) for every i
#pragma unroll
    for (int i = 0; i < KP; ++i) accP[i] = 0.f;

    // ---- reverse sweep: abar ----
    float abar = h_at(T - 1);
    if (ok) grad_logb[base + (size_t)(T - 1) * K + lane] = abar;
    for (int t = T - 2; t >= 0; --t) {
        const float a = ok ? ws_a[base + (size_t)t * K + lane] : 0.f;
        float den = 0.f, av[KP];
#pragma unroll
        for (int i = 0; i < KP; ++i) { av[i] = __shfl_sync(FULL_MASK, a, i); den = fmaf(av[i], col[i], den); }
        const float q = (ok && den > 0.f) ? abar / den : 0.f;              // abar_{t+1}(j) / sum_i alpha_t(i) P(i,j)
#pragma unroll
        for (int i = 0; i < KP; ++i) accP[i] = fmaf(av[i] * col[i], q, accP[i]);
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < KP; ++j) s = fmaf(row[j], __shfl_sync(FULL_MASK, q, j), s);
        abar = h_at(t) + a * s;
        if (ok) grad_logb[base + (size_t)t * K + lane] = abar;
    }
    if (ok && grad_logp0) atomicAdd(grad_logp0 + lane, (double)abar);
    if (ok) {
#pragma unroll
        for (int i = 0; i < KP; ++i) if (i < K) gP_s[i * K + lane] += (double)accP[i];
    }
    __syncwarp();

    // ---- forward sweep: bbar ----  lane i accumulates sum_t bbar_t(i) v_t(i,j) for every j
#pragma unroll
    for (int i = 0; i < KP; ++i) accP[i] = 0.f;
    float bbar = h_at(0);
    for (int t = 0; t + 1 < T; ++t) {
        const float u = bprob(t + 1) * (ok ? ws_b[base + (size_t)(t + 1) * K + lane] : 0.f);
        float den = 0.f, uv[KP];
#pragma unroll
        for (int j = 0; j < KP; ++j) { uv[j] = __shfl_sync(FULL_MASK, u, j); den = fmaf(row[j], uv[j], den); }
        const float q = (ok && den > 0.f) ? bbar / den : 0.f;              // bbar_t(i) / sum_j P(i,j) u_{t+1}(j)
#pragma unroll
        for (int j = 0; j < KP; ++j) accP[j] = fmaf(row[j] * uv[j], q, accP[j]);
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < KP; ++i) s = fmaf(col[i], __shfl_sync(FULL_MASK, q, i), s);
        const float c = u * s;                                             // c_{t+1}(j)
        if (ok) grad_logb[base + (size_t)(t + 1) * K + lane] += c;
        bbar = h_at(t + 1) + c;
    }
    if (ok) {
#pragma unroll
        for (int j = 0; j < KP; ++j) if (j < K) gP_s[lane * K + j] += (double)accP[j];
    }
    __syncwarp();
    if (grad_logP) for (int i = lane; i < K * K; i += 32) if (gP_s[i] != 0.0) atomicAdd(grad_logP + i, gP_s[i]);
}

}  // namespace hmmb200

using namespace hmmb200;

// dL/d log b [B,T,K] (overwritten), dL/d log P [K,K] and dL/d log p0 [K] (both ADDED to, double) for a loss on the posteriors.
// Arguments as for hmmb200_xi_sum_f32 plus gamma (the forward pass's posteriors) and G = dL/dgamma.
HMMB200_EXPORT int hmmb200_posterior_backward_f32(const float *emis, int emis_mode, float floor_eps, const float *trans_prob,
                                                  const void *fb_workspace, const float *gamma, const float *grad_gamma,
                                                  int B, int T, int K, float *grad_logb, double *grad_logP, double *grad_logp0,
                                                  void *stream) {
    if (B < 0 || T < 0 || K <= 0) return set_error(HMMB200_EINVAL, "posterior_backward: bad shape");
    if (B == 0 || T == 0) return HMMB200_OK;
    if (K > 32) return set_error(HMMB200_EUNSUPPORTED, "posterior_backward: K <= 32 (got %d)", K);
    if (!emis || !trans_prob || !fb_workspace || !gamma || !grad_gamma || !grad_logb) return set_error(HMMB200_EINVAL, "posterior_backward: null argument");
    if (int rc = require_sm100()) return rc;
    const size_t n = (size_t)B * T;
    const float *ws_a = (const float *)fb_workspace;
    const float *ws_b = (const float *)((const uint8_t *)fb_workspace + align256(n * K * sizeof(float)));
    cudaStream_t s = (cudaStream_t)stream;
    const int kp = pad4(K);
#define PB_CASE(N) if (kp <= N) { posterior_backward_kernel<N><<<(unsigned)B, 32, 0, s>>>(emis, emis_mode, floor_eps, trans_prob, ws_a, ws_b, gamma, grad_gamma, B, T, K, grad_logb, grad_logP, grad_logp0); return check_launch("posterior_backward_kernel"); }
    PB_CASE(4) PB_CASE(8) PB_CASE(12) PB_CASE(16) PB_CASE(24)
    PB_CASE(32)
#undef PB_CASE
    return HMMB200_OK;
}
