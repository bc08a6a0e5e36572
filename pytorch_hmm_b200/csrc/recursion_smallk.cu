// recursion_smallk.cu -- small-K (K <= 32) HMM recursions for sm_100a, warp-specialised.
//
//   fb_sweep_kernel     forward and backward sweeps (blockIdx.y = direction), scaled-probability space.
//                       Replaces the per-time-step ATen launches of pytorch_hmm/hmm.py:95-117.
//   fb_combine_warp_kernel / fb_combine_kernel   posterior / exp(log alpha) / exp(log beta) from the two scaled sweeps
//                       (hmm.py:120-128): coalesced warp-per-32-frames form for K % 4 == 0, scalar form otherwise.
//   viterbi_kernel      max-plus recursion, packed uint8 backpointers in shared memory, chunk-parallel on-device
//                       traceback (hmm.py:159-178; mixture_gaussian.py:312-336).
//
// The time recursion is a latency problem (T dependent steps), so each CTA splits roles:
//   * ONE consumer warp carries the dependent chain and nothing else.  A warp holds NS = 32/G sequences; lane
//     (sub, j) owns state j of sequence `sub` and keeps its column (or row) of the transition matrix in registers.
//     Per step it reads the previous K-vector back from a shared-memory ring with K/4 broadcast LDS.128, does K
//     FMAs (or K add + max), one multiply/add with the emission term and one STS.  No global access, no
//     transcendental, no integer division on the chain.
//   * helper warps stream emissions from HBM a chunk of CH frames ahead, convert them to what the consumer needs
//     (floor, max-normalise, exp) into a shared-memory ring, and trail behind the consumer draining its results to
//     HBM with coalesced stores; for Viterbi they also recompute the backpointers from the stored delta vectors
//     (same fp32 adds -> bit-identical), which takes the arg-max off the chain entirely.
//   * hand-off is two named barriers per ring buffer (FULL: helpers -> consumer, DONE: consumer -> helpers).
// Scaling: the forward/backward vectors are renormalised every step by a power of two taken from the exponent of the
// previous step's largest entry (one REDUX; exact; integer exponent bookkeeping), so alpha = a * 2^ksum * exp(sum m).
// A lone warp issues roughly one instruction every ~4-5 cycles on a dependent chain, so the consumer's instruction
// count IS its latency: packed fp32x2 FMAs/adds, 3-input max, no per-step branches, 4x unrolled (fits the L0 I-cache).
#include "recursion_smallk.cuh"

#include <stdlib.h>

namespace hmmb200 {

// ----------------------------------------------------------------------------------------------------------
// host-side dispatch
// ----------------------------------------------------------------------------------------------------------
template <int G, int KP>
static int launch_fb(FbParams p, cudaStream_t s) {
    constexpr int NS = 32 / G;
    p.bulk = bulk_feed_param(p.emis, p.T, p.K);
#ifdef HMMB200_DEBUG_HOOKS
    if (getenv("HMMB200_NO_BULK") || getenv("HMMB200_NO_BULK_FB")) p.bulk = 0;
#endif
    const size_t smem = FB_SMEM_BYTES + (p.bulk ? raw_stage_bytes(p.K, NS) : 0);
    cudaError_t e = cudaFuncSetAttribute(fb_sweep_kernel<G, KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_error(HMMB200_ELAUNCH, "fb smem opt-in: %s", cudaGetErrorString(e));
    dim3 grid((p.B + NS - 1) / NS, 2);
    fb_sweep_kernel<G, KP><<<grid, FB_THREADS, smem, s>>>(p);
    return check_launch("fb_sweep_kernel");
}

template <int G, int KP>
static int launch_vit(VitParams p, cudaStream_t s) {
    constexpr int NS = 32 / G;
    bool in_smem; size_t smem;
    p.bulk = bulk_feed_param(p.emis, p.T, p.K);
#ifdef HMMB200_DEBUG_HOOKS
    if (getenv("HMMB200_NO_BULK") || getenv("HMMB200_NO_BULK_VIT")) p.bulk = 0;
#endif
    const size_t raw = p.bulk ? raw_stage_bytes(p.K, NS) : 0;
    vit_plan(p.T, G, p.chunk, p.n_chunks, in_smem, smem, 200 * 1024 - raw);
    p.psi_in_smem = in_smem ? 1 : 0;
    if (smem > 200 * 1024 - raw) return set_error(HMMB200_EUNSUPPORTED, "viterbi: T=%d too long for the traceback tables", p.T);
    smem += raw;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(viterbi_kernel<G, KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return set_error(HMMB200_ELAUNCH, "viterbi smem opt-in: %s", cudaGetErrorString(e));
    }
    dim3 grid((p.B + NS - 1) / NS);
    viterbi_kernel<G, KP><<<grid, VIT_THREADS, smem, s>>>(p);
    return check_launch("viterbi_kernel");
}

static int dispatch_fb(const FbParams &p, cudaStream_t s) { HMMB200_DISPATCH_GK(launch_fb, p.K, p, s); }
static int dispatch_vit(const VitParams &p, cudaStream_t s) { HMMB200_DISPATCH_GK(launch_vit, p.K, p, s); }

}  // namespace hmmb200

using namespace hmmb200;

HMMB200_EXPORT size_t hmmb200_fb_workspace_bytes(int B, int T, int K) {
    if (B <= 0 || T <= 0 || K <= 0) return 0;
    if (K > 32) return largek_shape_ok(K) ? largek_fb_workspace_bytes(B, T, K) : 0;
    size_t n = (size_t)B * T;
    return 2 * align256(n * K * sizeof(float)) + 2 * align256(n * sizeof(float));
}

HMMB200_EXPORT int hmmb200_forward_backward_f32(const float *emis, int emis_mode, float floor_eps, int add_rowmax,
                                                const float *trans_prob, const float *init_prob, int B, int T, int K,
                                                float *gamma, float *fwd_prob, float *bwd_prob,
                                                float *log_alpha, float *log_beta, float *loglik,
                                                void *workspace, size_t workspace_bytes, void *stream) {
    if (B < 0 || T < 0 || K <= 0) return set_error(HMMB200_EINVAL, "forward_backward: bad shape B=%d T=%d K=%d", B, T, K);
    if (B == 0 || T == 0) return HMMB200_OK;
    if (K > 32 && !largek_shape_ok(K))
        return set_error(HMMB200_EUNSUPPORTED, "forward_backward: K <= 2048 states supported (got %d)", K);
    if (!emis || !trans_prob || !init_prob) return set_error(HMMB200_EINVAL, "forward_backward: null input");
    if (emis_mode < 0 || emis_mode > 3) return set_error(HMMB200_EINVAL, "forward_backward: bad emis_mode %d", emis_mode);
    if (!workspace || workspace_bytes < hmmb200_fb_workspace_bytes(B, T, K))
        return set_error(HMMB200_EWORKSPACE, "forward_backward: workspace %zu < %zu bytes", workspace_bytes,
                         hmmb200_fb_workspace_bytes(B, T, K));
    if (int rc = require_sm100()) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    if (K > 32)
        return largek_forward_backward(emis, emis_mode, floor_eps, add_rowmax, trans_prob, init_prob, B, T, K, gamma, fwd_prob,
                                       bwd_prob, log_alpha, log_beta, loglik, workspace, s);
    size_t n = (size_t)B * T;
    uint8_t *w = (uint8_t *)workspace;
    FbParams p;
    p.emis = emis; p.mode = emis_mode; p.eps = floor_eps; p.add_rowmax = add_rowmax;
    p.trans = trans_prob; p.init = init_prob; p.B = B; p.T = T; p.K = K;
    p.ws_a = (float *)w;  w += align256(n * K * sizeof(float));
    p.ws_b = (float *)w;  w += align256(n * K * sizeof(float));
    p.ws_la = (float *)w; w += align256(n * sizeof(float));
    p.ws_lb = (float *)w;
    p.loglik = loglik; p.pdl = 0;
    if (int rc = dispatch_fb(p, s)) return rc;
    if (gamma || fwd_prob || bwd_prob || log_alpha || log_beta) {
        CombineParams c;
        c.ws_a = p.ws_a; c.ws_b = p.ws_b; c.ws_la = p.ws_la; c.ws_lb = p.ws_lb;
        c.n_frames = (int64_t)n; c.K = K;
        c.gamma = gamma; c.fwd = fwd_prob; c.bwd = bwd_prob; c.log_alpha = log_alpha; c.log_beta = log_beta; c.bf16 = 0;
        if (int rc = launch_combine(c, s)) return rc;
    }
    return HMMB200_OK;
}

// K <= 32: do the decoded path and the traceback tables of a group of sequences fit one CTA's shared memory?  (NS T bytes of path:
// T up to ~24 k at K <= 4, ~48 k at K <= 8, ~95 k at K <= 16, ~190 k above.)  Longer sequences take the cluster kernel's sweep
// (one CTA per group, trellis in HBM) and its path-only traceback: same arithmetic, same results, no length limit.
static bool vit_smallk_fits(int T, int K) {
    int G = group_lanes(K), L, nC; bool in_smem; size_t smem;
    const size_t budget = 200 * 1024 - raw_stage_bytes(K, 32 / G);
    vit_plan(T, G, L, nC, in_smem, smem, budget);
    return smem <= budget;
}

HMMB200_EXPORT size_t hmmb200_viterbi_workspace_bytes(int B, int T, int K) {
    if (B <= 0 || T <= 0 || K <= 0) return 0;
    if (K > 32) return largek_shape_ok(K) ? largek_viterbi_workspace_bytes(B, T, K) : 0;
    if (!vit_smallk_fits(T, K)) return largek_viterbi_workspace_bytes(B, T, K);
    int G = group_lanes(K), L, nC; bool in_smem; size_t smem;
    // (the launch may also carve the raw emission stage out of the same budget: plan with it, so that the query never under-reports)
    vit_plan(T, G, L, nC, in_smem, smem, 200 * 1024 - raw_stage_bytes(K, 32 / G));
    return in_smem ? 0 : (size_t)B * T * G;
}

HMMB200_EXPORT int hmmb200_viterbi_f32(const float *emis, int emis_mode, float floor_eps,
                                       const float *log_trans, const float *log_init, int B, int T, int K,
                                       float *delta, void *psi_out, int64_t *states, float *score,
                                       void *workspace, size_t workspace_bytes, void *stream) {
    uint8_t *psi = (uint8_t *)psi_out;
    if (B < 0 || T < 0 || K <= 0) return set_error(HMMB200_EINVAL, "viterbi: bad shape B=%d T=%d K=%d", B, T, K);
    if (B == 0 || T == 0) return HMMB200_OK;
    if (K > 32 && !largek_shape_ok(K)) return set_error(HMMB200_EUNSUPPORTED, "viterbi: K <= 2048 states supported (got %d)", K);
    if (!emis || !log_trans || !log_init || !states) return set_error(HMMB200_EINVAL, "viterbi: null argument");
    if (emis_mode < 0 || emis_mode > 3) return set_error(HMMB200_EINVAL, "viterbi: bad emis_mode %d", emis_mode);
    size_t need = hmmb200_viterbi_workspace_bytes(B, T, K);
    if (need && (!workspace || workspace_bytes < need))
        return set_error(HMMB200_EWORKSPACE, "viterbi: workspace %zu < %zu bytes", workspace_bytes, need);
    if (int rc = require_sm100()) return rc;
    if (K > 32 || !vit_smallk_fits(T, K))
        return largek_viterbi(emis, emis_mode, floor_eps, log_trans, log_init, B, T, K, delta, psi, states, score, workspace,
                              (cudaStream_t)stream);
    VitParams p;
    p.emis = emis; p.mode = emis_mode; p.eps = floor_eps; p.log_trans = log_trans; p.log_init = log_init;
    p.B = B; p.T = T; p.K = K; p.delta = delta; p.psi_out = psi; p.states = states; p.score = score;
    p.psi_ws = (uint8_t *)workspace; p.psi_in_smem = 1; p.chunk = 64; p.n_chunks = 0; p.pdl = 0;
    return dispatch_vit(p, (cudaStream_t)stream);
}
