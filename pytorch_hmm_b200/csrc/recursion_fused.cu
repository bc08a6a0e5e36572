// recursion_fused.cu -- forward sweep + backward sweep + Viterbi (with its on-device traceback) of a group of sequences in ONE CTA.
//
// This is the whole-batch pass the reference times (examples/benchmark.py:120-196: forward_backward, then viterbi_decode, on the
// same emissions) as one launch.  The three recursions are latency chains of T dependent steps each carried by a single
// "consumer" warp (recursion_smallk.cuh).  What a chain loses inside a busy SM is, in this order (ncu source-page samples on the
// consumers' loops): INSTRUCTION FETCH -- with tens of KB of unrolled helper code streaming through the SM the consumers spent a
// third of their samples in stall_no_inst, which is why the helper roles are written as small rolled loops -- and then issue slots
// taken by other warps on the consumer's scheduler.  The kernel owns the placement (warp w issues from sub-partition w % 4):
//     SMSP 1: forward consumer (warp 1), its two loaders, one of its drainers (+ one backward drainer)
//     SMSP 2: backward consumer (warp 2), its two loaders, one of its drainers
//     SMSP 3: Viterbi consumer (warp 3), its two loaders (+ one forward drainer)
//     SMSP 0: the five Viterbi drainers (the backpointer recomputation is the heaviest helper role)
// i.e. a consumer shares its scheduler -- and its 6 KB L0 instruction cache -- only with the small loops of its own pipeline.
// Measured on the headline shape (tools/diag_layout.py, 7 placements): 0.126 ms, against 0.132 - 0.152 ms for the others.
// Each pipeline keeps its own named barriers and shared-memory rings; the Viterbi traceback synchronises only the Viterbi warps, so
// it runs while the sweeps finish.  Launched with programmatic stream serialisation the set-up (transition columns into registers,
// ring carve-up) overlaps the tail of the emission kernel that produces the log-emissions.
#include "recursion_smallk.cuh"

#include <stdlib.h>

namespace hmmb200 {

constexpr int FU_WARPS = 3 + 2 * (FB_NL + FB_ND) + VIT_NL + VIT_ND;     // 18
constexpr int FU_THREADS = 32 * FU_WARPS;
static_assert(FU_WARPS == 18 && FB_NL == 2 && FB_ND == 2 && VIT_NL == 2 && VIT_ND == 5, "the warp map below is written for 18 warps");

// warp -> (pipeline: 0 forward, 1 backward, 2 Viterbi; role within the pipeline: 0 consumer, then loaders, then drainers)
template <int LAYOUT>
__device__ __forceinline__ void fused_role(int warp, int &pipe, int &role) {
    // LAYOUT 0: each consumer beside its own pipeline's helpers (SMSPs 1, 2, 3), the Viterbi drainers on SMSP 0
    // LAYOUT 1: consumers on the three lowest warp ids (SMSPs 0, 1, 2), Viterbi drainers 0-3 on SMSP 3
    //                          w0 w1 w2 w3 w4 w5 w6 w7 w8 w9 10 11 12 13 14 15 16 17
    constexpr int PIPE0[18] = {2, 0, 1, 2, 2, 0, 1, 2, 2, 0, 1, 2, 2, 0, 1, 0, 2, 1};
    constexpr int ROLE0[18] = {3, 0, 0, 0, 4, 1, 1, 1, 5, 2, 2, 2, 6, 3, 3, 4, 7, 4};
    constexpr int PIPE1[18] = {0, 1, 2, 2, 1, 0, 2, 2, 0, 2, 1, 2, 1, 0, 0, 2, 1, 2};
    constexpr int ROLE1[18] = {0, 0, 0, 3, 1, 1, 1, 4, 3, 2, 3, 5, 2, 2, 4, 6, 4, 7};
    unsigned long long pm = 0, rm = 0;
#pragma unroll
    for (int i = 0; i < 18; ++i) {
        pm |= (unsigned long long)(LAYOUT == 0 ? PIPE0[i] : PIPE1[i]) << (2 * i);
        rm |= (unsigned long long)(LAYOUT == 0 ? ROLE0[i] : ROLE1[i]) << (3 * i);
    }
    pipe = (int)((pm >> (2 * warp)) & 3);
    role = (int)((rm >> (3 * warp)) & 7);
}

struct FusedParams {
    FbParams fb;
    VitParams vit;
    int dbg;              // debug builds only: bit 0 warp layout, bits 4-6 skip the forward / backward / Viterbi pipeline (timing experiments)
    unsigned long long dbg_pm, dbg_rm;   // debug builds only: a warp map supplied at run time (2 / 3 bits per warp), 0 = none
};

// shared-memory layout: [forward sweep][backward sweep][raw stage asc (forward)][raw stage desc (backward)][raw stage asc (Viterbi)][Viterbi]
template <bool PAD> __host__ __device__ constexpr size_t fu_smem_f() { return FbSmem<0, PAD>::BYTES; }
template <bool PAD> __host__ __device__ constexpr size_t fu_smem_b() { return FbSmem<1, PAD>::BYTES; }

template <int G, int KP>
__global__ void __launch_bounds__(FU_THREADS, 1) fb_viterbi_kernel(const __grid_constant__ FusedParams p) {
    extern __shared__ __align__(16) uint8_t smem[];
    constexpr bool PAD = KP < G;
    int pipe, role;
#ifdef HMMB200_DEBUG_HOOKS
    if (p.dbg_rm != 0ull) {
        pipe = (int)((p.dbg_pm >> (2 * (threadIdx.x >> 5))) & 3);
        role = (int)((p.dbg_rm >> (3 * (threadIdx.x >> 5))) & 7);
    } else if (p.dbg & 1) fused_role<1>(threadIdx.x >> 5, pipe, role);
    else fused_role<0>(threadIdx.x >> 5, pipe, role);
#else
    fused_role<0>(threadIdx.x >> 5, pipe, role);
#endif
    constexpr size_t OFF_B = fu_smem_f<PAD>(), OFF_RAW = OFF_B + fu_smem_b<PAD>();
    const size_t raw = p.fb.bulk ? raw_stage_bytes(p.fb.K, 32 / G) : 0;
    const RawStage rs_f(smem + OFF_RAW, p.fb.K, p.fb.bulk), rs_b(smem + OFF_RAW + raw, p.fb.K, p.fb.bulk), rs_v(smem + OFF_RAW + 2 * raw, p.fb.K, p.fb.bulk);
    if (p.fb.bulk) {
        if (threadIdx.x == 0) { rs_f.init(FB_NL); rs_b.init(FB_NL); rs_v.init(VIT_NL); }
        __syncthreads();
    }
#ifdef HMMB200_DEBUG_HOOKS
    if ((p.dbg >> 4) & (1 << pipe)) return;                         // timing experiments: leave a pipeline out
#endif
    if (pipe == 0) {
        fb_roles<G, KP, 0, PAD>(p.fb, smem, role, PipeBars{1, 3, FB_THREADS}, rs_f);
    } else if (pipe == 1) {
        fb_roles<G, KP, 1, PAD>(p.fb, smem + OFF_B, role, PipeBars{5, 7, FB_THREADS}, rs_b);
    } else {
        vit_roles<G, KP>(p.vit, smem + OFF_RAW + 3 * raw, role, role * 32 + (int)(threadIdx.x & 31), PipeBars{9, 11, VIT_THREADS}, rs_v,
                         [] { bar_sync(13, VIT_THREADS); });
    }
}

template <int G, int KP>
static int launch_fused(FusedParams p, int pdl, cudaStream_t s) {
    constexpr int NS = 32 / G;
    bool in_smem; size_t vsmem;
    constexpr bool PAD = KP < G;
    p.fb.bulk = p.vit.bulk = bulk_feed_param(p.fb.emis, p.fb.T, p.fb.K);
    const size_t fixed = fu_smem_f<PAD>() + fu_smem_b<PAD>() + (p.fb.bulk ? 3 * raw_stage_bytes(p.fb.K, NS) : 0);
    const size_t budget = 227 * 1024 - fixed;
    vit_plan(p.vit.T, G, p.vit.chunk, p.vit.n_chunks, in_smem, vsmem, budget);
    if (!in_smem || vsmem > budget) return 1;                       // backpointers do not fit beside the sweeps: caller runs the separate kernels
    p.vit.psi_in_smem = 1;
    p.fb.pdl = p.vit.pdl = pdl;
    p.dbg = 0; p.dbg_pm = p.dbg_rm = 0ull;
#ifdef HMMB200_DEBUG_HOOKS
    if (const char *e = getenv("HMMB200_FUSED_MAP")) {               // "pm,rm" as decimal integers
        char *end = nullptr;
        p.dbg_pm = strtoull(e, &end, 10);
        if (end && *end == ',') p.dbg_rm = strtoull(end + 1, nullptr, 10);
    }
    if (const char *e = getenv("HMMB200_FUSED_DBG")) p.dbg = atoi(e);
    if (p.dbg & 2) p.fb.bulk = p.vit.bulk = 0;                     // bit 1: per-lane global loads instead of the bulk-copy feed (A/B timing)
#endif
    const size_t smem = fixed + vsmem;
    static bool done[64];                                           // per-device function attribute, set once (idempotent)
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !done[dev]) {
        cudaError_t e = cudaFuncSetAttribute(fb_viterbi_kernel<G, KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return set_error(HMMB200_ELAUNCH, "fb_viterbi smem opt-in: %s", cudaGetErrorString(e));
        if (dev >= 0 && dev < 64) done[dev] = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((p.fb.B + NS - 1) / NS);
    cfg.blockDim = dim3(FU_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    cudaError_t e = cudaLaunchKernelEx(&cfg, fb_viterbi_kernel<G, KP>, p);
    if (e != cudaSuccess) return set_error(HMMB200_ELAUNCH, "fb_viterbi_kernel: %s", cudaGetErrorString(e));
    return check_launch("fb_viterbi_kernel");
}

static int dispatch_fused(const FusedParams &p, int pdl, cudaStream_t s) { HMMB200_DISPATCH_GK(launch_fused, p.fb.K, p, pdl, s); }

}  // namespace hmmb200

using namespace hmmb200;

HMMB200_EXPORT size_t hmmb200_fb_viterbi_workspace_bytes(int B, int T, int K) {
    const size_t a = hmmb200_fb_workspace_bytes(B, T, K);
    if (a == 0) return 0;
    return align256(a) + hmmb200_viterbi_workspace_bytes(B, T, K);
}

HMMB200_EXPORT int hmmb200_fb_viterbi_f32(const float *emis, int fb_mode, int vit_mode, float floor_eps, int add_rowmax,
                                          const float *trans_prob, const float *init_prob,
                                          const float *log_trans, const float *log_init, int B, int T, int K,
                                          float *gamma, float *fwd_prob, float *bwd_prob, float *log_alpha, float *log_beta,
                                          float *loglik, float *delta, void *psi, int64_t *states, float *score,
                                          void *workspace, size_t workspace_bytes, int flags, void *stream) {
    if (B < 0 || T < 0 || K <= 0) return set_error(HMMB200_EINVAL, "fb_viterbi: bad shape B=%d T=%d K=%d", B, T, K);
    if (B == 0 || T == 0) return HMMB200_OK;
    if (!emis || !trans_prob || !init_prob || !log_trans || !log_init || !states) return set_error(HMMB200_EINVAL, "fb_viterbi: null argument");
    if (fb_mode < 0 || fb_mode > 3 || vit_mode < 0 || vit_mode > 3) return set_error(HMMB200_EINVAL, "fb_viterbi: bad emission mode");
    if (flags & HMMB200_FUSED_BF16_OUT) {
        if (K > 32) return set_error(HMMB200_EUNSUPPORTED, "fb_viterbi: bfloat16 outputs are available for K <= 32");
        if (log_alpha || log_beta) return set_error(HMMB200_EINVAL, "fb_viterbi: log_alpha / log_beta have no bfloat16 form");
    }
    const size_t fb_bytes = hmmb200_fb_workspace_bytes(B, T, K);
    const size_t need = hmmb200_fb_viterbi_workspace_bytes(B, T, K);
    if (fb_bytes == 0) return set_error(HMMB200_EUNSUPPORTED, "fb_viterbi: K <= 2048 states supported (got %d)", K);
    if (!workspace || workspace_bytes < need) return set_error(HMMB200_EWORKSPACE, "fb_viterbi: workspace %zu < %zu bytes", workspace_bytes, need);
    if (int rc = require_sm100()) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    uint8_t *w = (uint8_t *)workspace;
    if (K <= 32) {
        const size_t n = (size_t)B * T;
        FusedParams p;
        p.fb.emis = emis; p.fb.mode = fb_mode; p.fb.eps = floor_eps; p.fb.add_rowmax = add_rowmax;
        p.fb.trans = trans_prob; p.fb.init = init_prob; p.fb.B = B; p.fb.T = T; p.fb.K = K;
        uint8_t *q = w;
        p.fb.ws_a = (float *)q;  q += align256(n * K * sizeof(float));
        p.fb.ws_b = (float *)q;  q += align256(n * K * sizeof(float));
        p.fb.ws_la = (float *)q; q += align256(n * sizeof(float));
        p.fb.ws_lb = (float *)q;
        p.fb.loglik = loglik; p.fb.pdl = 0;
        p.vit.emis = emis; p.vit.mode = vit_mode; p.vit.eps = floor_eps; p.vit.log_trans = log_trans; p.vit.log_init = log_init;
        p.vit.B = B; p.vit.T = T; p.vit.K = K; p.vit.delta = delta; p.vit.psi_out = (uint8_t *)psi; p.vit.states = states;
        p.vit.score = score; p.vit.psi_ws = nullptr; p.vit.psi_in_smem = 1; p.vit.chunk = 64; p.vit.n_chunks = 0; p.vit.pdl = 0;
        const int rc = dispatch_fused(p, (flags & HMMB200_FUSED_PDL) ? 1 : 0, s);
        if (rc < 0) return rc;
        if (rc == 0) {
            if (gamma || fwd_prob || bwd_prob || log_alpha || log_beta) {
                CombineParams c;
                c.ws_a = p.fb.ws_a; c.ws_b = p.fb.ws_b; c.ws_la = p.fb.ws_la; c.ws_lb = p.fb.ws_lb;
                c.n_frames = (int64_t)n; c.K = K;
                c.gamma = gamma; c.fwd = fwd_prob; c.bwd = bwd_prob; c.log_alpha = log_alpha; c.log_beta = log_beta;
                c.bf16 = (flags & HMMB200_FUSED_BF16_OUT) ? 1 : 0;
                return launch_combine(c, s);
            }
            return HMMB200_OK;
        }
    }
    if (K > 32) {
        if (!largek_shape_ok(K)) return set_error(HMMB200_EUNSUPPORTED, "fb_viterbi: K <= 2048 states supported (got %d)", K);
        return largek_fb_viterbi(emis, fb_mode, vit_mode, floor_eps, add_rowmax, trans_prob, init_prob, log_trans, log_init, B, T, K,
                                 gamma, fwd_prob, bwd_prob, log_alpha, log_beta, loglik, delta, psi, states, score,
                                 w, w + align256(fb_bytes), s);
    }
    // shapes the fused kernel does not take (backpointers that do not fit beside the sweeps): the two stand-alone passes
    if (int rc = hmmb200_forward_backward_f32(emis, fb_mode, floor_eps, add_rowmax, trans_prob, init_prob, B, T, K, gamma, fwd_prob,
                                              bwd_prob, log_alpha, log_beta, loglik, w, fb_bytes, stream)) return rc;
    return hmmb200_viterbi_f32(emis, vit_mode, floor_eps, log_trans, log_init, B, T, K, delta, psi, states, score,
                               w + align256(fb_bytes), workspace_bytes - align256(fb_bytes), stream);
}
