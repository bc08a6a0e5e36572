// emission_full.cu -- full-covariance GMM emission log-likelihoods (sm_100a).  SURVEY 8(f) rank 3.
//
//   replaces  MixtureGaussianHMMLayer._full_gaussian_log_probs   pytorch_hmm/mixture_gaussian.py:216-240
// The reference solves L_kc y = (x - mu_kc) for every (frame, state, component) with torch.linalg.solve_triangular on a
// [B,T,S,C,D,1] broadcast.  Here the triangular solve is folded into the parameters once per update (host side, O(K C D^3)):
//   y = W_kc x + c_kc,   W_kc = L_kc^-1 (lower triangular),   c_kc = -W_kc mu_kc
//   l_kc(x) = log w_kc - 0.5 (|y|^2 + log det_kc + D log 2 pi),   log b_k = own_lse_c l_kc        (mixture_gaussian.py:141-155)
// so a frame costs K*C triangular matrix-vector products: a second contraction, [frames, D] x [D, K*C*D], followed by a
// sum of squares per component.  One thread owns one frame (its x row lives in registers); the rows of W_kc are broadcast to the
// whole CTA from shared memory as 16-byte loads (4 FMAs per load); components are double-buffered through cp.async.
// CUDA-core fp32 (FFMA): 0.5 K C D^2 FMAs per frame -- 154 k at K=12, C=4, D=80.
#include "common.cuh"

namespace hmmb200 {

constexpr int EF_THREADS = 128;            // frames per CTA tile (one per thread)

struct FullParams {
    const float *x;        // [n, D]
    const float *W;        // [KC, D, DP] lower-triangular inverse Cholesky factors, rows padded to DP = multiple of 4, zeros above the diagonal
    const float *cvec;     // [KC, D]   -W mu
    const float *cst;      // [KC]      log w - 0.5 (log det + D log 2 pi)
    int64_t n;
    int K, C, D, DP;
    float *logb;           // [n, K]
    float *comp;           // [n, KC] or null
};

__device__ __forceinline__ void ef_cp16(void *dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}

// rows [16 BAND, 16 BAND + 16) of the lower-triangular factor only reach columns < 16 (BAND + 1): the column loop of a band is a
// compile-time bound, so the x row stays in registers and at most 15 of every 16 (BAND + 1) products are above the diagonal
template <int DT, int BAND>
__device__ __forceinline__ void ef_band_rows(const float *Wb, int D, const float (&xr)[DT], float &q) {
    constexpr int NB4 = (4 * (BAND + 1) < DT / 4) ? 4 * (BAND + 1) : DT / 4;
    const int r_hi = min(D, 16 * BAND + 16);
    for (int r = 16 * BAND; r < r_hi; ++r) {
        const float4 *wr = reinterpret_cast<const float4 *>(Wb + r * DT);
        float y0 = Wb[D * DT + r], y1 = 0.f;
#pragma unroll
        for (int d4 = 0; d4 < NB4; ++d4) {
            const float4 w = wr[d4];
            y0 = fmaf(w.x, xr[4 * d4], y0); y1 = fmaf(w.y, xr[4 * d4 + 1], y1);
            y0 = fmaf(w.z, xr[4 * d4 + 2], y0); y1 = fmaf(w.w, xr[4 * d4 + 3], y1);
        }
        const float y = y0 + y1;
        q = fmaf(y, y, q);
    }
    if constexpr (16 * (BAND + 1) < DT) ef_band_rows<DT, BAND + 1>(Wb, D, xr, q);
}

template <int DT>            // DT = D padded to a multiple of 4 (compile-time: the x row and the inner loops live in registers)
__global__ void __launch_bounds__(EF_THREADS) gmm_emission_full_kernel(FullParams p) {
    extern __shared__ __align__(16) float ef_smem[];
    const int D = p.D, DP = DT, KC = p.K * p.C, C = p.C;
    const int wfloats = D * DP + DP;                       // one component: W rows + c (padded)
    float *const wbuf0 = ef_smem, *const wbuf1 = ef_smem + wfloats;
    float *lrow = ef_smem + 2 * wfloats;                   // [EF_THREADS][C] per-thread component values of the current state
    const int tid = threadIdx.x;
    auto stage = [&](int kc, int b) {                      // W_kc (D x DP) and c_kc into buffer b
        const float4 *src = reinterpret_cast<const float4 *>(p.W + (size_t)kc * D * DP);
        float *wb = b ? wbuf1 : wbuf0;
        float4 *dst = reinterpret_cast<float4 *>(wb);
        for (int i = tid; i < D * DP / 4; i += EF_THREADS) ef_cp16(dst + i, src + i);
        for (int i = tid; i < DP; i += EF_THREADS) wb[D * DP + i] = (i < D) ? p.cvec[(size_t)kc * D + i] : 0.f;
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    const int64_t n_tiles = (p.n + EF_THREADS - 1) / EF_THREADS;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t fr = tile * EF_THREADS + tid;
        const bool live = fr < p.n;
        float xr[DT];
#pragma unroll
        for (int d = 0; d < DT; d += 4) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (live && d < D) {
                if ((D & 3) == 0 && ((((uintptr_t)p.x) & 15) == 0)) v = *reinterpret_cast<const float4 *>(p.x + fr * D + d);
                else { v.x = p.x[fr * D + d]; if (d + 1 < D) v.y = p.x[fr * D + d + 1]; if (d + 2 < D) v.z = p.x[fr * D + d + 2]; if (d + 3 < D) v.w = p.x[fr * D + d + 3]; }
            }
            xr[d] = v.x; xr[d + 1] = v.y; xr[d + 2] = v.z; xr[d + 3] = v.w;
        }
        __syncthreads();                                   // previous tile done with the W buffers
        stage(0, 0);
        float *mine = lrow + tid * C;
        for (int kc = 0; kc < KC; ++kc) {
            const int b = kc & 1;
            if (kc + 1 < KC) stage(kc + 1, b ^ 1);
            if (kc + 1 < KC) asm volatile("cp.async.wait_group 1;" ::: "memory");
            else asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncthreads();
            const float *Wb = b ? wbuf1 : wbuf0;
            float q = 0.f;
            ef_band_rows<DT, 0>(Wb, D, xr, q);
            const float l = fmaf(-0.5f, q, __ldg(p.cst + kc));
            if (p.comp && live) p.comp[fr * KC + kc] = l;
            mine[kc % C] = l;
            if (kc % C == C - 1) {                         // the state's C components are in: the reference's private log-sum-exp
                float out;
                if (C == 1) out = l;
                else {
                    float m = mine[0];
                    for (int c = 1; c < C; ++c) m = fmaxf(m, mine[c]);
                    if (isinf(m)) m = 0.f;
                    float s = 0.f;
                    for (int c = 0; c < C; ++c) s += expf(mine[c] - m);
                    out = logf(fmaxf(s, 1e-8f)) + m;
                }
                if (live) p.logb[fr * p.K + kc / C] = out;
            }
            __syncthreads();                               // buffer b may be refilled by the next-but-one component
        }
    }
}

}  // namespace hmmb200

using namespace hmmb200;

// W [K*C, D, DP] with DP = (D + 3) & ~3 (rows zero-padded; zeros above the diagonal), cvec [K*C, D], cst [K*C]: see the file header.
HMMB200_EXPORT int hmmb200_gmm_emission_full_f32(const float *x, const float *W, const float *cvec, const float *cst, int64_t n_frames,
                                                 int K, int C, int D, float *logb, float *comp, void *stream) {
    if (n_frames < 0 || K <= 0 || C <= 0 || D <= 0) return set_error(HMMB200_EINVAL, "gmm_emission_full: bad shape");
    if (n_frames == 0) return HMMB200_OK;
    if (!x || !W || !cvec || !cst || !logb) return set_error(HMMB200_EINVAL, "gmm_emission_full: null argument");
    if (D > 96) return set_error(HMMB200_EUNSUPPORTED, "gmm_emission_full: feature_dim <= 96 (got %d)", D);
    if ((((uintptr_t)W) & 15) != 0) return set_error(HMMB200_EINVAL, "gmm_emission_full: W must be 16-byte aligned");
    if (int rc = require_sm100()) return rc;
    FullParams p;
    p.x = x; p.W = W; p.cvec = cvec; p.cst = cst; p.n = n_frames; p.K = K; p.C = C; p.D = D; p.DP = (D + 3) & ~3;
    p.logb = logb; p.comp = comp;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t n_tiles = (n_frames + EF_THREADS - 1) / EF_THREADS;
    cudaStream_t s = (cudaStream_t)stream;
    const size_t smem = (2 * ((size_t)D * p.DP + p.DP) + (size_t)EF_THREADS * C) * sizeof(float);
    if (smem > 200 * 1024) return set_error(HMMB200_EUNSUPPORTED, "gmm_emission_full: D = %d, C = %d need %zu bytes of shared memory", D, C, smem);
    const int grid = (int)min((int64_t)sms * 2, n_tiles);
#define EF_LAUNCH(N)                                                                                                            \
    case N:                                                                                                                     \
        if (smem > 48 * 1024) cudaFuncSetAttribute(gmm_emission_full_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        gmm_emission_full_kernel<N><<<grid, EF_THREADS, smem, s>>>(p);                                                          \
        break;
    switch (p.DP) {
        EF_LAUNCH(4) EF_LAUNCH(8) EF_LAUNCH(12) EF_LAUNCH(16) EF_LAUNCH(20) EF_LAUNCH(24) EF_LAUNCH(28) EF_LAUNCH(32)
        EF_LAUNCH(40) EF_LAUNCH(48) EF_LAUNCH(64) EF_LAUNCH(80) EF_LAUNCH(96)
        default: return set_error(HMMB200_EUNSUPPORTED, "gmm_emission_full: padded feature_dim %d has no kernel instance (use D in {<=32 step 4, 40, 48, 64, 80, 96})", p.DP);
    }
#undef EF_LAUNCH
    return check_launch("gmm_emission_full_kernel");
}
