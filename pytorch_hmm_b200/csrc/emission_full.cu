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
__device__ __forceinline__ float2 ef_ffma2(float2 a, float2 b, float2 c) {
    unsigned long long ra = *reinterpret_cast<unsigned long long *>(&a), rb = *reinterpret_cast<unsigned long long *>(&b);
    unsigned long long rc = *reinterpret_cast<unsigned long long *>(&c), rd;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2 *>(&rd);
}

// The two running sums of a row (even and odd columns) are the halves of ONE packed accumulator: a 16-byte load of the row feeds two
// fma.rn.f32x2 instead of four FFMA -- same products, same order, same bits, 3 instead of 5 instructions per four FMAs (the kernel was
// issue-bound: 71 % of its instructions FFMA at 54 % issue utilisation).
// NF frames per thread: every 16-byte load of a row of W (a broadcast: the same address in all lanes, but 512 bytes on the return path
// per warp) feeds 4 NF FMAs -- with one frame per thread the kernel sat on the shared-memory return path at a third of the FMA rate.
template <int DT, int BAND, int NF>
__device__ __forceinline__ void ef_band_rows(const float *Wb, int D, const float2 (&xr)[NF][DT / 2], float (&q)[NF]) {
    constexpr int NB4 = (4 * (BAND + 1) < DT / 4) ? 4 * (BAND + 1) : DT / 4;
    const int r_hi = min(D, 16 * BAND + 16);
    for (int r = 16 * BAND; r < r_hi; ++r) {
        const float4 *wr = reinterpret_cast<const float4 *>(Wb + r * DT);
        const float c0 = Wb[D * DT + r];
        float2 y2[NF];
#pragma unroll
        for (int f = 0; f < NF; ++f) y2[f] = make_float2(c0, 0.f);
#pragma unroll
        for (int d4 = 0; d4 < NB4; ++d4) {
            const float4 w = wr[d4];
#pragma unroll
            for (int f = 0; f < NF; ++f) {
                y2[f] = ef_ffma2(make_float2(w.x, w.y), xr[f][2 * d4], y2[f]);
                y2[f] = ef_ffma2(make_float2(w.z, w.w), xr[f][2 * d4 + 1], y2[f]);
            }
        }
#pragma unroll
        for (int f = 0; f < NF; ++f) { const float y = y2[f].x + y2[f].y; q[f] = fmaf(y, y, q[f]); }
    }
    if constexpr (16 * (BAND + 1) < DT) ef_band_rows<DT, BAND + 1, NF>(Wb, D, xr, q);
}

template <int DT, int NF>    // DT = D padded to a multiple of 4 (compile-time: the x rows and the inner loops live in registers); NF frames per thread
__global__ void __launch_bounds__(EF_THREADS) gmm_emission_full_kernel(FullParams p) {
    extern __shared__ __align__(16) float ef_smem[];
    const int D = p.D, DP = DT, KC = p.K * p.C, C = p.C;
    const int wfloats = D * DP + DP;                       // one component: W rows + c (padded)
    float *const wbuf0 = ef_smem, *const wbuf1 = ef_smem + wfloats;
    float *lrow = ef_smem + 2 * wfloats;                   // [NF][EF_THREADS][C] per-frame component values of the current state
    const int tid = threadIdx.x;
    auto stage = [&](int kc, int b) {                      // W_kc (D x DP) and c_kc into buffer b
        const float4 *src = reinterpret_cast<const float4 *>(p.W + (size_t)kc * D * DP);
        float *wb = b ? wbuf1 : wbuf0;
        float4 *dst = reinterpret_cast<float4 *>(wb);
        for (int i = tid; i < D * DP / 4; i += EF_THREADS) ef_cp16(dst + i, src + i);
        for (int i = tid; i < DP; i += EF_THREADS) wb[D * DP + i] = (i < D) ? p.cvec[(size_t)kc * D + i] : 0.f;
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    constexpr int TILE = EF_THREADS * NF;
    const int64_t n_tiles = (p.n + TILE - 1) / TILE;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        int64_t fr[NF];
        bool live[NF];
        float2 xr[NF][DT / 2];
#pragma unroll
        for (int f = 0; f < NF; ++f) {
            fr[f] = tile * TILE + f * EF_THREADS + tid;    // (consecutive threads take consecutive frames)
            live[f] = fr[f] < p.n;
#pragma unroll
            for (int d = 0; d < DT; d += 4) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (live[f] && d < D) {
                    const float *xp = p.x + fr[f] * D + d;
                    if ((D & 3) == 0 && ((((uintptr_t)p.x) & 15) == 0)) v = *reinterpret_cast<const float4 *>(xp);
                    else { v.x = xp[0]; if (d + 1 < D) v.y = xp[1]; if (d + 2 < D) v.z = xp[2]; if (d + 3 < D) v.w = xp[3]; }
                }
                xr[f][d / 2] = make_float2(v.x, v.y); xr[f][d / 2 + 1] = make_float2(v.z, v.w);
            }
        }
        __syncthreads();                                   // previous tile done with the W buffers
        stage(0, 0);
        for (int kc = 0; kc < KC; ++kc) {
            const int b = kc & 1;
            if (kc + 1 < KC) stage(kc + 1, b ^ 1);
            if (kc + 1 < KC) asm volatile("cp.async.wait_group 1;" ::: "memory");
            else asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncthreads();
            const float *Wb = b ? wbuf1 : wbuf0;
            float q[NF];
#pragma unroll
            for (int f = 0; f < NF; ++f) q[f] = 0.f;
            ef_band_rows<DT, 0, NF>(Wb, D, xr, q);
            const float cstv = __ldg(p.cst + kc);
#pragma unroll
            for (int f = 0; f < NF; ++f) {
                float *mine = lrow + ((size_t)f * EF_THREADS + tid) * C;
                const float l = fmaf(-0.5f, q[f], cstv);
                if (p.comp && live[f]) p.comp[fr[f] * KC + kc] = l;
                mine[kc % C] = l;
                if (kc % C == C - 1) {                     // the state's C components are in: the reference's private log-sum-exp
                    float out;
                    if (C == 1) out = l;
                    else {
                        float m = mine[0];
                        for (int c = 1; c < C; ++c) m = fmaxf(m, mine[c]);
                        if (isinf(m)) m = 0.f;
                        float sm = 0.f;
                        for (int c = 0; c < C; ++c) sm += expf(mine[c] - m);
                        out = logf(fmaxf(sm, 1e-8f)) + m;
                    }
                    if (live[f]) p.logb[fr[f] * p.K + kc / C] = out;
                }
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
    // two frames per thread when their x rows fit the register file (D <= 80: 2 x 80 of the 255 registers)
    const int NF = (p.DP <= 80) ? 2 : 1;
    const int64_t n_tiles = (n_frames + EF_THREADS * NF - 1) / (EF_THREADS * NF);
    cudaStream_t s = (cudaStream_t)stream;
    const size_t smem = (2 * ((size_t)D * p.DP + p.DP) + (size_t)NF * EF_THREADS * C) * sizeof(float);
    if (smem > 200 * 1024) return set_error(HMMB200_EUNSUPPORTED, "gmm_emission_full: D = %d, C = %d need %zu bytes of shared memory", D, C, smem);
    // persistent CTAs, as many per SM as fit (the kernel waits on shared-memory loads between its FFMA runs: with two 4-warp CTAs per
    // SM the schedulers issued 54 % of the time)
    int grid = 0;
#define EF_LAUNCH(N)                                                                                                            \
    case N: {                                                                                                                   \
        auto kern = (N <= 80) ? gmm_emission_full_kernel<N, (N <= 80 ? 2 : 1)> : gmm_emission_full_kernel<N, 1>;                \
        if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);               \
        int occ = 2;                                                                                                            \
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, EF_THREADS, smem) != cudaSuccess || occ < 1) { cudaGetLastError(); occ = 2; } \
        grid = (int)min((int64_t)sms * occ, n_tiles);                                                                           \
        kern<<<grid, EF_THREADS, smem, s>>>(p);                                                                                 \
        break;                                                                                                                  \
    }
    switch (p.DP) {
        EF_LAUNCH(4) EF_LAUNCH(8) EF_LAUNCH(12) EF_LAUNCH(16) EF_LAUNCH(20) EF_LAUNCH(24) EF_LAUNCH(28) EF_LAUNCH(32)
        EF_LAUNCH(40) EF_LAUNCH(48) EF_LAUNCH(64) EF_LAUNCH(80) EF_LAUNCH(96)
        default: return set_error(HMMB200_EUNSUPPORTED, "gmm_emission_full: padded feature_dim %d has no kernel instance (use D in {<=32 step 4, 40, 48, 64, 80, 96})", p.DP);
    }
#undef EF_LAUNCH
    return check_launch("gmm_emission_full_kernel");
}
