// bw.cu -- Baum-Welch E-step sufficient statistics for a GMM-HMM (sm_100a).
//
// The reference has no Baum-Welch code; the formulas are docs/01_hmm_theory.md:196-227 (gamma :204, xi :209, pi :216, a_ij :221)
// and the standard Gaussian-mixture extension.  Statistics are accumulated in double:
//   gamma1[K]   += gamma_0[k]                      xi[K,K]     += sum_t xi_t(i,j)
//   occ[K,C]    += sum_t gamma_t[k] r_t[k,c]       sx[K,C,D]   += sum_t gamma_t[k] r_t[k,c] x_t
//   sxx[K,C,D]  += sum_t gamma_t[k] r_t[k,c] x_t^2           (r = component responsibility within the state)
// Kernels:
//   gmm_components_kernel   per-component log-likelihoods log w_kc + log N(x | mu_kc, var_kc)  [n, K*C]
//   bw_xi_kernel            xi and gamma_0 from the scaled forward/backward vectors left in the fb workspace
//   bw_gmm_stats_kernel     occ / sx / sxx: register-tiled (component x dim) outer-product accumulation, fp32 partials per
//                           CTA, one double atomicAdd per statistic per CTA
// Multi-GPU: each rank accumulates its shard; the host all-reduces the ~8 k doubles once per EM iteration (NCCL).
#include "common.cuh"

namespace hmmb200 {

__global__ void __launch_bounds__(128) gmm_components_kernel(const float *x, const float *packed, int64_t n, int KC, int D,
                                                             int NP2, float *comp, const float *skip_if_one) {
    if (skip_if_one != nullptr && *skip_if_one == 1.f) return;         // the tcgen05 emission kernel wrote comp already
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n * KC; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t fr = idx / KC;
    const int kc = (int)(idx % KC), pr = kc >> 1, hi = kc & 1;
    const float *xn = x + fr * D;
    float acc = 0.f;
    for (int d = 0; d < D; ++d) {
        const float u = fmaf(xn[d], __ldg(packed + ((size_t)d * NP2 + pr) * 4 + hi), __ldg(packed + ((size_t)d * NP2 + pr) * 4 + 2 + hi));
        acc = fmaf(u, u, acc);
    }
    comp[idx] = fmaf(-0.5f, acc, __ldg(packed + (size_t)D * NP2 * 4 + kc));
    }
}

// one warp per (sequence, block of frames); lane j owns column j of xi.  KP = K padded to a multiple of 4 bounds the unrolled loops.
template <int KP>
__global__ void __launch_bounds__(128) bw_xi_kernel(const float *emis, int mode, float eps, const float *trans,
                                                    const float *ws_a, const float *ws_b, const float *wseq, int B, int T, int K,
                                                    int frames_per_warp, double *xi, double *gamma1) {
    extern __shared__ double xi_s[];                       // [K*K] per CTA
    for (int i = threadIdx.x; i < K * K; i += blockDim.x) xi_s[i] = 0.0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int blocks_per_seq = (T - 1 + frames_per_warp - 1) / frames_per_warp;
    const int64_t wid = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
    const bool ok = lane < K;
    float col[KP];
#pragma unroll
    for (int i = 0; i < KP; ++i) col[i] = (ok && i < K) ? trans[i * K + lane] : 0.f;
    float acc[KP];
#pragma unroll
    for (int i = 0; i < KP; ++i) acc[i] = 0.f;
    if (wid < (int64_t)B * max(blocks_per_seq, 1) && T > 1) {
        const int b = (int)(wid / blocks_per_seq), blk = (int)(wid % blocks_per_seq);
        const int t0 = blk * frames_per_warp, t1 = min(T - 1, t0 + frames_per_warp);
        const float wb = wseq ? wseq[b] : 1.f;                 // per-sequence weight (autograd: d loss / d loglik_b)
        for (int t = t0; t < t1; ++t) {
            // u_j = b~_{t+1}(j) * beta_{t+1}(j)
            float e = ok ? emis[((size_t)b * T + t + 1) * K + lane] : 0.f, bt;
            if (mode == HMMB200_EMIS_PROB_FLOOR) bt = ok ? e + eps : 0.f;
            else if (mode == HMMB200_EMIS_LOG_EXP_FLOOR) bt = ok ? expf(e) + eps : 0.f;
            else {
                float mx = ok ? e : -INFINITY;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(FULL_MASK, mx, o));
                if (!(mx > -INFINITY)) mx = 0.f;
                bt = ok ? expf(e - mx) + ((mode == HMMB200_EMIS_LOG_NORM_FLOOR) ? eps : 0.f) : 0.f;
            }
            const float u = ok ? bt * ws_b[((size_t)b * T + t + 1) * K + lane] : 0.f;
            const float a = ok ? ws_a[((size_t)b * T + t) * K + lane] : 0.f;
            float v[KP], cs = 0.f;
#pragma unroll
            for (int i = 0; i < KP; ++i) { v[i] = __shfl_sync(FULL_MASK, a, i) * col[i] * u; cs += v[i]; }
            float Z = cs;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) Z += __shfl_xor_sync(FULL_MASK, Z, o);
            const float inv = (Z > 0.f) ? wb / Z : 0.f;
#pragma unroll
            for (int i = 0; i < KP; ++i) acc[i] = fmaf(v[i], inv, acc[i]);
        }
        if (blk == 0 && gamma1 != nullptr) {               // gamma_0 = a_0 .* b_0 / sum
            const float g = ok ? ws_a[(size_t)b * T * K + lane] * ws_b[(size_t)b * T * K + lane] : 0.f;
            float Z = g;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) Z += __shfl_xor_sync(FULL_MASK, Z, o);
            if (ok && Z > 0.f) atomicAdd(gamma1 + lane, (double)(wb * g / Z));
        }
    } else if (T == 1 && wid < B && gamma1 != nullptr) {
        const int b = (int)wid;
        const float g = ok ? ws_a[(size_t)b * K + lane] * ws_b[(size_t)b * K + lane] : 0.f;
        float Z = g;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) Z += __shfl_xor_sync(FULL_MASK, Z, o);
        if (ok && Z > 0.f) atomicAdd(gamma1 + lane, (double)((wseq ? wseq[b] : 1.f) * g / Z));
    }
    if (ok) {
#pragma unroll
        for (int i = 0; i < KP; ++i) if (i < K) atomicAdd(&xi_s[i * K + lane], (double)acc[i]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < K * K; i += blockDim.x) if (xi_s[i] != 0.0) atomicAdd(xi + i, xi_s[i]);
}

// occ / sx / sxx are one [K*C, N] x [N, D] product (twice: with x and with x^2) over the N frames of the batch.  Per CTA a tile
// of BW_F frames is staged in shared memory: w = gamma_t(k) * responsibility(c | k) as [f][KCp], x and x^2 as [f][Dp].  A thread
// owns a 4 (components) x 8 (dims) block of both accumulators and one of FG frame sub-sequences of the tile: per frame it reads
// five 16-byte vectors for 64 FMAs.  Partials stay in fp32 registers over the CTA's tiles and are committed once with double
// atomics (every statistic gets (CTAs x FG) adds per call).
constexpr int BW_F = 64;
constexpr int BW_TC = 4, BW_TD = 8;
__global__ void __launch_bounds__(512) bw_gmm_stats_kernel(const float *x, const float *comp, const float *logb, const float *gamma,
                                                            int64_t n, int K, int C, int D, int FG, double *occ, double *sx, double *sxx) {
    extern __shared__ __align__(16) float sm_bw[];
    const int KC = K * C;
    const int KCp = (KC + BW_TC - 1) / BW_TC * BW_TC, Dp = (D + BW_TD - 1) / BW_TD * BW_TD;
    float *w_s = sm_bw;                    // [BW_F][KCp]
    float *x_s = w_s + BW_F * KCp;         // [BW_F][Dp]
    float *x2_s = x_s + BW_F * Dp;         // [BW_F][Dp]
    // blockIdx.y selects a slice of the component groups when (K*C/4) x (D/8) cells exceed one CTA
    const int ndg = Dp / BW_TD, ncg_all = KCp / BW_TC;
    const int ncg = (ncg_all + gridDim.y - 1) / gridDim.y;
    const int cells = ncg * ndg;
    const int tid = threadIdx.x;
    const int fg = tid / cells, cell = tid % cells;        // the launch sizes the CTA to cells * FG threads
    const int gc = blockIdx.y * ncg + cell / ndg, gd = cell % ndg;
    const bool live = gc < ncg_all && fg < FG;               // the CTA is rounded up to whole warps: surplus threads only help staging
    float ax[BW_TC][BW_TD], axx[BW_TC][BW_TD], aocc[BW_TC];
#pragma unroll
    for (int i = 0; i < BW_TC; ++i) {
        aocc[i] = 0.f;
#pragma unroll
        for (int j = 0; j < BW_TD; ++j) { ax[i][j] = 0.f; axx[i][j] = 0.f; }
    }
    const int64_t n_tiles = (n + BW_F - 1) / BW_F;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t base = tile * BW_F;
        const int nf = (int)min((int64_t)BW_F, n - base);
        __syncthreads();
        // staging: one warp per frame row (no integer division by runtime sizes on the fill path)
        for (int f = tid >> 5; f < BW_F; f += (int)(blockDim.x >> 5)) {
            const int lane = tid & 31;
            const bool fok = f < nf;
            const int64_t fr = base + (fok ? f : 0);
            for (int kc = lane; kc < KCp; kc += 32) {
                float w = 0.f;
                if (fok && kc < KC) {
                    const int k = kc / C;
                    w = gamma[fr * K + k] * expf(comp[fr * KC + kc] - logb[fr * K + k]);   // gamma_t(k) * responsibility(c | k)
                }
                w_s[f * KCp + kc] = w;
            }
            for (int d = lane; d < Dp; d += 32) {
                const float v = (fok && d < D) ? x[fr * D + d] : 0.f;
                x_s[f * Dp + d] = v; x2_s[f * Dp + d] = v * v;
            }
        }
        __syncthreads();
#pragma unroll 2
        for (int f = fg; live && f < nf; f += FG) {
            const float4 w4 = *reinterpret_cast<const float4 *>(w_s + f * KCp + gc * BW_TC);
            const float4 xa = *reinterpret_cast<const float4 *>(x_s + f * Dp + gd * BW_TD);
            const float4 xb = *reinterpret_cast<const float4 *>(x_s + f * Dp + gd * BW_TD + 4);
            const float4 qa = *reinterpret_cast<const float4 *>(x2_s + f * Dp + gd * BW_TD);
            const float4 qb = *reinterpret_cast<const float4 *>(x2_s + f * Dp + gd * BW_TD + 4);
            const float wv[BW_TC] = {w4.x, w4.y, w4.z, w4.w};
            const float xv[BW_TD] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
            const float qv[BW_TD] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
#pragma unroll
            for (int i = 0; i < BW_TC; ++i) {
                aocc[i] += wv[i];
#pragma unroll
                for (int j = 0; j < BW_TD; ++j) { ax[i][j] = fmaf(wv[i], xv[j], ax[i][j]); axx[i][j] = fmaf(wv[i], qv[j], axx[i][j]); }
            }
        }
    }
    // fold the FG frame sub-sequences inside the CTA (through the staging buffer) so that only `cells` threads touch the global
    // double accumulators: every statistic then receives one atomic per CTA instead of FG
    constexpr int NACC = 2 * BW_TC * BW_TD + BW_TC;
    for (int r = 1; r < FG; ++r) {
        __syncthreads();
        float *slot = sm_bw + (size_t)cell * NACC;
        if (fg == r) {
#pragma unroll
            for (int i = 0; i < BW_TC; ++i) {
                slot[2 * BW_TC * BW_TD + i] = aocc[i];
#pragma unroll
                for (int j = 0; j < BW_TD; ++j) { slot[i * BW_TD + j] = ax[i][j]; slot[BW_TC * BW_TD + i * BW_TD + j] = axx[i][j]; }
            }
        }
        __syncthreads();
        if (fg == 0) {
#pragma unroll
            for (int i = 0; i < BW_TC; ++i) {
                aocc[i] += slot[2 * BW_TC * BW_TD + i];
#pragma unroll
                for (int j = 0; j < BW_TD; ++j) { ax[i][j] += slot[i * BW_TD + j]; axx[i][j] += slot[BW_TC * BW_TD + i * BW_TD + j]; }
            }
        }
    }
    if (fg != 0) return;
#pragma unroll
    for (int i = 0; i < BW_TC; ++i) {
        const int kc = gc * BW_TC + i;
        if (live && kc < KC) {
            if (gd == 0) atomicAdd(occ + kc, (double)aocc[i]);
#pragma unroll
            for (int j = 0; j < BW_TD; ++j) {
                const int d = gd * BW_TD + j;
                if (d < D) { atomicAdd(sx + (size_t)kc * D + d, (double)ax[i][j]); atomicAdd(sxx + (size_t)kc * D + d, (double)axx[i][j]); }
            }
        }
    }
}

template <int KP>
static int launch_xi_kp(const float *emis, int mode, float eps, const float *trans, const float *ws_a, const float *ws_b,
                        const float *wseq, int B, int T, int K, double *xi, double *gamma1, cudaStream_t s) {
    const int fpw = 64, warps = 4;
    const int blocks_per_seq = T > 1 ? (T - 1 + fpw - 1) / fpw : 1;
    const int64_t n_warps = (int64_t)B * blocks_per_seq;
    bw_xi_kernel<KP><<<(unsigned)((n_warps + warps - 1) / warps), warps * 32, (size_t)K * K * sizeof(double), s>>>(
        emis, mode, eps, trans, ws_a, ws_b, wseq, B, T, K, fpw, xi, gamma1);
    return check_launch("bw_xi_kernel");
}

static int launch_xi(const float *emis, int mode, float eps, const float *trans, const float *ws_a, const float *ws_b,
                     const float *wseq, int B, int T, int K, double *xi, double *gamma1, cudaStream_t s) {
    const int kp = pad4(K);
#define XI_CASE(N) if (kp <= N) return launch_xi_kp<N>(emis, mode, eps, trans, ws_a, ws_b, wseq, B, T, K, xi, gamma1, s)
    XI_CASE(4); XI_CASE(8); XI_CASE(12); XI_CASE(16); XI_CASE(20); XI_CASE(24); XI_CASE(28);
#undef XI_CASE
    return launch_xi_kp<32>(emis, mode, eps, trans, ws_a, ws_b, wseq, B, T, K, xi, gamma1, s);
}

static int launch_gmm_stats(const float *x, const float *comp, const float *logb, const float *gamma, int64_t n, int K, int C, int D,
                            double *occ, double *sx, double *sxx, cudaStream_t s) {
    const int KCp = (K * C + BW_TC - 1) / BW_TC * BW_TC, Dp = (D + BW_TD - 1) / BW_TD * BW_TD;
    size_t smem = (size_t)BW_F * (KCp + 2 * Dp) * sizeof(float);
    smem = smem > 512 * (2 * BW_TC * BW_TD + BW_TC) * sizeof(float) ? smem : 512 * (2 * BW_TC * BW_TD + BW_TC) * sizeof(float);   // also the fold buffer
    if (smem > 200 * 1024) return set_error(HMMB200_EUNSUPPORTED, "gmm_stats: K*C + D too large");
    if (smem > 48 * 1024) cudaFuncSetAttribute(bw_gmm_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t n_tiles = (n + BW_F - 1) / BW_F;
    const int ndg = Dp / BW_TD, ncg_all = KCp / BW_TC;
    if (ndg > 512) return set_error(HMMB200_EUNSUPPORTED, "gmm_stats: D = %d too large", D);
    int gy = 1;                                                     // slices of the component groups (<= 512 cells per CTA)
    while (((ncg_all + gy - 1) / gy) * ndg > 512) ++gy;
    const int cells = ((ncg_all + gy - 1) / gy) * ndg;
    int FG = 512 / cells;                                           // frame sub-sequences per tile: up to 512 threads per CTA
    FG = FG < 1 ? 1 : (FG > 8 ? 8 : FG);
    const int threads = (cells * FG + 31) & ~31;                   // whole warps (the staging loops are warp-per-row)
    dim3 grid((unsigned)min((int64_t)sms, n_tiles), (unsigned)gy);      // one CTA per SM (128 registers x ~480 threads)
    bw_gmm_stats_kernel<<<grid, threads, smem, s>>>(x, comp, logb, gamma, n, K, C, D, FG, occ, sx, sxx);
    return check_launch("bw_gmm_stats_kernel");
}

}  // namespace hmmb200

using namespace hmmb200;

HMMB200_EXPORT int hmmb200_gmm_components_f32(const float *x, const float *packed, int64_t n_frames, int K, int C, int D,
                                              float *comp, void *stream) {
    if (n_frames < 0 || K <= 0 || C <= 0 || D <= 0) return set_error(HMMB200_EINVAL, "gmm_components: bad shape");
    if (n_frames == 0) return HMMB200_OK;
    if (!x || !packed || !comp) return set_error(HMMB200_EINVAL, "gmm_components: null argument");
    if (int rc = require_sm100()) return rc;
    const int KC = K * C;
    const int64_t total = n_frames * KC;
    gmm_components_kernel<<<(unsigned)min((total + 127) / 128, (int64_t)148 * 16), 128, 0, (cudaStream_t)stream>>>(x, packed, n_frames, KC, D, (KC + 1) / 2, comp, nullptr);
    return check_launch("gmm_components_kernel");
}

// log b AND the per-component values in one pass over x: the tcgen05 emission kernel holds the components in its epilogue just
// before the mixture log-sum-exp and writes both; when the pack is outside its range the two fp32 kernels run instead.
HMMB200_EXPORT int hmmb200_gmm_emission_components_f32(const float *x, const float *packed, int64_t n_frames, int K, int C, int D,
                                                       float *logb, float *comp, void *stream) {
    if (!comp) return set_error(HMMB200_EINVAL, "gmm_emission_components: null argument");
    const float *flag = nullptr;
    if (int rc = gmm_emission_dispatch(x, packed, n_frames, K, C, D, logb, comp, (cudaStream_t)stream, 0, &flag)) return rc;
    if (n_frames == 0) return HMMB200_OK;
    const int KC = K * C;
    const int64_t total = n_frames * KC;
    gmm_components_kernel<<<(unsigned)min((total + 127) / 128, (int64_t)148 * 16), 128, 0, (cudaStream_t)stream>>>(x, packed, n_frames, KC, D, (KC + 1) / 2, comp, flag);
    return check_launch("gmm_components_kernel");
}

// stats layout (doubles): gamma1[K] | xi[K*K] | occ[K*C] | sx[K*C*D] | sxx[K*C*D]   (accumulated; zero it before the first call)
HMMB200_EXPORT size_t hmmb200_bw_stats_doubles(int K, int C, int D) {
    if (K <= 0 || C <= 0 || D <= 0) return 0;
    return (size_t)K + (size_t)K * K + (size_t)K * C + 2 * (size_t)K * C * D;
}

HMMB200_EXPORT int hmmb200_bw_accumulate_f32(const float *x, const float *comp, const float *logb, const float *gamma,
                                             const float *emis, int emis_mode, float floor_eps, const float *trans_prob,
                                             const void *fb_workspace, int B, int T, int K, int C, int D,
                                             double *stats, void *stream) {
    if (B < 0 || T < 0 || K <= 0 || C <= 0 || D <= 0) return set_error(HMMB200_EINVAL, "bw_accumulate: bad shape");
    if (B == 0 || T == 0) return HMMB200_OK;
    if (K > 32) return set_error(HMMB200_EUNSUPPORTED, "bw_accumulate: K <= 32 (got %d)", K);
    if (!x || !comp || !logb || !gamma || !emis || !trans_prob || !fb_workspace || !stats)
        return set_error(HMMB200_EINVAL, "bw_accumulate: null argument");
    if (int rc = require_sm100()) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    double *gamma1 = stats, *xi = stats + K, *occ = xi + (size_t)K * K, *sx = occ + (size_t)K * C, *sxx = sx + (size_t)K * C * D;
    const size_t n = (size_t)B * T;
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const float *ws_a = (const float *)fb_workspace;
    const float *ws_b = (const float *)((const uint8_t *)fb_workspace + al(n * K * sizeof(float)));
    if (int rc = launch_xi(emis, emis_mode, floor_eps, trans_prob, ws_a, ws_b, nullptr, B, T, K, xi, gamma1, s)) return rc;
    return launch_gmm_stats(x, comp, logb, gamma, (int64_t)n, K, C, D, occ, sx, sxx, s);
}

HMMB200_EXPORT int hmmb200_gmm_stats_f32(const float *x, const float *comp, const float *logb, const float *weight, int64_t n_frames,
                                         int K, int C, int D, double *occ, double *sx, double *sxx, void *stream) {
    if (n_frames < 0 || K <= 0 || C <= 0 || D <= 0) return set_error(HMMB200_EINVAL, "gmm_stats: bad shape");
    if (n_frames == 0) return HMMB200_OK;
    if (!x || !comp || !logb || !weight || !occ || !sx || !sxx) return set_error(HMMB200_EINVAL, "gmm_stats: null argument");
    if (int rc = require_sm100()) return rc;
    return launch_gmm_stats(x, comp, logb, weight, n_frames, K, C, D, occ, sx, sxx, (cudaStream_t)stream);
}

// Weighted transition / initial-state statistics alone: xi[i][j] += sum_b w_b sum_t xi_t(i,j), gamma1[k] += sum_b w_b gamma_0(k).
// With w_b = d loss / d loglik_b these are the gradients of the loss w.r.t. log P and log p0.
HMMB200_EXPORT int hmmb200_xi_sum_f32(const float *emis, int emis_mode, float floor_eps, const float *trans_prob,
                                      const void *fb_workspace, const float *seq_weights, int B, int T, int K,
                                      double *xi, double *gamma1, void *stream) {
    if (B < 0 || T < 0 || K <= 0) return set_error(HMMB200_EINVAL, "xi_sum: bad shape");
    if (B == 0 || T == 0) return HMMB200_OK;
    if (K > 32) return set_error(HMMB200_EUNSUPPORTED, "xi_sum: K <= 32 (got %d)", K);
    if (!emis || !trans_prob || !fb_workspace || !xi) return set_error(HMMB200_EINVAL, "xi_sum: null argument");
    if (int rc = require_sm100()) return rc;
    const size_t n = (size_t)B * T;
    auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
    const float *ws_a = (const float *)fb_workspace;
    const float *ws_b = (const float *)((const uint8_t *)fb_workspace + al(n * K * sizeof(float)));
    return launch_xi(emis, emis_mode, floor_eps, trans_prob, ws_a, ws_b, seq_weights, B, T, K, xi, gamma1, (cudaStream_t)stream);
}
