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

#include <stdlib.h>

namespace hmmb200 {

// bw_tc.cu: the statistics GEMM on tcgen05 (0 launched, 1 shape not covered, < 0 error)
int launch_gmm_stats_tc(const float *x, const float *comp, const float *logb, const float *gamma, int64_t n_full, int K, int C, int D,
                        double *occ, double *sx, double *sxx, cudaStream_t s);

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

// xi_t(i,j) = a_t(i) P(i,j) u_{t+1}(j) / Z_t,  u = b~ .* beta (scaled vectors from the fb workspace),  Z_t = sum_ij of the numerator.
// One warp per (sequence, block of frames).  LPF lanes share a frame and split the K rows of xi (R = ceil(K / LPF) rows each), so a
// warp works on 32 / LPF consecutive frames at once: the three input rows of those frames are contiguous in memory (coalesced), and a
// lane spends 3 R KP multiply-adds per frame on R KP cells -- P(i,.) u(.) products, their row sums for Z, then one FMA per cell with
// a(i) / Z -- instead of one warp-wide shuffle per row of every single frame.  KP = K padded to a multiple of 4.
template <int KP, int LPF>
__global__ void __launch_bounds__(384) bw_xi_kernel(const float *emis, int mode, float eps, const float *trans,
                                                    const float *ws_a, const float *ws_b, const float *wseq, int B, int T, int K,
                                                    int frames_per_warp, double *xi, double *gamma1) {
    extern __shared__ double xi_s[];                       // [K*K + K] per CTA: xi, gamma_0
    constexpr int R = (KP + LPF - 1) / LPF, FPI = 32 / LPF;    // rows per lane, frames per warp iteration
    double *g1_s = xi_s + K * K;
    for (int i = threadIdx.x; i < K * K + K; i += blockDim.x) xi_s[i] = 0.0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int fr = lane / LPF, part = lane % LPF;
    const int blocks_per_seq = max((T - 1 + frames_per_warp - 1) / frames_per_warp, 1);
    const int64_t n_tasks = (int64_t)B * blocks_per_seq;
    const int64_t n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    float Prow[R][KP], acc[R][KP];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int j = 0; j < KP; ++j) {
            const int i = part * R + r;
            Prow[r][j] = (i < K && j < K) ? trans[i * K + j] : 0.f;
            acc[r][j] = 0.f;
        }
    // Persistent warps: a warp walks tasks (sequence, block of frames) with its xi partial sums in registers and commits them ONCE;
    // with a CTA per four tasks every one of the K*K global accumulators took ~2000 same-address atomics per call, which serialise
    // in the L2 and were most of the kernel's time.
    for (int64_t task = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp; task < n_tasks; task += n_warps) {
        const int b = (int)(task / blocks_per_seq), blk = (int)(task % blocks_per_seq);
        const float wb = wseq ? wseq[b] : 1.f;                 // per-sequence weight (autograd: d loss / d loglik_b)
        if (T > 1) {
            const int t0 = blk * frames_per_warp, t1 = min(T - 1, t0 + frames_per_warp);
            for (int tb = t0; tb < t1; tb += FPI) {
                const int t = tb + fr;
                const bool live = t < t1;
                const size_t row1 = ((size_t)b * T + (live ? t + 1 : t0 + 1)) * K, row0 = ((size_t)b * T + (live ? t : t0)) * K;
                // u_j = b~_{t+1}(j) * beta_{t+1}(j)
                float u[KP], bb[KP];
                float mx = -INFINITY;
#pragma unroll
                for (int j = 0; j < KP; ++j) { u[j] = (j < K) ? emis[row1 + j] : 0.f; bb[j] = (j < K) ? ws_b[row1 + j] : 0.f; }
                float a[R];
#pragma unroll
                for (int r = 0; r < R; ++r) { const int i = part * R + r; a[r] = (i < K) ? ws_a[row0 + i] : 0.f; }
#pragma unroll
                for (int j = 0; j < KP; ++j) if (j < K) mx = fmaxf(mx, u[j]);
                if (!(mx > -INFINITY)) mx = 0.f;
#pragma unroll
                for (int j = 0; j < KP; ++j) {
                    float bt;
                    if (mode == HMMB200_EMIS_PROB_FLOOR) bt = u[j] + eps;
                    else if (mode == HMMB200_EMIS_LOG_EXP_FLOOR) bt = expf(u[j]) + eps;
                    else bt = expf(u[j] - mx) + ((mode == HMMB200_EMIS_LOG_NORM_FLOOR) ? eps : 0.f);
                    u[j] = (j < K) ? bt * bb[j] : 0.f;
                }
                float pu[R][KP], z = 0.f;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    float rs = 0.f;
#pragma unroll
                    for (int j = 0; j < KP; ++j) { pu[r][j] = Prow[r][j] * u[j]; rs += pu[r][j]; }
                    z = fmaf(a[r], rs, z);
                }
#pragma unroll
                for (int o = LPF / 2; o > 0; o >>= 1) z += __shfl_xor_sync(FULL_MASK, z, o);
                const float inv = (live && z > 0.f) ? wb / z : 0.f;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const float sc = a[r] * inv;
#pragma unroll
                    for (int j = 0; j < KP; ++j) acc[r][j] = fmaf(sc, pu[r][j], acc[r][j]);
                }
            }
        }
        if (blk == 0 && gamma1 != nullptr) {               // gamma_0 = a_0 .* b_0 / sum
            const bool ok = lane < K;
            const float g = ok ? ws_a[(size_t)b * T * K + lane] * ws_b[(size_t)b * T * K + lane] : 0.f;
            float Z = g;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) Z += __shfl_xor_sync(FULL_MASK, Z, o);
            if (ok && Z > 0.f) atomicAdd(g1_s + lane, (double)(wb * g / Z));
        }
    }
    // fold the frame groups of the warp (lanes that share `part`), then one shared-memory double add per cell and warp
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int j = 0; j < KP; ++j) {
            float v = acc[r][j];
#pragma unroll
            for (int o = LPF; o < 32; o <<= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
            const int i = part * R + r;
            if (fr == 0 && i < K && j < K && v != 0.f) atomicAdd(&xi_s[i * K + j], (double)v);
        }
    __syncthreads();
    for (int i = threadIdx.x; i < K * K; i += blockDim.x) if (xi_s[i] != 0.0) atomicAdd(xi + i, xi_s[i]);
    if (gamma1 != nullptr)
        for (int i = threadIdx.x; i < K; i += blockDim.x) if (g1_s[i] != 0.0) atomicAdd(gamma1 + i, g1_s[i]);
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

// The same statistics with the tile feed taken off the compute warps (the form used when K*C % 4 == 0, D % 8 == 0 and the tensors are
// 16-byte aligned, i.e. every BASELINE shape).  A tile of BW_F frames is FOUR contiguous byte ranges of the inputs (x, comp, gamma, log b
// are dense [n, .] arrays): one thread moves them with cp.async.bulk into a BW_STAGES-deep shared-memory ring, completion counted on an
// mbarrier per stage, so the HBM latency of tile t+2 hides behind the FMAs of tile t (the first version staged every tile with ordinary
// loads between two block barriers: ~0.9 ms per 512 000 frames, eight times the FMA time).  Per tile: wait, turn comp into
// w = gamma * exp(comp - log b) ([BW_F][K*C], 6 elements per thread), barrier, 4x8 register-tile FMAs straight from the raw x rows
// (x^2 formed in registers), barrier, refill the stage.
constexpr int BW_STAGES = 3;
__device__ __forceinline__ uint32_t bw_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bw_mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bw_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void bw_mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bw_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bw_mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "BWWAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra BWDONE_%=;\n\t"
        "bra BWWAIT_%=;\n\t"
        "BWDONE_%=:\n\t"
        "}" ::"r"(bw_smem_u32(bar)), "r"(parity), "r"(1000000u) : "memory");
}
__device__ __forceinline__ void bw_bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(bw_smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(bw_smem_u32(bar)) : "memory");
}

__device__ __forceinline__ float2 bw_ffma2(float2 a, float2 b, float2 c) {
    unsigned long long ra = *reinterpret_cast<unsigned long long *>(&a), rb = *reinterpret_cast<unsigned long long *>(&b);
    unsigned long long rc = *reinterpret_cast<unsigned long long *>(&c), rd;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2 *>(&rd);
}
__device__ __forceinline__ float2 bw_mul2(float2 a, float2 b) {
    unsigned long long ra = *reinterpret_cast<unsigned long long *>(&a), rb = *reinterpret_cast<unsigned long long *>(&b), rd;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    return *reinterpret_cast<float2 *>(&rd);
}

__global__ void __launch_bounds__(512, 1) bw_gmm_stats_bulk_kernel(const float *x, const float *comp, const float *logb, const float *gamma,
                                                                    int64_t n, int K, int C, int D, int FG, double *occ, double *sx, double *sxx) {
    extern __shared__ __align__(16) float sm_bw[];
    const int KC = K * C;
    // ring stage: [x: F*D][comp: F*KC][gamma: F*K][logb: F*K] floats, every piece a multiple of 16 bytes (F = 64)
    const int stage_floats = BW_F * (D + KC + 2 * K);
    uint64_t *full = reinterpret_cast<uint64_t *>(sm_bw);                  // [BW_STAGES]
    float *ring = sm_bw + 16;
    float *w_s = ring + (size_t)BW_STAGES * stage_floats;                  // [BW_F][KC]
    const int ndg = D / BW_TD, ncg = KC / BW_TC;
    const int cells = ncg * ndg;
    const int tid = threadIdx.x;
    const int fg = tid / cells, cell = tid % cells;
    // component group fastest: the lanes of a quarter-warp read ONE x address (broadcast) and consecutive 16-byte pieces of the w row,
    // i.e. every 16-byte shared-memory load of the FMA loop is conflict-free (with the dim group fastest the x loads of neighbouring
    // lanes sat 32 bytes apart: two-way bank conflicts, and the loop was bound by shared-memory wavefronts instead of FMAs)
    const int gc = cell % ncg, gd = cell / ncg;
    const bool live = fg < FG;
    float2 ax[BW_TC][BW_TD / 2], axx[BW_TC][BW_TD / 2];
    float aocc[BW_TC];
#pragma unroll
    for (int i = 0; i < BW_TC; ++i) {
        aocc[i] = 0.f;
#pragma unroll
        for (int j = 0; j < BW_TD / 2; ++j) { ax[i][j] = make_float2(0.f, 0.f); axx[i][j] = make_float2(0.f, 0.f); }
    }
    const int64_t n_tiles = n / BW_F;                                      // full tiles only (the launcher sends the tail elsewhere)
    const int64_t my_tiles = (n_tiles > blockIdx.x) ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    auto issue = [&](int64_t it) {                                         // this CTA's it-th tile -> stage it % BW_STAGES
        const int64_t base = (blockIdx.x + it * gridDim.x) * BW_F;
        float *st = ring + (size_t)(it % BW_STAGES) * stage_floats;
        uint64_t *bar = full + (it % BW_STAGES);
        const uint32_t bx = BW_F * D * 4, bc = BW_F * KC * 4, bk = BW_F * K * 4;
        bw_mbar_expect_tx(bar, bx + bc + 2 * bk);
        bw_bulk_g2s(st, x + base * D, bx, bar);
        bw_bulk_g2s(st + BW_F * D, comp + base * KC, bc, bar);
        bw_bulk_g2s(st + BW_F * (D + KC), gamma + base * K, bk, bar);
        bw_bulk_g2s(st + BW_F * (D + KC + K), logb + base * K, bk, bar);
    };
    if (tid == 0) {
        for (int i = 0; i < BW_STAGES; ++i) bw_mbar_init(full + i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int64_t it = 0; it < min((int64_t)BW_STAGES, my_tiles); ++it) issue(it);
    }
    __syncthreads();
    // element walk of the w pass without integer divisions in the loop: e = tid + i * blockDim  ->  (f, kc)
    const int step_f = (int)blockDim.x / KC, step_kc = (int)blockDim.x % KC;
    const float inv_c = 1.f / (float)C;
    for (int64_t it = 0; it < my_tiles; ++it) {
        const float *st = ring + (size_t)(it % BW_STAGES) * stage_floats;
        bw_mbar_wait(full + (it % BW_STAGES), (uint32_t)((it / BW_STAGES) & 1));
        const float *x_s = st, *c_s = st + BW_F * D, *g_s = st + BW_F * (D + KC), *l_s = st + BW_F * (D + KC + K);
        {
            int f = tid / KC, kc = tid % KC;
            for (int e = tid; e < BW_F * KC; e += blockDim.x) {
                const int k = __float2int_rz(((float)kc + 0.5f) * inv_c);
                w_s[e] = g_s[f * K + k] * __expf(c_s[e] - l_s[f * K + k]);     // gamma_t(k) * responsibility(c | k)
                f += step_f; kc += step_kc;
                if (kc >= KC) { kc -= KC; ++f; }
            }
        }
        __syncthreads();
        if (live) {
#pragma unroll 2
            for (int f = fg; f < BW_F; f += FG) {
                const float4 w4 = *reinterpret_cast<const float4 *>(w_s + f * KC + gc * BW_TC);
                const float4 xa = *reinterpret_cast<const float4 *>(x_s + f * D + gd * BW_TD);
                const float4 xb = *reinterpret_cast<const float4 *>(x_s + f * D + gd * BW_TD + 4);
                const float wv[BW_TC] = {w4.x, w4.y, w4.z, w4.w};
                const float2 xv[BW_TD / 2] = {make_float2(xa.x, xa.y), make_float2(xa.z, xa.w), make_float2(xb.x, xb.y), make_float2(xb.z, xb.w)};
                float2 qv[BW_TD / 2];
#pragma unroll
                for (int j = 0; j < BW_TD / 2; ++j) qv[j] = bw_mul2(xv[j], xv[j]);
#pragma unroll
                for (int i = 0; i < BW_TC; ++i) {
                    aocc[i] += wv[i];
                    const float2 ww = make_float2(wv[i], wv[i]);
#pragma unroll
                    for (int j = 0; j < BW_TD / 2; ++j) { ax[i][j] = bw_ffma2(ww, xv[j], ax[i][j]); axx[i][j] = bw_ffma2(ww, qv[j], axx[i][j]); }
                }
            }
        }
        __syncthreads();                                                   // every thread is done with the stage and with w_s
        if (tid == 0 && it + BW_STAGES < my_tiles) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy reads above before the bulk copy's writes
            issue(it + BW_STAGES);
        }
    }
    // fold the FG frame sub-sequences inside the CTA, then one double atomic per statistic and CTA (as bw_gmm_stats_kernel)
    constexpr int NACC = 2 * BW_TC * BW_TD + BW_TC;
    float *fold = ring;
    for (int r = 1; r < FG; ++r) {
        __syncthreads();
        float *slot = fold + (size_t)cell * NACC;
        if (fg == r) {
#pragma unroll
            for (int i = 0; i < BW_TC; ++i) {
                slot[2 * BW_TC * BW_TD + i] = aocc[i];
#pragma unroll
                for (int j = 0; j < BW_TD / 2; ++j) {
                    slot[i * BW_TD + 2 * j] = ax[i][j].x; slot[i * BW_TD + 2 * j + 1] = ax[i][j].y;
                    slot[BW_TC * BW_TD + i * BW_TD + 2 * j] = axx[i][j].x; slot[BW_TC * BW_TD + i * BW_TD + 2 * j + 1] = axx[i][j].y;
                }
            }
        }
        __syncthreads();
        if (fg == 0) {
#pragma unroll
            for (int i = 0; i < BW_TC; ++i) {
                aocc[i] += slot[2 * BW_TC * BW_TD + i];
#pragma unroll
                for (int j = 0; j < BW_TD / 2; ++j) {
                    ax[i][j].x += slot[i * BW_TD + 2 * j]; ax[i][j].y += slot[i * BW_TD + 2 * j + 1];
                    axx[i][j].x += slot[BW_TC * BW_TD + i * BW_TD + 2 * j]; axx[i][j].y += slot[BW_TC * BW_TD + i * BW_TD + 2 * j + 1];
                }
            }
        }
    }
    if (fg != 0 || my_tiles == 0) return;
#pragma unroll
    for (int i = 0; i < BW_TC; ++i) {
        const int kc = gc * BW_TC + i;
        if (gd == 0) atomicAdd(occ + kc, (double)aocc[i]);
#pragma unroll
        for (int j = 0; j < BW_TD / 2; ++j) {
            const int d = gd * BW_TD + 2 * j;
            atomicAdd(sx + (size_t)kc * D + d, (double)ax[i][j].x);
            atomicAdd(sx + (size_t)kc * D + d + 1, (double)ax[i][j].y);
            atomicAdd(sxx + (size_t)kc * D + d, (double)axx[i][j].x);
            atomicAdd(sxx + (size_t)kc * D + d + 1, (double)axx[i][j].y);
        }
    }
}

// The same [K*C, frames] x [frames, 2D] product on the tensor cores: mma.sync m16n8k8 TF32 with a 3-term hi/lo split
// (hi = the value with its low 13 mantissa bits cleared, lo = value - hi, both exact; hi*hi + hi*lo + lo*hi leaves a relative error
// of ~2^-21 per product), fp32 accumulators in registers over all the CTA's tiles.  The FMA form above peaks at 57 % of the fp32 pipe
// (0.25 ms per 512 000 frames); this one needs a third of its instructions.  Same bulk-copy ring and w pass.
//   16 warps = 4 frame groups (16 frames = 2 k-steps of a 64-frame tile each) x 4 column groups (x dims [0, D/2), [D/2, D), then the
//   same two ranges of x^2); a warp holds MT x NTW accumulator tiles (MT = ceil(K*C / 16) <= 4, NTW = ceil(D / 16) <= 6) plus, in the
//   first column group, one tile against a column of ones (the occupancies).
__device__ __forceinline__ void bw_mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void bw_split(float v, uint32_t &hi, uint32_t &lo) {
    hi = __float_as_uint(v) & 0xffffe000u;
    lo = __float_as_uint(v - __uint_as_float(hi));
}

template <int MT, int NTW>
__global__ void __launch_bounds__(512, 1) bw_gmm_stats_mma_kernel(const float *x, const float *comp, const float *logb, const float *gamma,
                                                                   int64_t n, int K, int C, int D, double *occ, double *sx, double *sxx) {
    extern __shared__ __align__(16) float sm_bw[];
    const int KC = K * C;
    constexpr int WP = MT * 16 + 8;                                        // w row pitch: the A-fragment loads (4 frames x 8 rows) are conflict-free
    const int stage_floats = BW_F * (D + KC + 2 * K);
    uint64_t *full = reinterpret_cast<uint64_t *>(sm_bw);
    float *ring = sm_bw + 16;
    float *w_s = ring + (size_t)BW_STAGES * stage_floats;                  // [BW_F][WP], columns KC .. WP-1 stay zero
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t4 = lane & 3;
    const int fgp = warp >> 2, cgp = warp & 3;                             // frame group, column group
    const bool sq = cgp >= 2;                                              // this warp's columns are x^2
    const int nt_all = D / 8;                                              // n-tiles of x (and of x^2)
    const int nt0 = (cgp & 1) * NTW;                                       // first n-tile of the warp within its half
    float acc[MT][NTW][4], aocc[MT][4];
#pragma unroll
    for (int m = 0; m < MT; ++m) {
#pragma unroll
        for (int q = 0; q < 4; ++q) aocc[m][q] = 0.f;
#pragma unroll
        for (int j = 0; j < NTW; ++j)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[m][j][q] = 0.f;
    }
    const int64_t n_tiles = n / BW_F;
    const int64_t my_tiles = (n_tiles > blockIdx.x) ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    auto issue = [&](int64_t it) {
        const int64_t base = (blockIdx.x + it * gridDim.x) * BW_F;
        float *st = ring + (size_t)(it % BW_STAGES) * stage_floats;
        uint64_t *bar = full + (it % BW_STAGES);
        const uint32_t bx = BW_F * D * 4, bc = BW_F * KC * 4, bk = BW_F * K * 4;
        bw_mbar_expect_tx(bar, bx + bc + 2 * bk);
        bw_bulk_g2s(st, x + base * D, bx, bar);
        bw_bulk_g2s(st + BW_F * D, comp + base * KC, bc, bar);
        bw_bulk_g2s(st + BW_F * (D + KC), gamma + base * K, bk, bar);
        bw_bulk_g2s(st + BW_F * (D + KC + K), logb + base * K, bk, bar);
    };
    if (tid == 0) {
        for (int i = 0; i < BW_STAGES; ++i) bw_mbar_init(full + i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int64_t it = 0; it < min((int64_t)BW_STAGES, my_tiles); ++it) issue(it);
    }
    for (int e = tid; e < BW_F * WP; e += blockDim.x) w_s[e] = 0.f;
    __syncthreads();
    const int step_f = (int)blockDim.x / KC, step_kc = (int)blockDim.x % KC;
    const float inv_c = 1.f / (float)C;
    const uint32_t one = __float_as_uint(1.f);
    for (int64_t it = 0; it < my_tiles; ++it) {
        const float *st = ring + (size_t)(it % BW_STAGES) * stage_floats;
        bw_mbar_wait(full + (it % BW_STAGES), (uint32_t)((it / BW_STAGES) & 1));
        const float *x_s = st, *c_s = st + BW_F * D, *g_s = st + BW_F * (D + KC), *l_s = st + BW_F * (D + KC + K);
        {
            int f = tid / KC, kc = tid % KC;
            for (int e = tid; e < BW_F * KC; e += blockDim.x) {
                const int k = __float2int_rz(((float)kc + 0.5f) * inv_c);
                w_s[f * WP + kc] = g_s[f * K + k] * __expf(c_s[e] - l_s[f * K + k]);   // gamma_t(k) * responsibility(c | k)
                f += step_f; kc += step_kc;
                if (kc >= KC) { kc -= KC; ++f; }
            }
        }
        __syncthreads();
#pragma unroll
        for (int ks = 0; ks < BW_F / 8 / 4; ++ks) {                        // this warp's k-steps (8 frames each) of the tile
            const int f0 = (fgp * (BW_F / 8 / 4) + ks) * 8;
            uint32_t ah[MT][4], al[MT][4];
#pragma unroll
            for (int m = 0; m < MT; ++m) {
                const float *wr = w_s + (f0 + t4) * WP + m * 16 + g;       // A(row = kc, col = frame) = w[frame][kc]
                bw_split(wr[0], ah[m][0], al[m][0]);
                bw_split(wr[8], ah[m][1], al[m][1]);
                bw_split(wr[4 * WP], ah[m][2], al[m][2]);
                bw_split(wr[4 * WP + 8], ah[m][3], al[m][3]);
            }
#pragma unroll
            for (int j = 0; j < NTW; ++j) {
                const int nt = nt0 + j;
                if (nt < nt_all && nt < (cgp & 1) * NTW + NTW) {
                    const float *xr = x_s + (f0 + t4) * D + nt * 8 + g;    // B(row = frame, col = dim) = x[frame][dim]
                    float b0 = xr[0], b1 = xr[4 * D];
                    if (sq) { b0 *= b0; b1 *= b1; }
                    uint32_t bh[2], bl[2];
                    bw_split(b0, bh[0], bl[0]);
                    bw_split(b1, bh[1], bl[1]);
#pragma unroll
                    for (int m = 0; m < MT; ++m) {
                        bw_mma_tf32(acc[m][j], ah[m], bh);
                        bw_mma_tf32(acc[m][j], ah[m], bl);
                        bw_mma_tf32(acc[m][j], al[m], bh);
                    }
                }
            }
            if (cgp == 0) {                                                 // occupancies: w^T x 1
                const uint32_t b1v[2] = {one, one};
#pragma unroll
                for (int m = 0; m < MT; ++m) { bw_mma_tf32(aocc[m], ah[m], b1v); bw_mma_tf32(aocc[m], al[m], b1v); }
            }
        }
        __syncthreads();                                                   // every thread is done with the stage and with w_s
        if (tid == 0 && it + BW_STAGES < my_tiles) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue(it + BW_STAGES);
        }
    }
    if (my_tiles == 0) return;
    // fold the four frame groups through shared memory (fp32), then one double atomic per statistic and CTA.
    // accumulator element q of tile (m, j): row kc = 16 m + g + 8 (q >> 1), column d = 8 nt + 2 t4 + (q & 1)
    float *fold = ring;                                                    // [MT*16][2*D + 1] floats
    const int FW = 2 * D + 1;
    for (int e = tid; e < MT * 16 * FW; e += blockDim.x) fold[e] = 0.f;
    __syncthreads();
    for (int r = 0; r < 4; ++r) {
        if (fgp == r) {
#pragma unroll
            for (int m = 0; m < MT; ++m) {
#pragma unroll
                for (int j = 0; j < NTW; ++j) {
                    const int nt = nt0 + j;
                    if (nt < nt_all && nt < (cgp & 1) * NTW + NTW) {
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            fold[(m * 16 + g + 8 * (q >> 1)) * FW + (sq ? D : 0) + nt * 8 + 2 * t4 + (q & 1)] += acc[m][j][q];
                    }
                }
                if (cgp == 0 && t4 == 0) {
                    fold[(m * 16 + g) * FW + 2 * D] += aocc[m][0];
                    fold[(m * 16 + g + 8) * FW + 2 * D] += aocc[m][2];
                }
            }
        }
        __syncthreads();
    }
    for (int e = tid; e < KC * FW; e += blockDim.x) {
        const int kc = e / FW, c = e % FW;
        const double v = (double)fold[e];
        if (c < D) atomicAdd(sx + (size_t)kc * D + c, v);
        else if (c < 2 * D) atomicAdd(sxx + (size_t)kc * D + (c - D), v);
        else atomicAdd(occ + kc, v);
    }
}

// The same statistics with the rows staged by bulk copies.  bw_xi_kernel loads a frame's 27 values with scalar loads and waits for
// them before it computes -- 12 warps per SM with nothing in flight while they compute: 1.3 TB/s of HBM (ncu: a third of the samples
// on the first use of the loaded data).  Here a warp's NEXT task (32 frames x three arrays, 4.6 KB) travels as three cp.async.bulk
// copies into the warp's other staging buffer while the current one is consumed from shared memory.  Needs K % 4 == 0 (16-byte rows),
// K == KP and T > 1; same arithmetic per frame, same order of the frames within a warp's accumulators.
constexpr int XI_FPW = 32;
template <int KP, int LPF>
__global__ void __launch_bounds__(384) bw_xi_bulk_kernel(const float *emis, int mode, float eps, const float *trans,
                                                         const float *ws_a, const float *ws_b, const float *wseq, int B, int T,
                                                         double *xi, double *gamma1) {
    extern __shared__ __align__(16) uint8_t xi_raw[];
    constexpr int K = KP, R = (KP + LPF - 1) / LPF, FPI = 32 / LPF, KV = KP / 4;
    constexpr int BUF = XI_FPW * KP;                           // floats per staged array
    const int nw = blockDim.x >> 5;
    uint64_t *bars = reinterpret_cast<uint64_t *>(xi_raw);     // [nw][2]
    double *xi_s = reinterpret_cast<double *>(xi_raw + (((size_t)nw * 2 * sizeof(uint64_t) + 15) & ~(size_t)15));
    double *g1_s = xi_s + K * K;
    float *stg_all = reinterpret_cast<float *>(reinterpret_cast<uint8_t *>(xi_s) + (((size_t)(K * K + K) * sizeof(double) + 15) & ~(size_t)15));
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int fr = lane / LPF, part = lane % LPF;
    float *stg = stg_all + (size_t)warp * 2 * 3 * BUF;
    uint64_t *bar = bars + warp * 2;
    for (int i = threadIdx.x; i < K * K + K; i += blockDim.x) xi_s[i] = 0.0;
    if (lane == 0) { bw_mbar_init(bar, 1); bw_mbar_init(bar + 1, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    const int bps = (T - 1 + XI_FPW - 1) / XI_FPW;
    const int64_t n_tasks = (int64_t)B * bps, n_warps = (int64_t)gridDim.x * nw;
    float Prow[R][KP], acc[R][KP];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int j = 0; j < KP; ++j) {
            const int i = part * R + r;
            Prow[r][j] = (i < K) ? trans[i * K + j] : 0.f;
            acc[r][j] = 0.f;
        }
    auto issue = [&](int64_t task, int buf) {                  // lane 0: the three arrays of a task into staging buffer `buf`
        const int b = (int)(task / bps), blk = (int)(task % bps);
        const int t0 = blk * XI_FPW, n = min(T - 1, t0 + XI_FPW) - t0;
        const uint32_t bytes = (uint32_t)n * K * sizeof(float);
        float *dst = stg + (size_t)buf * 3 * BUF;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the buffer's earlier generic reads come first
        bw_mbar_expect_tx(bar + buf, 3 * bytes);
        bw_bulk_g2s(dst, emis + ((size_t)b * T + t0 + 1) * K, bytes, bar + buf);
        bw_bulk_g2s(dst + BUF, ws_b + ((size_t)b * T + t0 + 1) * K, bytes, bar + buf);
        bw_bulk_g2s(dst + 2 * BUF, ws_a + ((size_t)b * T + t0) * K, bytes, bar + buf);
    };
    int64_t task = (int64_t)blockIdx.x * nw + warp;
    if (task < n_tasks && lane == 0) issue(task, 0);
    for (int it = 0; task < n_tasks; ++it, task += n_warps) {
        const int buf = it & 1;
        if (task + n_warps < n_tasks && lane == 0) issue(task + n_warps, buf ^ 1);
        const int b = (int)(task / bps), blk = (int)(task % bps);
        const float wb = wseq ? wseq[b] : 1.f;                 // per-sequence weight (autograd: d loss / d loglik_b)
        const int n = min(T - 1, blk * XI_FPW + XI_FPW) - blk * XI_FPW;
        bw_mbar_wait(bar + buf, (uint32_t)((it >> 1) & 1));
        const float *es = stg + (size_t)buf * 3 * BUF, *bs = es + BUF, *as = es + 2 * BUF;
        for (int tb = 0; tb < n; tb += FPI) {
            const bool live = tb + fr < n;
            const int row = live ? tb + fr : 0;
            float u[KP], bb[KP];
#pragma unroll
            for (int q = 0; q < KV; ++q) {
                const float4 e4 = reinterpret_cast<const float4 *>(es + row * KP)[q], b4 = reinterpret_cast<const float4 *>(bs + row * KP)[q];
                u[4 * q] = e4.x; u[4 * q + 1] = e4.y; u[4 * q + 2] = e4.z; u[4 * q + 3] = e4.w;
                bb[4 * q] = b4.x; bb[4 * q + 1] = b4.y; bb[4 * q + 2] = b4.z; bb[4 * q + 3] = b4.w;
            }
            float a[R];
#pragma unroll
            for (int r = 0; r < R; ++r) { const int i = part * R + r; a[r] = (i < K) ? as[row * KP + i] : 0.f; }
            float mx = -INFINITY;
#pragma unroll
            for (int j = 0; j < KP; ++j) mx = fmaxf(mx, u[j]);
            if (!(mx > -INFINITY)) mx = 0.f;
#pragma unroll
            for (int j = 0; j < KP; ++j) {
                float bt;
                if (mode == HMMB200_EMIS_PROB_FLOOR) bt = u[j] + eps;
                else if (mode == HMMB200_EMIS_LOG_EXP_FLOOR) bt = expf(u[j]) + eps;
                else bt = expf(u[j] - mx) + ((mode == HMMB200_EMIS_LOG_NORM_FLOOR) ? eps : 0.f);
                u[j] = bt * bb[j];
            }
            float pu[R][KP], z = 0.f;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                float rs = 0.f;
#pragma unroll
                for (int j = 0; j < KP; ++j) { pu[r][j] = Prow[r][j] * u[j]; rs += pu[r][j]; }
                z = fmaf(a[r], rs, z);
            }
#pragma unroll
            for (int o = LPF / 2; o > 0; o >>= 1) z += __shfl_xor_sync(FULL_MASK, z, o);
            const float inv = (live && z > 0.f) ? wb / z : 0.f;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const float sc = a[r] * inv;
#pragma unroll
                for (int j = 0; j < KP; ++j) acc[r][j] = fmaf(sc, pu[r][j], acc[r][j]);
            }
        }
        if (blk == 0 && gamma1 != nullptr) {               // gamma_0 = a_0 .* b_0 / sum
            const bool ok = lane < K;
            const float g = ok ? ws_a[(size_t)b * T * K + lane] * ws_b[(size_t)b * T * K + lane] : 0.f;
            float Z = g;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) Z += __shfl_xor_sync(FULL_MASK, Z, o);
            if (ok && Z > 0.f) atomicAdd(g1_s + lane, (double)(wb * g / Z));
        }
        __syncwarp();                                          // every lane is done with this buffer before it is refilled
    }
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int j = 0; j < KP; ++j) {
            float v = acc[r][j];
#pragma unroll
            for (int o = LPF; o < 32; o <<= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
            const int i = part * R + r;
            if (fr == 0 && i < K && v != 0.f) atomicAdd(&xi_s[i * K + j], (double)v);
        }
    __syncthreads();
    for (int i = threadIdx.x; i < K * K; i += blockDim.x) if (xi_s[i] != 0.0) atomicAdd(xi + i, xi_s[i]);
    if (gamma1 != nullptr)
        for (int i = threadIdx.x; i < K; i += blockDim.x) if (g1_s[i] != 0.0) atomicAdd(gamma1 + i, g1_s[i]);
}

template <int KP, int LPF>
static int launch_xi_kp(const float *emis, int mode, float eps, const float *trans, const float *ws_a, const float *ws_b,
                        const float *wseq, int B, int T, int K, double *xi, double *gamma1, cudaStream_t s) {
    const int fpw = 64, warps = 12;
    int dev0 = 0, sms0 = 148;
    cudaGetDevice(&dev0);
    cudaDeviceGetAttribute(&sms0, cudaDevAttrMultiProcessorCount, dev0);
    bool bulk = (K == KP) && T > 1 && ((((uintptr_t)emis | (uintptr_t)ws_a | (uintptr_t)ws_b) & 15) == 0);
#ifdef HMMB200_DEBUG_HOOKS
    if (getenv("HMMB200_XI_NO_BULK")) bulk = false;
#endif
    if (bulk) {
        const size_t smem = (((size_t)warps * 2 * sizeof(uint64_t) + 15) & ~(size_t)15) + (((size_t)(K * K + K) * sizeof(double) + 15) & ~(size_t)15) +
                            (size_t)warps * 2 * 3 * XI_FPW * KP * sizeof(float);
        if (smem <= 200 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(bw_xi_bulk_kernel<KP, LPF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return set_error(HMMB200_ELAUNCH, "bw_xi smem opt-in: %s", cudaGetErrorString(e));
            const int64_t nt = (int64_t)B * ((T - 1 + XI_FPW - 1) / XI_FPW);
            const int64_t nc = min((nt + warps - 1) / warps, (int64_t)sms0);
            bw_xi_bulk_kernel<KP, LPF><<<(unsigned)nc, warps * 32, smem, s>>>(emis, mode, eps, trans, ws_a, ws_b, wseq, B, T, xi, gamma1);
            return check_launch("bw_xi_bulk_kernel");
        }
    }
    const int blocks_per_seq = T > 1 ? (T - 1 + fpw - 1) / fpw : 1;
    const int64_t n_tasks = (int64_t)B * blocks_per_seq;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t ctas = min((n_tasks + warps - 1) / warps, (int64_t)sms);
    bw_xi_kernel<KP, LPF><<<(unsigned)ctas, warps * 32, (size_t)(K * K + K) * sizeof(double), s>>>(
        emis, mode, eps, trans, ws_a, ws_b, wseq, B, T, K, fpw, xi, gamma1);
    return check_launch("bw_xi_kernel");
}

static int launch_xi(const float *emis, int mode, float eps, const float *trans, const float *ws_a, const float *ws_b,
                     const float *wseq, int B, int T, int K, double *xi, double *gamma1, cudaStream_t s) {
    const int kp = pad4(K);
    // lanes per frame: the smallest power of two that leaves a lane at most ~48 cells of xi
#define XI_CASE(N, L) if (kp <= N) return launch_xi_kp<N, L>(emis, mode, eps, trans, ws_a, ws_b, wseq, B, T, K, xi, gamma1, s)
    XI_CASE(4, 1); XI_CASE(8, 2); XI_CASE(12, 4); XI_CASE(16, 8); XI_CASE(20, 16); XI_CASE(24, 16); XI_CASE(28, 32);
#undef XI_CASE
    return launch_xi_kp<32, 32>(emis, mode, eps, trans, ws_a, ws_b, wseq, B, T, K, xi, gamma1, s);
}

static int launch_gmm_stats_generic(const float *x, const float *comp, const float *logb, const float *gamma, int64_t n, int K, int C, int D,
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

static int launch_gmm_stats(const float *x, const float *comp, const float *logb, const float *gamma, int64_t n, int K, int C, int D,
                            double *occ, double *sx, double *sxx, cudaStream_t s) {
    const int KC = K * C;
    auto al16 = [](const void *q) { return ((uintptr_t)q & 15) == 0; };
    const int cells = (KC / BW_TC) * (D / BW_TD);
    const size_t stage = (size_t)BW_F * (D + KC + 2 * K) * sizeof(float);
    const size_t smem = 64 + BW_STAGES * stage + (size_t)BW_F * KC * sizeof(float);
    const bool bulk = KC % BW_TC == 0 && D % BW_TD == 0 && cells >= 1 && cells <= 512 && n >= BW_F && smem <= 200 * 1024 &&
                      BW_STAGES * stage >= (size_t)cells * (2 * BW_TC * BW_TD + BW_TC) * sizeof(float) &&
                      al16(x) && al16(comp) && al16(logb) && al16(gamma);
    if (!bulk) return launch_gmm_stats_generic(x, comp, logb, gamma, n, K, C, D, occ, sx, sxx, s);
    static bool attr_done[64];
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (dev < 0 || dev >= 64 || !attr_done[dev]) {
        cudaError_t e = cudaFuncSetAttribute(bw_gmm_stats_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return set_error(HMMB200_ELAUNCH, "gmm_stats smem opt-in: %s", cudaGetErrorString(e));
        if (dev >= 0 && dev < 64) attr_done[dev] = true;
    }
    int FG = 512 / cells;
    FG = FG < 1 ? 1 : (FG > 8 ? 8 : FG);
    const int threads = (cells * FG + 31) & ~31;
    const int64_t n_full = n / BW_F;
    // tensor-core form when the register tiling covers the shape: K*C <= 64 rows (MT <= 4), D <= 96 (NTW <= 6)
    const int MT = (KC + 15) / 16, NTW = (D / 8 + 1) / 2;
    const size_t smem_mma = 64 + BW_STAGES * stage + (size_t)BW_F * (MT * 16 + 8) * sizeof(float);
    const bool mma_ok = MT <= 4 && NTW <= 6 && smem_mma <= 200 * 1024 && BW_STAGES * stage >= (size_t)MT * 16 * (2 * D + 1) * sizeof(float);
    bool launched = false;
    int path = 0;                                                   // 0 best available, 1 mma.sync, 2 CUDA-core FMAs (debug builds: A/B timing)
#ifdef HMMB200_DEBUG_HOOKS
    if (const char *e = getenv("HMMB200_STATS_PATH")) path = atoi(e);
#endif
    if (path == 0) {                                                // tcgen05 form (bw_tc.cu)
        const int rc = launch_gmm_stats_tc(x, comp, logb, gamma, n_full * BW_F, K, C, D, occ, sx, sxx, s);
        if (rc < 0) return rc;
        launched = (rc == 0);
    }
    if (!launched && mma_ok && path <= 1) {
#define BW_MMA_CASE(M_, N_)                                                                                                            \
        if (!launched && MT == M_ && NTW == N_) {                                                                                       \
            cudaFuncSetAttribute(bw_gmm_stats_mma_kernel<M_, N_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);             \
            bw_gmm_stats_mma_kernel<M_, N_><<<(unsigned)min((int64_t)sms, n_full), 512, smem_mma, s>>>(x, comp, logb, gamma, n_full * BW_F, \
                                                                                                    K, C, D, occ, sx, sxx);           \
            launched = true;                                                                                                            \
        }
        BW_MMA_CASE(1, 1) BW_MMA_CASE(1, 2) BW_MMA_CASE(1, 3) BW_MMA_CASE(1, 4) BW_MMA_CASE(1, 5) BW_MMA_CASE(1, 6)
        BW_MMA_CASE(2, 1) BW_MMA_CASE(2, 2) BW_MMA_CASE(2, 3) BW_MMA_CASE(2, 4) BW_MMA_CASE(2, 5) BW_MMA_CASE(2, 6)
        BW_MMA_CASE(3, 1) BW_MMA_CASE(3, 2) BW_MMA_CASE(3, 3) BW_MMA_CASE(3, 4) BW_MMA_CASE(3, 5) BW_MMA_CASE(3, 6)
        BW_MMA_CASE(4, 1) BW_MMA_CASE(4, 2) BW_MMA_CASE(4, 3) BW_MMA_CASE(4, 4) BW_MMA_CASE(4, 5) BW_MMA_CASE(4, 6)
#undef BW_MMA_CASE
    }
    if (!launched)
        bw_gmm_stats_bulk_kernel<<<(unsigned)min((int64_t)sms, n_full), threads, smem, s>>>(x, comp, logb, gamma, n_full * BW_F, K, C, D, FG,
                                                                                          occ, sx, sxx);
    if (int rc = check_launch("bw_gmm_stats_bulk_kernel")) return rc;
    const int64_t tail = n - n_full * BW_F;                         // fewer than BW_F frames: the plain kernel
    if (tail > 0) {
        const int64_t o = n_full * BW_F;
        return launch_gmm_stats_generic(x + o * D, comp + o * KC, logb + o * K, gamma + o * K, tail, K, C, D, occ, sx, sxx, s);
    }
    return HMMB200_OK;
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
