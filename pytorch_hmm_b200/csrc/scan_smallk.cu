// scan_smallk.cu -- time-parallel forward-backward for long sequences at small batch (K <= 32), sm_100a.
//
// The recursions of pytorch_hmm/hmm.py:95-117 are T dependent steps; with a handful of sequences the warp-per-sequence
// kernels (recursion_smallk.cu) leave the chip idle and take T x ~65 ns.  The step alpha_t^T = alpha_{t-1}^T M_t with
// M_t = P diag(b_t) is a product of K x K matrices, i.e. associative, so time is cut into S segments of L frames:
//   1. scan_products_kernel    every (sequence, segment, row i) thread pushes the unit vector e_i through its segment:
//                              the rows of Pi_s = M_t0 ... M_t1 (K x the sequential work, but S*K-way parallel);
//   2. scan_boundaries_kernel  one warp per sequence chains the S segment products: forward vectors at the segment starts
//                              (alpha^T Pi_s), backward vectors at the segment ends (Pi_s beta) -- the SAME products serve
//                              both directions -- and the log-likelihood;
//   3. scan_fill_kernel        every (sequence, segment) thread redoes its L frames from the true boundary vectors, forward
//                              then backward, and writes posterior / exp(log alpha) / exp(log beta) (hmm.py:120-128).
// Everything is scaled-probability fp32 with exact power-of-two renormalisation per step (integer exponents), the
// boundary chaining in double.  Emissions are converted once (scan_emissions_kernel) to the probability form b~ and the
// per-frame log scale m of the emission mode, exactly as the sequential kernels' loader warps do.
// A max-plus (Viterbi) scan is deliberately NOT offered: re-associating fp32 additions changes delta in the last bits,
// and the contract for Viterbi is bit-exactness (BASELINE.json north_star).
#include "common.cuh"

#include <limits.h>

namespace hmmb200 {

struct ScanParams {
    const float *emis;
    int mode;
    float eps;
    int add_m;
    const float *trans, *init;        // [K,K], [K] effective probabilities
    int B, T, K, L, S;
    float *bt;                        // [B,T,K]  b~
    float *mrow;                      // [B,T]    m
    float *prod;                      // [B,S,K,K] rows of the normalised segment products
    int *pexp;                        // [B,S,K]  their power-of-two exponents
    double *msum;                     // [B,S]    sum of m over the segment's frames
    double *bndA;                     // [B,S+1,K] forward vector at the frame before each segment (common scale exp(bndLA))
    double *bndLA;                    // [B,S+1]   its log scale
    double *bndB;                     // [B,S,K]   backward vector at each segment's last frame (common scale exp(bndLB))
    double *bndLB;                    // [B,S]
    float *ws_a;                      // [B,T,K]  scaled alpha (phase 3 scratch)
    float *ws_la;                     // [B,T]
    float *gamma, *fwd, *bwd, *log_alpha, *log_beta, *loglik;
    int vec_out;                      // every output pointer is 16-byte aligned
};

// segment s covers frames t0..t1 (frame 0 is the start vector p0 .* b_0 and belongs to no product)
__device__ __forceinline__ void seg_range(int s, int L, int T, int &t0, int &t1) {
    t0 = 1 + s * L;
    t1 = min(T - 1, (s + 1) * L);
}

// exact power-of-two renormalisation: v *= 2^-e with e = exponent of the largest entry; returns e (0 for an all-zero vector)
template <int KP>
__device__ __forceinline__ int renorm(float (&v)[KP]) {
    float mx = 0.f;
#pragma unroll
    for (int k = 0; k < KP; ++k) mx = fmaxf(mx, v[k]);
    const unsigned eb = __float_as_uint(mx) >> 23;
    if (eb == 0u || eb >= 255u) return 0;
    const float r = __uint_as_float((254u - eb) << 23);
#pragma unroll
    for (int k = 0; k < KP; ++k) v[k] *= r;
    return (int)eb - 127;
}

__global__ void __launch_bounds__(256) scan_emissions_kernel(ScanParams p) {
    const int64_t fr = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (fr >= (int64_t)p.B * p.T) return;
    const int K = p.K;
    const float *e = p.emis + fr * K;
    float mx = 0.f;
    if (p.mode == HMMB200_EMIS_LOG || p.mode == HMMB200_EMIS_LOG_NORM_FLOOR) {
        mx = -INFINITY;
        for (int k = 0; k < K; ++k) mx = fmaxf(mx, e[k]);
        if (!(mx > -INFINITY)) mx = 0.f;
    }
    for (int k = 0; k < K; ++k) {
        float v;
        if (p.mode == HMMB200_EMIS_PROB_FLOOR) v = e[k] + p.eps;
        else if (p.mode == HMMB200_EMIS_LOG_EXP_FLOOR) v = expf(e[k]) + p.eps;
        else v = expf(e[k] - mx) + ((p.mode == HMMB200_EMIS_LOG_NORM_FLOOR) ? p.eps : 0.f);
        p.bt[fr * K + k] = v;
    }
    p.mrow[fr] = mx;
}

// shared-memory copy of P, row-major [K][KP] (rows padded with zeros): every thread reads the same element -> broadcast
template <int KP>
__device__ __forceinline__ void load_P(const float *trans, int K, float *P_s) {
    for (int i = threadIdx.x; i < KP * KP; i += blockDim.x) {     // rows AND columns beyond K are zero: no bounds tests later
        const int r = i / KP, c = i % KP;
        P_s[i] = (r < K && c < K) ? trans[r * K + c] : 0.f;
    }
    __syncthreads();
}

// Rows of P come from shared memory as broadcast LDS.128; they are fetched PD rows ahead of their use so that the ~30-cycle
// shared-memory latency overlaps the FMAs of the rows in hand (one warp per scheduler: nothing else would hide it).
template <int KP>
struct PRows {
    static constexpr int Q = KP / 4, PD = 3;
    float4 r[PD][Q];
    __device__ __forceinline__ void fetch(const float *P_s, int i) {
#pragma unroll
        for (int q = 0; q < Q; ++q) r[i % PD][q] = reinterpret_cast<const float4 *>(P_s + i * KP)[q];
    }
};

// out[j] = sum_i v[i] P[i][j]
template <int KP>
__device__ __forceinline__ void vec_mat(const float (&v)[KP], const float *P_s, int K, float (&out)[KP]) {
    PRows<KP> pr;
#pragma unroll
    for (int j = 0; j < KP; ++j) out[j] = 0.f;
#pragma unroll
    for (int i = 0; i < PRows<KP>::PD - 1; ++i) if (i < KP) pr.fetch(P_s, i);
#pragma unroll
    for (int i = 0; i < KP; ++i) {                          // straight-line code (padded rows of P are zero): no branches
        if (i + PRows<KP>::PD - 1 < KP) pr.fetch(P_s, i + PRows<KP>::PD - 1);
#pragma unroll
        for (int q = 0; q < KP / 4; ++q) {
            const float4 r = pr.r[i % PRows<KP>::PD][q];
            out[4 * q + 0] = fmaf(v[i], r.x, out[4 * q + 0]);
            out[4 * q + 1] = fmaf(v[i], r.y, out[4 * q + 1]);
            out[4 * q + 2] = fmaf(v[i], r.z, out[4 * q + 2]);
            out[4 * q + 3] = fmaf(v[i], r.w, out[4 * q + 3]);
        }
    }
    (void)K;
}

// out[i] = sum_j P[i][j] u[j]
template <int KP>
__device__ __forceinline__ void mat_vec(const float (&u)[KP], const float *P_s, int K, float (&out)[KP]) {
    PRows<KP> pr;
#pragma unroll
    for (int i = 0; i < PRows<KP>::PD - 1; ++i) if (i < KP) pr.fetch(P_s, i);
#pragma unroll
    for (int i = 0; i < KP; ++i) {
        float a0 = 0.f, a1 = 0.f;
        if (i + PRows<KP>::PD - 1 < KP) pr.fetch(P_s, i + PRows<KP>::PD - 1);
#pragma unroll
        for (int q = 0; q < KP / 4; ++q) {
            const float4 r = pr.r[i % PRows<KP>::PD][q];
            a0 = fmaf(r.x, u[4 * q + 0], a0);
            a1 = fmaf(r.y, u[4 * q + 1], a1);
            a0 = fmaf(r.z, u[4 * q + 2], a0);
            a1 = fmaf(r.w, u[4 * q + 3], a1);
        }
        out[i] = a0 + a1;
    }
    (void)K;
}

template <int KP>
__device__ __forceinline__ void load_row(const float *src, int K, float (&v)[KP]) {
    if (K == KP) {                                          // K % 4 == 0: rows are 16-byte aligned, KP/4 vector loads
#pragma unroll
        for (int q = 0; q < KP / 4; ++q) {
            const float4 x = __ldg(reinterpret_cast<const float4 *>(src) + q);
            v[4 * q] = x.x; v[4 * q + 1] = x.y; v[4 * q + 2] = x.z; v[4 * q + 3] = x.w;
        }
    } else {
#pragma unroll
        for (int k = 0; k < KP; ++k) v[k] = (k < K) ? __ldg(src + k) : 0.f;
    }
}

template <int KP>
__device__ __forceinline__ void store_row(float *dst, int K, bool vec, const float (&v)[KP]) {
    if (vec) {
#pragma unroll
        for (int q = 0; q < KP / 4; ++q) reinterpret_cast<float4 *>(dst)[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    } else {
#pragma unroll
        for (int k = 0; k < KP; ++k) if (k < K) dst[k] = v[k];
    }
}

// ---- phase 1: rows of the segment products --------------------------------------------------------------------------
template <int KP>
__global__ void __launch_bounds__(128) scan_products_kernel(ScanParams p) {
    __shared__ __align__(16) float P_s[32 * KP];
    const int K = p.K, T = p.T, L = p.L, S = p.S;
    load_P<KP>(p.trans, K, P_s);
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (int64_t)p.B * S * K) return;
    const int i = (int)(gid % K);
    const int s = (int)((gid / K) % S);
    const int b = (int)(gid / ((int64_t)K * S));
    int t0, t1;
    seg_range(s, L, T, t0, t1);
    float v[KP], w[KP];
#pragma unroll
    for (int k = 0; k < KP; ++k) v[k] = (k == i) ? 1.f : 0.f;
    int kexp = 0;
    double ms = 0.0;
    const float *bt = p.bt + ((size_t)b * T) * K;
    // frames are consumed in groups of GS whose rows are fetched one group ahead (register double buffer): with ~1 warp per
    // scheduler nothing else hides the L2 latency of the emission rows
    constexpr int GS = (KP <= 16) ? 4 : 2;
    float cur[GS][KP], nxt[GS][KP];
    auto fetch = [&](int tb, float (&dst)[GS][KP]) {
#pragma unroll
        for (int g = 0; g < GS; ++g) {
            const int t = min(tb + g, t1);
            load_row<KP>(bt + (size_t)t * K, K, dst[g]);
        }
    };
    fetch(t0, cur);
    for (int tb = t0; tb <= t1; tb += GS) {
        if (tb + GS <= t1) fetch(tb + GS, nxt);
#pragma unroll
        for (int g = 0; g < GS; ++g) {
            if (tb + g <= t1) {
                vec_mat<KP>(v, P_s, K, w);
#pragma unroll
                for (int k = 0; k < KP; ++k) v[k] = w[k] * cur[g][k];
                kexp += renorm<KP>(v);
                if (i == 0 && p.add_m) ms += (double)__ldg(p.mrow + (size_t)b * T + tb + g);
            }
        }
#pragma unroll
        for (int g = 0; g < GS; ++g)
#pragma unroll
            for (int k = 0; k < KP; ++k) cur[g][k] = nxt[g][k];
    }
    float *dst = p.prod + (((size_t)b * S + s) * K + i) * K;
#pragma unroll
    for (int k = 0; k < KP; ++k) if (k < K) dst[k] = v[k];
    p.pexp[((size_t)b * S + s) * K + i] = kexp;
    if (i == 0) p.msum[(size_t)b * S + s] = ms;
}

// ---- phase 2: chain the segment products --------------------------------------------------------------------------------
// One CTA of two warps per sequence: warp 0 chains forward (alpha^T Pi_s), warp 1 backward (Pi_s beta), lane = state.
// The products are streamed through shared memory in double-buffered batches (cp.async), so the serial chain never
// waits for L2; renormalisation is an exact power of two (integer exponent), logs are taken only when a scale is stored.
__device__ __forceinline__ void sc_cp4(void *dst_smem, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ double pow2d(int e) { return __longlong_as_double((long long)(1023 + e) << 52); }   // |e| < 1022
__device__ __forceinline__ int expo_d(double x) { return (int)((__double_as_longlong(x) >> 52) & 0x7ff) - 1023; }
__device__ __forceinline__ double wmax_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(FULL_MASK, v, o));
    return v;
}
__device__ __forceinline__ int wmax_i(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(FULL_MASK, v, o));
    return v;
}

// Boundary vectors are kept as doubles WITHOUT per-step normalisation (every row of a segment product has its own power-of-two
// exponent; taking them relative to row 0's keeps everything within double range), so a step is K broadcast shuffles + K FMAs
// and two warp votes; the exponent is pulled out only when the vector drifts past 2^+-300.  Phase 3 normalises what it reads.
template <int KP>
__global__ void __launch_bounds__(64) scan_boundaries_kernel(ScanParams p, int nbat) {
    extern __shared__ __align__(16) uint8_t sc_smem[];
    const int K = p.K, T = p.T, S = p.S;
    const int b = blockIdx.x, warp = threadIdx.x >> 5, j = threadIdx.x & 31;
    const bool ok = j < K;
    const int KK = K * K;
    // per warp: 2 buffers x nbat x (1 double + K*K floats + K ints)
    const size_t buf_bytes = (size_t)nbat * (KK * 4 + K * 4 + 8);
    uint8_t *mybuf = sc_smem + (size_t)warp * 2 * buf_bytes;
    auto buf_prod = [&](int q) { return reinterpret_cast<float *>(mybuf + (size_t)q * buf_bytes + (size_t)nbat * 8); };
    auto buf_exp = [&](int q) { return reinterpret_cast<int *>(mybuf + (size_t)q * buf_bytes + (size_t)nbat * 8 + (size_t)nbat * KK * 4); };
    auto buf_ms = [&](int q) { return reinterpret_cast<double *>(mybuf + (size_t)q * buf_bytes); };
    const float *prod = p.prod + (size_t)b * S * KK;
    const int *pexp = p.pexp + (size_t)b * S * K;
    const double *msum = p.msum + (size_t)b * S;
    // segment of slot r of batch q: forward ascending from 0, backward descending from S-1 (down to 1)
    const int n_steps = (warp == 0) ? S : max(S - 1, 0);
    const int n_batches = (n_steps + nbat - 1) / nbat;
    auto seg_of = [&](int q, int r) { return (warp == 0) ? q * nbat + r : S - 1 - (q * nbat + r); };
    auto issue = [&](int q) {
        const int cnt = min(nbat, n_steps - q * nbat);
        float *bp = buf_prod(q & 1);
        int *be = buf_exp(q & 1);
        double *bm = buf_ms(q & 1);
        for (int r = 0; r < cnt; ++r) {
            const int sg = seg_of(q, r);
            for (int x = j; x < KK; x += 32) sc_cp4(bp + r * KK + x, prod + (size_t)sg * KK + x);
            if (ok) sc_cp4(be + r * K + j, pexp + (size_t)sg * K + j);
            if (j < 2) sc_cp4(reinterpret_cast<float *>(bm + r) + j, reinterpret_cast<const float *>(msum + sg) + j);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    // pull a power of two out of the vector when it drifts (rare): exponent of the largest entry
    auto settle = [&](double &v, long long &ksum) {
        const bool big = ok && v > 0x1p300, small = !ok || v < 0x1p-300;
        if (__any_sync(FULL_MASK, big) || __all_sync(FULL_MASK, small)) {
            const int e = wmax_i((ok && v > 0.0) ? expo_d(v) : INT_MIN);
            if (e != INT_MIN) { v *= pow2d(-e); ksum += e; }
        }
    };
    auto clampd = [](int d) { return d < -900 ? -900 : (d > 900 ? 900 : d); };

    if (warp == 0) {
        // ---------------- forward: a_0 = p0 .* b~_0;  a_{s+1}[j] = sum_i a_s[i] 2^e_i Pi^_s[i][j] ----------------
        double a = ok ? (double)p.init[j] * (double)p.bt[((size_t)b * T) * K + j] : 0.0;
        double macc = p.add_m ? (double)p.mrow[(size_t)b * T] : 0.0;
        long long ksum = 0;
        double *outA = p.bndA + (size_t)b * (S + 1) * K;
        double *outL = p.bndLA + (size_t)b * (S + 1);
        if (ok) outA[j] = a;
        if (j == 0) outL[0] = macc;
        if (n_batches > 0) issue(0);
        for (int q = 0; q < n_batches; ++q) {
            if (q + 1 < n_batches) { issue(q + 1); asm volatile("cp.async.wait_group 1;" ::: "memory"); }
            else asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncwarp();
            const int cnt = min(nbat, n_steps - q * nbat);
            const float *bp = buf_prod(q & 1);
            const int *be = buf_exp(q & 1);
            const double *bm = buf_ms(q & 1);
            for (int r = 0; r < cnt; ++r) {
                const int sg = q * nbat + r;
                float col[KP];                               // column j of the product, fetched before the dependent chain
#pragma unroll
                for (int i = 0; i < KP; ++i) col[i] = (i < K && ok) ? bp[r * KK + i * K + j] : 0.f;
                const int e0 = be[r * K];
                const double wj = ok ? a * pow2d(clampd(be[r * K + j] - e0)) : 0.0;
                double an0 = 0.0, an1 = 0.0;
#pragma unroll
                for (int i = 0; i < KP; i += 2) {
                    if (i < K) an0 = fma(__shfl_sync(FULL_MASK, wj, i), (double)col[i], an0);
                    if (i + 1 < K) an1 = fma(__shfl_sync(FULL_MASK, wj, i + 1), (double)col[i + 1], an1);
                }
                a = an0 + an1;
                ksum += e0;
                settle(a, ksum);
                macc += bm[r];
                if (ok) outA[(size_t)(sg + 1) * K + j] = a;
                if (j == 0) outL[sg + 1] = macc + 0.69314718055994530942 * (double)ksum;
            }
            __syncwarp();
        }
        if (p.loglik) {
            double tot = ok ? a : 0.0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(FULL_MASK, tot, o);
            if (j == 0) p.loglik[b] = (float)(macc + 0.69314718055994530942 * (double)ksum + log(tot));
        }
    } else {
        // ---------------- backward: beta(last frame) = 1;  beta(end of s-1)[i] = 2^e_i sum_c Pi^_s[i][c] beta(end of s)[c] ----------------
        double bb = ok ? 1.0 : 0.0, macc = 0.0;
        long long ksum = 0;
        double *outB = p.bndB + (size_t)b * S * K;
        double *outL = p.bndLB + (size_t)b * S;
        if (S > 0) {
            if (ok) outB[(size_t)(S - 1) * K + j] = 1.0;
            if (j == 0) outL[S - 1] = 0.0;
        }
        if (n_batches > 0) issue(0);
        for (int q = 0; q < n_batches; ++q) {
            if (q + 1 < n_batches) { issue(q + 1); asm volatile("cp.async.wait_group 1;" ::: "memory"); }
            else asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncwarp();
            const int cnt = min(nbat, n_steps - q * nbat);
            const float *bp = buf_prod(q & 1);
            const int *be = buf_exp(q & 1);
            const double *bm = buf_ms(q & 1);
            for (int r = 0; r < cnt; ++r) {
                const int sg = S - 1 - (q * nbat + r);                // >= 1
                float row[KP];                               // row j of the product
#pragma unroll
                for (int c = 0; c < KP; ++c) row[c] = (c < K && ok) ? bp[r * KK + j * K + c] : 0.f;
                const int e0 = be[r * K];
                const double sc = ok ? pow2d(clampd(be[r * K + j] - e0)) : 0.0;
                double d0 = 0.0, d1 = 0.0;
#pragma unroll
                for (int c = 0; c < KP; c += 2) {
                    if (c < K) d0 = fma((double)row[c], __shfl_sync(FULL_MASK, bb, c), d0);
                    if (c + 1 < K) d1 = fma((double)row[c + 1], __shfl_sync(FULL_MASK, bb, c + 1), d1);
                }
                bb = (d0 + d1) * sc;
                ksum += e0;
                settle(bb, ksum);
                macc += bm[r];
                if (ok) outB[(size_t)(sg - 1) * K + j] = bb;
                if (j == 0) outL[sg - 1] = macc + 0.69314718055994530942 * (double)ksum;
            }
            __syncwarp();
        }
    }
}

// normalise a boundary vector (doubles of arbitrary common scale) into floats with largest entry in [1,2); returns the log scale
template <int KP>
__device__ __forceinline__ double load_boundary(const double *src, int K, float (&v)[KP]) {
    double d[KP], mx = 0.0;
#pragma unroll
    for (int k = 0; k < KP; ++k) { d[k] = (k < K) ? src[k] : 0.0; mx = fmax(mx, d[k]); }
    int e = 0;
    if (mx > 0.0) e = expo_d(mx);
    const double sc = pow2d(-e);
#pragma unroll
    for (int k = 0; k < KP; ++k) v[k] = (float)(d[k] * sc);
    return 0.69314718055994530942 * (double)e;
}

// ---- phase 3: fill every segment from its true boundary vectors ------------------------------------------------------
template <int KP>
__global__ void __launch_bounds__(64) scan_fill_kernel(ScanParams p) {
    __shared__ __align__(16) float P_s[32 * KP];
    const int K = p.K, T = p.T, L = p.L, S = p.S;
    load_P<KP>(p.trans, K, P_s);
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int n_seg = max(S, 1);                            // T == 1: one pseudo segment that only handles frame 0
    if (gid >= (int64_t)p.B * n_seg) return;
    const int s = (int)(gid % n_seg), b = (int)(gid / n_seg);
    int t0, t1;
    seg_range(s, L, T, t0, t1);
    const float *bt = p.bt + ((size_t)b * T) * K;
    const float *mr = p.mrow + (size_t)b * T;
    float *wa = p.ws_a + ((size_t)b * T) * K;
    float *wla = p.ws_la + (size_t)b * T;
    constexpr int GS = (KP <= 12) ? 4 : ((KP <= 16) ? 2 : 1);   // frames per prefetch group (see scan_products_kernel)
    float v[KP], w[KP];
    // ---- forward ----
    const double la = p.bndLA[(size_t)b * (S + 1) + s] + load_boundary<KP>(p.bndA + ((size_t)b * (S + 1) + s) * K, K, v);
    if (s == 0) {                                           // frame 0 itself
        store_row<KP>(wa, K, K == KP, v);
        wla[0] = (float)la;
    }
    if (S > 0) {
        int ksum = 0;
        double msum = 0.0;
        float cur[GS][KP], nxt[GS][KP], mc[GS], mn[GS];
        auto fetch = [&](int tb, float (&dst)[GS][KP], float (&dm)[GS]) {
#pragma unroll
            for (int g = 0; g < GS; ++g) {
                const int t = min(tb + g, t1);
                load_row<KP>(bt + (size_t)t * K, K, dst[g]);
                dm[g] = p.add_m ? __ldg(mr + t) : 0.f;
            }
        };
        fetch(t0, cur, mc);
        for (int tb = t0; tb <= t1; tb += GS) {
            if (tb + GS <= t1) fetch(tb + GS, nxt, mn);
#pragma unroll
            for (int g = 0; g < GS; ++g) {
                const int t = tb + g;
                if (t <= t1) {
                    vec_mat<KP>(v, P_s, K, w);
#pragma unroll
                    for (int k = 0; k < KP; ++k) v[k] = w[k] * cur[g][k];
                    ksum += renorm<KP>(v);
                    msum += (double)mc[g];
                    store_row<KP>(wa + (size_t)t * K, K, K == KP, v);
                    wla[t] = (float)(la + msum + 0.69314718055994530942 * (double)ksum);
                }
            }
#pragma unroll
            for (int g = 0; g < GS; ++g) {
                mc[g] = mn[g];
#pragma unroll
                for (int k = 0; k < KP; ++k) cur[g][k] = nxt[g][k];
            }
        }
    }
    // ---- backward + outputs ----
    double lb;
    if (S > 0) {
        lb = p.bndLB[(size_t)b * S + s] + load_boundary<KP>(p.bndB + ((size_t)b * S + s) * K, K, w);
    } else {
#pragma unroll
        for (int k = 0; k < KP; ++k) w[k] = (k < K) ? 1.f : 0.f;
        lb = 0.0;
        t1 = 0;
    }
    int kb = 0;
    double mb = 0.0;
    const int t_stop = (s == 0) ? 0 : t0;
    {
        float ca[GS][KP], cb[GS][KP], na[GS][KP], nb[GS][KP], cl[GS], cm[GS], nl[GS], nm[GS];
        auto fetch = [&](int tb, float (&da)[GS][KP], float (&db)[GS][KP], float (&dl)[GS], float (&dm)[GS]) {
#pragma unroll
            for (int g = 0; g < GS; ++g) {
                const int t = max(tb - g, t_stop);
                if (K == KP) {
#pragma unroll
                    for (int q = 0; q < KP / 4; ++q) {
                        const float4 x = reinterpret_cast<const float4 *>(wa + (size_t)t * K)[q];
                        da[g][4 * q] = x.x; da[g][4 * q + 1] = x.y; da[g][4 * q + 2] = x.z; da[g][4 * q + 3] = x.w;
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < KP; ++k) da[g][k] = (k < K) ? wa[(size_t)t * K + k] : 0.f;
                }
                load_row<KP>(bt + (size_t)t * K, K, db[g]);
                dl[g] = wla[t];
                dm[g] = p.add_m ? __ldg(mr + t) : 0.f;
            }
        };
        fetch(t1, ca, cb, cl, cm);
        for (int tb = t1; tb >= t_stop; tb -= GS) {
            if (tb - GS >= t_stop) fetch(tb - GS, na, nb, nl, nm);
#pragma unroll
            for (int g = 0; g < GS; ++g) {
                const int t = tb - g;
                if (t >= t_stop) {
                    // w = beta_t (scaled), true beta_t = w * exp(lbt)
                    const double lbt = lb + mb + 0.69314718055994530942 * (double)kb;
                    const float lat = cl[g];
                    float z = 0.f;
#pragma unroll
                    for (int k = 0; k < KP; ++k) z = fmaf(ca[g][k], w[k], z);
                    const float inv = 1.f / z, ea = expf(lat), eb = expf((float)lbt);
                    const size_t o = ((size_t)b * T + t) * K;
                    const bool vec = (K == KP) && p.vec_out;
                    float r[KP];
                    if (p.gamma) {
#pragma unroll
                        for (int k = 0; k < KP; ++k) r[k] = ca[g][k] * w[k] * inv;
                        store_row<KP>(p.gamma + o, K, vec, r);
                    }
                    if (p.fwd) {
#pragma unroll
                        for (int k = 0; k < KP; ++k) r[k] = ca[g][k] * ea;
                        store_row<KP>(p.fwd + o, K, vec, r);
                    }
                    if (p.bwd) {
#pragma unroll
                        for (int k = 0; k < KP; ++k) r[k] = w[k] * eb;
                        store_row<KP>(p.bwd + o, K, vec, r);
                    }
                    if (p.log_alpha) {
#pragma unroll
                        for (int k = 0; k < KP; ++k) r[k] = logf(ca[g][k]) + lat;
                        store_row<KP>(p.log_alpha + o, K, vec, r);
                    }
                    if (p.log_beta) {
#pragma unroll
                        for (int k = 0; k < KP; ++k) r[k] = logf(w[k]) + (float)lbt;
                        store_row<KP>(p.log_beta + o, K, vec, r);
                    }
                    if (t > t_stop) {                           // beta_{t-1} = P (b~_t .* beta_t)
                        float u[KP];
#pragma unroll
                        for (int k = 0; k < KP; ++k) u[k] = cb[g][k] * w[k];
                        mat_vec<KP>(u, P_s, K, w);
                        kb += renorm<KP>(w);
                        mb += (double)cm[g];
                    }
                }
            }
#pragma unroll
            for (int g = 0; g < GS; ++g) {
                cl[g] = nl[g]; cm[g] = nm[g];
#pragma unroll
                for (int k = 0; k < KP; ++k) { ca[g][k] = na[g][k]; cb[g][k] = nb[g][k]; }
            }
        }
    }
}

static size_t sc_al(size_t x) { return (x + 255) & ~(size_t)255; }

static int scan_segment_len(int T) {
    // serial chain lengths: phase 1 L steps, phase 2 S = T/L steps (both directions side by side), phase 3 2L steps.
    // With comparable per-step cost this is minimised near L = sqrt(T/3); a multiple of 16 in [32, 1024].
    int L = 32;
    while (L < 1024 && (long long)L * L * 3 < T) L += 16;
    return L;
}

struct ScanLayout {
    size_t bt, mrow, prod, pexp, msum, bndA, bndLA, bndB, bndLB, ws_a, ws_la, total;
    int L, S;
};

static ScanLayout scan_layout(int B, int T, int K) {
    ScanLayout l;
    l.L = scan_segment_len(T);
    l.S = (T <= 1) ? 0 : (T - 1 + l.L - 1) / l.L;
    const size_t n = (size_t)B * T, ns = (size_t)B * (size_t)max(l.S, 1);
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += sc_al(bytes); return o; };
    l.bt = take(n * K * 4); l.mrow = take(n * 4); l.prod = take(ns * K * K * 4); l.pexp = take(ns * K * 4);
    l.msum = take(ns * 8); l.bndA = take((ns + B) * K * 8); l.bndLA = take((ns + B) * 8); l.bndB = take(ns * K * 8);
    l.bndLB = take(ns * 8); l.ws_a = take(n * K * 4); l.ws_la = take(n * 4);
    l.total = off;
    return l;
}

template <int G, int KP>
static int launch_scan(ScanParams p, cudaStream_t s) {
    (void)G;
    const int64_t n1 = (int64_t)p.B * p.S * p.K;
    if (n1 > 0) {
        scan_products_kernel<KP><<<(unsigned)((n1 + 127) / 128), 128, 0, s>>>(p);
        if (int rc = check_launch("scan_products_kernel")) return rc;
    }
    {
        // two warps x two buffers x nbat segments of (K*K floats + K ints + 1 double) within the default 48 KB
        const size_t per_seg = (size_t)p.K * p.K * 4 + (size_t)p.K * 4 + 8;
        int nbat = (int)(40 * 1024 / (4 * per_seg));
        nbat = nbat < 1 ? 1 : (nbat > 32 ? 32 : nbat);
        scan_boundaries_kernel<KP><<<p.B, 64, 4 * (size_t)nbat * per_seg, s>>>(p, nbat);
        if (int rc = check_launch("scan_boundaries_kernel")) return rc;
    }
    if (p.gamma || p.fwd || p.bwd || p.log_alpha || p.log_beta) {
        const int64_t n3 = (int64_t)p.B * max(p.S, 1);
        scan_fill_kernel<KP><<<(unsigned)((n3 + 63) / 64), 64, 0, s>>>(p);
        if (int rc = check_launch("scan_fill_kernel")) return rc;
    }
    return HMMB200_OK;
}

static int dispatch_scan(const ScanParams &p, cudaStream_t s) {
    const int kp = pad4(p.K);
    if (kp <= 4) return launch_scan<4, 4>(p, s);
    if (kp <= 8) return launch_scan<8, 8>(p, s);
    if (kp <= 12) return launch_scan<16, 12>(p, s);
    if (kp <= 16) return launch_scan<16, 16>(p, s);
    if (kp <= 20) return launch_scan<32, 20>(p, s);
    if (kp <= 24) return launch_scan<32, 24>(p, s);
    if (kp <= 28) return launch_scan<32, 28>(p, s);
    return launch_scan<32, 32>(p, s);
}

}  // namespace hmmb200

using namespace hmmb200;

HMMB200_EXPORT size_t hmmb200_fb_scan_workspace_bytes(int B, int T, int K) {
    if (B <= 0 || T <= 0 || K <= 0 || K > 32) return 0;
    return scan_layout(B, T, K).total;
}

HMMB200_EXPORT int hmmb200_forward_backward_scan_f32(const float *emis, int emis_mode, float floor_eps, int add_rowmax,
                                                     const float *trans_prob, const float *init_prob, int B, int T, int K,
                                                     float *gamma, float *fwd_prob, float *bwd_prob,
                                                     float *log_alpha, float *log_beta, float *loglik,
                                                     void *workspace, size_t workspace_bytes, void *stream) {
    if (B < 0 || T < 0 || K <= 0) return set_error(HMMB200_EINVAL, "forward_backward_scan: bad shape B=%d T=%d K=%d", B, T, K);
    if (B == 0 || T == 0) return HMMB200_OK;
    if (K > 32) return set_error(HMMB200_EUNSUPPORTED, "forward_backward_scan: K <= 32 (got %d)", K);
    if (!emis || !trans_prob || !init_prob) return set_error(HMMB200_EINVAL, "forward_backward_scan: null input");
    if (emis_mode < 0 || emis_mode > 3) return set_error(HMMB200_EINVAL, "forward_backward_scan: bad emis_mode %d", emis_mode);
    const ScanLayout l = scan_layout(B, T, K);
    if (!workspace || workspace_bytes < l.total)
        return set_error(HMMB200_EWORKSPACE, "forward_backward_scan: workspace %zu < %zu bytes", workspace_bytes, l.total);
    if (int rc = require_sm100()) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    uint8_t *w = (uint8_t *)workspace;
    ScanParams p;
    p.emis = emis; p.mode = emis_mode; p.eps = floor_eps;
    p.add_m = ((emis_mode == HMMB200_EMIS_LOG) || (emis_mode == HMMB200_EMIS_LOG_NORM_FLOOR && add_rowmax)) ? 1 : 0;
    p.trans = trans_prob; p.init = init_prob; p.B = B; p.T = T; p.K = K; p.L = l.L; p.S = l.S;
    p.bt = (float *)(w + l.bt); p.mrow = (float *)(w + l.mrow); p.prod = (float *)(w + l.prod); p.pexp = (int *)(w + l.pexp);
    p.msum = (double *)(w + l.msum); p.bndA = (double *)(w + l.bndA); p.bndLA = (double *)(w + l.bndLA);
    p.bndB = (double *)(w + l.bndB); p.bndLB = (double *)(w + l.bndLB); p.ws_a = (float *)(w + l.ws_a); p.ws_la = (float *)(w + l.ws_la);
    p.gamma = gamma; p.fwd = fwd_prob; p.bwd = bwd_prob; p.log_alpha = log_alpha; p.log_beta = log_beta; p.loglik = loglik;
    {
        auto al16 = [](const void *q) { return q == nullptr || ((uintptr_t)q & 15) == 0; };
        p.vec_out = (al16(gamma) && al16(fwd_prob) && al16(bwd_prob) && al16(log_alpha) && al16(log_beta)) ? 1 : 0;
    }
    const int64_t n = (int64_t)B * T;
    scan_emissions_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(p);
    if (int rc = check_launch("scan_emissions_kernel")) return rc;
    return dispatch_scan(p, s);
}
