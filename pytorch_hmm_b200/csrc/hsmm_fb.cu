// hsmm_fb.cu -- explicit-duration forward-backward, second version (BASELINE config 4): fp32 scaled probabilities, the forward and the
// backward sweep of a sequence as two concurrent warps of one CTA, and only ONE multiply-add of the duration sum on the time chain.
//
// (new functionality: the reference has no HSMM backward pass; the recursion is hsmm.cu's, restated)
//     Bg(t,s)  = pi(s) (t = 0) | sum_{s' != s} E(t-1,s') A(s',s)                      a segment of s begins at t
//     E(t,s)   = sum_d Bg(t-d+1,s) dur(s,d) c(s) prod_{tau=t-d+1..t} b~_tau(s)          a segment of s ends at t
//     bend(T-1,s) = 1;  bend(t,s) = sum_{s' != s} A(s,s') bbeg(t+1,s')
//     bbeg(t,s)   = sum_d dur(s,d) c(s) prod_{tau=t..t+d-1} b~_tau(s) bend(t+d-1,s)
// Of the Dmax terms of E(t,s) only d = 1 contains Bg(t,s), i.e. depends on the previous step's result:
//     E(t,s) = Bg(t,s) b~_t(s) durc(1,s)  +  sum_{d >= 2} R_{t-1}[start t-d+1][s] b~_t(s) durc(d,s)
// where the ring R holds the RUNNING products Bg(st,s) prod b~ of the open segments.  The sum over d >= 2 is formed while the
// previous step's vector is still in flight, so a step costs what a step of the plain HMM recursion costs (one K x K matrix-vector
// product between two shared-memory exchanges) plus one FMA -- not a Dmax-long dependent chain in double precision (hsmm.cu: ~2100
// cycles per step, 4.3 ms for B = 128, T = 2000).
// Lanes: LPS lanes per state (3 at K <= 10); lane (s, r) owns the ring slots i = r, r + LPS, ... of state s (its own shared-memory
// words: no synchronisation needed for the ring), the LPS partial sums meet in shuffles.  Scaling: every step multiplies the state by
// rho_t = 2^-e, e = exponent of the previous step's largest entry (exact; the integer exponents are summed), emissions enter as
// b~ = exp(f - max_s f) (hsmm_prep_kernel).  Posteriors (third kernel) from the two scaled sweeps:
//     P(begin at t) = Bg bbeg / p(o),  P(end at t) = E bend / p(o),  gamma_t(s) = sum_{tau<=t} P(begin) - sum_{tau<t} P(end)   (double).
#include "common.cuh"

#include <type_traits>

namespace hmmb200 {

struct HsmmFb2Params {
    const float *f, *segc, *logdur, *logA, *logpi;
    int B, T, K, Dm;
    float *gamma, *total;                     // [B,T,K], [B]
    float *bbegin_out, *bend_out;             // [B,T,K] log values or null
    float *bq;                                // [B,T,K] exp(f - max)
    float *ws_E, *ws_Bg, *ws_be, *ws_bb;      // [B,T,K] scaled forward / backward values
    int *ws_kf, *ws_kb;                       // [B,T] cumulative exponents at frame t
    float *ws_m;                              // [B,T] per-frame emission maxima m_t
    float *ws_S;                              // [B] sum_s E(T-1,s) (scaled)
};

// b~_t(s) = exp(f_t(s) - m_t), m_t = max_s f_t(s): one thread per frame
__global__ void __launch_bounds__(256) hsmm_prep_kernel(HsmmFb2Params p) {
    const int64_t n = (int64_t)p.B * p.T;
    const int K = p.K;
    for (int64_t fr = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; fr < n; fr += (int64_t)gridDim.x * blockDim.x) {
        const float *row = p.f + fr * K;
        float mx = -INFINITY;
        for (int s = 0; s < K; ++s) mx = fmaxf(mx, row[s]);
        if (!(mx > -INFINITY)) mx = 0.f;
        float *o = p.bq + fr * K;
        for (int s = 0; s < K; ++s) o[s] = __expf(row[s] - mx);
        p.ws_m[fr] = mx;
    }
}

// One warp = one sweep (DIR 0 forward, 1 backward) of one sequence.  KP = K padded to a multiple of 4, LPS lanes per state,
// NSLOT = ceil(Dmax / LPS) ring slots per lane; KT / DT = compile-time K / Dmax (0: run-time) so that the per-step addresses are
// immediates.  The step is ~100 instructions of a lone warp, so it is written for instruction count: lanes that own no state run the
// same code on a scratch column (pitch K + 1) instead of being predicated, the power-of-two rescale is derived every fourth step only
// (exact; between two rescales the state drifts by at most a few decades, far from the fp32 range), pointers advance by constants.
constexpr int HS_RESCALE = 4;
template <int KP, int LPS, int NSLOT, int KT, int DT, int DIR>
__device__ __forceinline__ void hsmm_sweep(const HsmmFb2Params &p, float *sm, int b) {
    const int K = KT > 0 ? KT : p.K, Dm = DT > 0 ? DT : p.Dm, T = p.T;
    const int KS = K + 1;                                       // row pitch of ring / durc2: column K is the scratch column
    const int lane = threadIdx.x & 31;
    const int s = lane / LPS, r = lane % LPS;
    const bool ok = s < K;
    const int sc = ok ? s : K;
    // shared memory of this sweep: vec [2][KP] (ping-pong), ring [NSLOT*LPS][KS], durc2 [2 Dm][KS], emission chunks [2][32 K]
    float *vec = sm;                                            // (first: 16-byte aligned for the float4 reads)
    float *ring = vec + 2 * KP;
    float *durc2 = ring + NSLOT * LPS * KS;
    float *ech = durc2 + 2 * Dm * KS;
    for (int i = lane; i < NSLOT * LPS * KS; i += 32) ring[i] = 0.f;
    for (int i = lane; i < 2 * Dm * KS; i += 32) {
        const int d = (i / KS) % Dm, st = i % KS;
        durc2[i] = (st < K) ? expf(p.logdur[st * Dm + d] + (p.segc ? p.segc[st] : 0.f)) : 0.f;
    }
    for (int i = lane; i < 2 * KP; i += 32) vec[i] = 0.f;
    // the lane's matrix slice: forward  A(s', s) over s' (column s), backward A(s, s') over s' (row s); the diagonal is excluded
    float Am[KP];
#pragma unroll
    for (int j = 0; j < KP; ++j) {
        float v = 0.f;
        if (ok && j < K && j != s) v = expf((DIR == 0) ? p.logA[j * K + s] : p.logA[s * K + j]);
        Am[j] = v;
    }
    __syncwarp();
    const float durc1 = durc2[sc];                              // durc(1, s)
    const float pi = (DIR == 0) ? ((ok && p.logpi) ? expf(p.logpi[s]) : (ok ? 1.f : 0.f)) : (ok ? 1.f : 0.f);
    const size_t base = (size_t)b * T;
    const bool writer = ok && r == 0;
    float *ws_v = ((DIR == 0) ? p.ws_E : p.ws_bb) + (base + ((DIR == 0) ? 0 : T - 1)) * K + sc;    // E (forward) / bbeg (backward)
    float *ws_g = ((DIR == 0) ? p.ws_Bg : p.ws_be) + (base + ((DIR == 0) ? 0 : T - 1)) * K + sc;   // Bg / bend
    int *ws_k = ((DIR == 0) ? p.ws_kf : p.ws_kb) + base + ((DIR == 0) ? 0 : T - 1);
    const int wstep = (DIR == 0) ? K : -K, kstep = (DIR == 0) ? 1 : -1;
    // emission chunks of 32 frames (in sweep order), double-buffered in shared memory; the next chunk travels in registers
    const int nch = (T + 31) / 32;
    float nxt[KP];
    auto fetch = [&](int c) {                                   // chunk c covers sweep positions 32 c .. 32 c + 31
        const int n0 = c * 32, nf = min(32, T - n0);
        const int f_lo = (DIR == 0) ? n0 : T - n0 - nf;         // lowest frame of the chunk
        const float *src = p.bq + (base + f_lo) * K;
#pragma unroll
        for (int q = 0; q < KP; ++q) { const int e = lane + 32 * q; nxt[q] = (q < K && e < nf * K) ? src[e] : 0.f; }
    };
    auto stash = [&](int c) {
        float *dst = ech + (c & 1) * 32 * K;
#pragma unroll
        for (int q = 0; q < KP; ++q) if (q < K) dst[lane + 32 * q] = nxt[q];
    };
    fetch(0);
    stash(0);
    __syncwarp();
    // the lane's slots: ring_l[j * LPS * KS]; their duration weights: dp[-j * LPS * KS] with dp = durc2 + (a0 + Dm) * KS + sc,
    // a0 = (position - r) mod Dm.  Slot j is a real slot when r + LPS j < Dm (always, except possibly for the last j).
    // The lane's slots are its own (no other lane reads them) and every one of them is rewritten every step, so they live in REGISTERS:
    // as a shared-memory ring they cost seven loads and seven stores per step whose latencies a single in-order warp cannot hide,
    // and the stores kept the compiler from hoisting the step's other loads.
    float rg[NSLOT];
#pragma unroll
    for (int j = 0; j < NSLOT; ++j) rg[j] = 0.f;
    int a0 = (Dm - r % Dm) % Dm;
    const float *dp = durc2 + (a0 + Dm) * KS + sc;
    // slot j of this lane exists when r + LPS j < Dm; with a compile-time Dmax that is a run-time question for the last j only
    auto slot_ok = [&](int j) -> bool {
        if (DT > 0) return (LPS * j + LPS - 1 < DT) || (r + LPS * j < DT);
        return r + LPS * j < Dm;
    };
    const int src0 = ok ? s * LPS : lane;
    int ksum = 0;
    float last_e = 0.f;
    auto step = [&](auto rescale_tag, int n, float bq) {
        constexpr bool RESCALE = decltype(rescale_tag)::value;
        // ---- off the chain: the open segments (d >= 2) from the ring as the previous step left it ----
        float x[NSLOT], pp = 0.f;
#pragma unroll
        for (int j = 0; j < NSLOT; ++j) {
            const bool vj = slot_ok(j);
            const float old = rg[j];                            // (slots past Dm exist and stay zero)
            // a slot whose duration index is 0 opens NOW: its old content (duration Dm + 1) expires
            x[j] = (a0 == LPS * j || !vj) ? 0.f : old * bq;
            pp = fmaf(x[j], vj ? dp[-j * LPS * KS] : 0.f, pp);
        }
        {   // sum over the LPS lanes of the state, in the same order in every lane (all get the identical value)
            float tot = __shfl_sync(FULL_MASK, pp, src0);
#pragma unroll
            for (int o = 1; o < LPS; ++o) tot += __shfl_sync(FULL_MASK, pp, ok ? src0 + o : lane);
            pp = tot;
        }
        // ---- the chain: previous vector -> matrix-vector product -> this step's vector ----
        const float4 *pv = reinterpret_cast<const float4 *>(vec + ((n + 1) & 1) * KP);
        float v[KP];
#pragma unroll
        for (int i4 = 0; i4 < KP / 4; ++i4) { const float4 q = pv[i4]; v[4 * i4] = q.x; v[4 * i4 + 1] = q.y; v[4 * i4 + 2] = q.z; v[4 * i4 + 3] = q.w; }
        float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
        for (int j = 0; j < KP; j += 2) { acc0 = fmaf(v[j], Am[j], acc0); acc1 = fmaf(v[j + 1], Am[j + 1], acc1); }
        float acc = acc0 + acc1;
        if (n == 0) acc = pi;                                   // Bg(0,s) = pi(s)  /  bend(T-1,s) = 1
        float g, e, fresh_v;
        if (RESCALE) {
            float mx = 0.f;
#pragma unroll
            for (int j = 0; j < KP; ++j) mx = fmaxf(mx, v[j]);
            if (n == 0) mx = 1.f;
            const unsigned ebits = __float_as_uint(mx) >> 23;   // (mx >= 0: no sign bit)
            const bool scale_ok = ebits != 0u && ebits != 255u;
            const float rho = scale_ok ? __uint_as_float((254u - ebits) << 23) : 1.f;
            if (scale_ok) ksum += (int)ebits - 127;
            g = acc * rho;                                      // Bg(t,s) / bend(t,s), scaled
            e = fmaf(g, bq * durc1, pp * rho);                  // E(t,s) / bbeg(t,s), scaled
            fresh_v = g * bq;
#pragma unroll
            for (int j = 0; j < NSLOT; ++j)
                if (slot_ok(j)) rg[j] = (a0 == LPS * j) ? fresh_v : x[j] * rho;
        } else {
            g = acc;
            e = fmaf(g, bq * durc1, pp);
            fresh_v = g * bq;
#pragma unroll
            for (int j = 0; j < NSLOT; ++j)
                if (slot_ok(j)) rg[j] = (a0 == LPS * j) ? fresh_v : x[j];
        }
        if (r == 0) vec[(n & 1) * KP + s] = e;                  // (lanes without a state write the padding of the vector: Am = 0 there)
        if (writer) { *ws_v = e; *ws_g = g; }
        if (lane == 0) *ws_k = ksum;
        ws_v += wstep; ws_g += wstep; ws_k += kstep;
        last_e = e;
        ++a0; dp += KS;
        if (a0 == Dm) { a0 = 0; dp -= Dm * KS; }
        __syncwarp();
    };
    using Yes = std::integral_constant<bool, true>;
    using No = std::integral_constant<bool, false>;
    for (int c = 0; c < nch; ++c) {
        if (c + 1 < nch) fetch(c + 1);
        const int n0 = c * 32, nf = min(32, T - n0);
        const float *eb = ech + (c & 1) * 32 * K + (ok ? s : 0);
        int u = 0;
        for (; u + HS_RESCALE <= nf; u += HS_RESCALE) {         // (n0 is a multiple of 32: the rescale steps are n % 4 == 0)
            float bqv[HS_RESCALE];
#pragma unroll
            for (int i = 0; i < HS_RESCALE; ++i) bqv[i] = eb[((DIR == 0) ? u + i : nf - 1 - u - i) * K];
            step(Yes{}, n0 + u, bqv[0]);
#pragma unroll
            for (int i = 1; i < HS_RESCALE; ++i) step(No{}, n0 + u + i, bqv[i]);
        }
        for (; u < nf; ++u) step(Yes{}, n0 + u, eb[((DIR == 0) ? u : nf - 1 - u) * K]);
        if (c + 1 < nch) stash(c + 1);
        __syncwarp();
    }
    if (DIR == 0) {
        float tot = writer ? last_e : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(FULL_MASK, tot, o);
        if (lane == 0) p.ws_S[b] = tot;
    }
}

__host__ __device__ inline size_t hsmm_fb2_sweep_floats(int KP, int LPS, int NSLOT, int K, int Dm) {
    return ((size_t)2 * KP + (size_t)NSLOT * LPS * (K + 1) + (size_t)2 * Dm * (K + 1) + (size_t)2 * 32 * K + 3) & ~(size_t)3;
}

template <int KP, int LPS, int NSLOT, int KT, int DT>
__global__ void __launch_bounds__(64) hsmm_fb2_kernel(HsmmFb2Params p) {
    extern __shared__ __align__(16) float sm_h2[];
    const size_t per = hsmm_fb2_sweep_floats(KP, LPS, NSLOT, p.K, p.Dm);
    if ((threadIdx.x >> 5) == 0) hsmm_sweep<KP, LPS, NSLOT, KT, DT, 0>(p, sm_h2, blockIdx.x);
    else hsmm_sweep<KP, LPS, NSLOT, KT, DT, 1>(p, sm_h2 + per, blockIdx.x);
}

// posteriors: one CTA per sequence, thread = (part, state): 32 parts split the frames into contiguous runs (the K threads of a part
// read K contiguous floats per frame).  Pass 1 forms P(begin) / P(end) per frame (left in gamma / over the bend scratch) and the
// run totals; the runs meet in a scan over the parts through shared memory (double); pass 2 turns them into gamma.
__global__ void __launch_bounds__(1024) hsmm_post_kernel(HsmmFb2Params p) {
    __shared__ double part_sum[32 * 32], part_m[32];
    const int K = p.K, T = p.T, b = blockIdx.x;
    const int part = threadIdx.x / K, s = threadIdx.x % K;
    const size_t base = (size_t)b * T;
    const int kT = p.ws_kf[base + T - 1];
    const float inv = 1.f / p.ws_S[b];
    const int per = (T + 31) / 32, t_lo = min(T, part * per), t_hi = min(T, t_lo + per);
    // emission maxima: run totals (every thread of a part adds the same numbers; the s == 0 thread publishes)
    double mrun = 0.0;
    for (int t = t_lo; t < t_hi; ++t) mrun += (double)p.ws_m[base + t];
    if (s == 0) part_m[part] = mrun;
    __syncthreads();
    double Mtot = 0.0, Mbefore = 0.0;                           // sum of all m_t / of those before this thread's run
    for (int q = 0; q < 32; ++q) { if (q < part) Mbefore += part_m[q]; Mtot += part_m[q]; }
    if (threadIdx.x == 0) p.total[b] = (float)(log((double)p.ws_S[b]) + 0.69314718055994530942 * (double)kT + Mtot);
    const bool want_log = p.bend_out != nullptr || p.bbegin_out != nullptr;
    double run = 0.0, Mc = Mbefore;
#pragma unroll 4
    for (int t = t_lo; t < t_hi; ++t) {
        const size_t o = (base + t) * K + s;
        const int kb = p.ws_kb[base + t];
        int ke = p.ws_kf[base + t] + kb - kT;
        ke = max(-126, min(126, ke));
        const float sc = __int_as_float((ke + 127) << 23) * inv;
        const float be = p.ws_be[o], bb = p.ws_bb[o];
        const float pb = p.ws_Bg[o] * bb * sc;                  // P(a segment of s begins at t)
        const float pe = p.ws_E[o] * be * sc;                   // P(a segment of s ends at t)
        if (want_log) {
            const double mt = (double)p.ws_m[base + t];
            if (p.bbegin_out) p.bbegin_out[o] = (float)(log((double)bb) + 0.69314718055994530942 * (double)kb + (Mtot - Mc));
            Mc += mt;
            if (p.bend_out) p.bend_out[o] = (float)(log((double)be) + 0.69314718055994530942 * (double)kb + (Mtot - Mc));
        }
        p.gamma[o] = pb;
        p.ws_be[o] = pe;
        run += (double)pb - (double)pe;
    }
    part_sum[part * 32 + s] = run;
    __syncthreads();
    double cum = 0.0;                                           // sum over the frames before this thread's run of (begin - end)
    for (int q = 0; q < part; ++q) cum += part_sum[q * 32 + s];
#pragma unroll 4
    for (int t = t_lo; t < t_hi; ++t) {
        const size_t o = (base + t) * K + s;
        cum += (double)p.gamma[o];
        const float g = (float)cum;
        cum -= (double)p.ws_be[o];
        p.gamma[o] = fminf(fmaxf(g, 0.f), 1.f);
    }
}

size_t hsmm_fb2_workspace_bytes(int B, int T, int K) {
    const size_t n = (size_t)B * T;
    return 5 * n * K * sizeof(float) + 3 * n * sizeof(float) + (size_t)B * sizeof(float) + 256;
}

// 0 launched, 1 shape not covered by this version (caller runs hsmm.cu's kernel)
int launch_hsmm_fb2(const float *f, const float *segc, const float *logdur, const float *logA, const float *logpi, int B, int T, int K, int Dm,
                    float *gamma, float *total, float *bbegin, float *bend, void *workspace, cudaStream_t s) {
    int KP, LPS, NSLOT;
    if (K <= 10 && Dm <= 21) { KP = 12; LPS = 3; NSLOT = 7; }
    else if (K <= 10 && Dm <= 42) { KP = 12; LPS = 3; NSLOT = 14; }
    else if (K <= 16 && Dm <= 32) { KP = 16; LPS = 2; NSLOT = 16; }
    else if (K <= 32 && Dm <= 32) { KP = 32; LPS = 1; NSLOT = 32; }
    else return 1;
    const size_t smem = 2 * hsmm_fb2_sweep_floats(KP, LPS, NSLOT, K, Dm) * sizeof(float);
    if (smem > 48 * 1024) return 1;
    const size_t n = (size_t)B * T;
    uint8_t *w = (uint8_t *)workspace;
    HsmmFb2Params p;
    p.f = f; p.segc = segc; p.logdur = logdur; p.logA = logA; p.logpi = logpi; p.B = B; p.T = T; p.K = K; p.Dm = Dm;
    p.gamma = gamma; p.total = total; p.bbegin_out = bbegin; p.bend_out = bend;
    p.ws_m = (float *)w;  w += n * sizeof(float);
    p.bq = (float *)w;    w += n * K * sizeof(float);
    p.ws_E = (float *)w;  w += n * K * sizeof(float);
    p.ws_Bg = (float *)w; w += n * K * sizeof(float);
    p.ws_be = (float *)w; w += n * K * sizeof(float);
    p.ws_bb = (float *)w; w += n * K * sizeof(float);
    p.ws_kf = (int *)w;   w += n * sizeof(int);
    p.ws_kb = (int *)w;   w += n * sizeof(int);
    p.ws_S = (float *)w;
    hsmm_prep_kernel<<<(unsigned)min((n + 255) / 256, (size_t)148 * 8), 256, 0, s>>>(p);
    if (int rc = check_launch("hsmm_prep_kernel")) return rc;
    if (K == 10 && Dm == 20) hsmm_fb2_kernel<12, 3, 7, 10, 20><<<B, 64, smem, s>>>(p);       // BASELINE config 4
    else if (KP == 12 && NSLOT == 7) hsmm_fb2_kernel<12, 3, 7, 0, 0><<<B, 64, smem, s>>>(p);
    else if (KP == 12) hsmm_fb2_kernel<12, 3, 14, 0, 0><<<B, 64, smem, s>>>(p);
    else if (KP == 16) hsmm_fb2_kernel<16, 2, 16, 0, 0><<<B, 64, smem, s>>>(p);
    else hsmm_fb2_kernel<32, 1, 32, 0, 0><<<B, 64, smem, s>>>(p);
    if (int rc = check_launch("hsmm_fb2_kernel")) return rc;
    hsmm_post_kernel<<<B, 32 * K, 0, s>>>(p);
    return check_launch("hsmm_post_kernel");
}

}  // namespace hmmb200
