// recursion_xlk.cu -- HMM recursions for 512 < K <= 2048 states: one launch per time step (sm_100a).
//
// Above 512 states the transition matrix no longer fits the registers of one 8-CTA cluster (recursion_largek.cu), so the step
//     out(a) = sum_b v(b) M(b,a)   (forward: M = P, backward: M = P^T)      /      out(j) = max_i (v(i) + logP(i,j))   (Viterbi)
// is a plain batched matrix product over the whole GPU: grid = (K/64 column blocks) x (B/8 sequence blocks), M streamed from L2
// (4 MB at K = 1024) once per step and CTA, the K-reduction split over four thread groups.  The T steps are T launches on the
// stream; the state lives in HBM between them.  Scaling, emission handling, outputs and the traceback are those of the cluster
// kernels (same workspace layout, same lk_* post-processing kernels), so results agree with them to rounding (forward-backward) or
// bit for bit (Viterbi: exact maxima, one fp32 add of log b).  This is the capacity path, not a tuned one: ~1000x the reference's
// CPU recursion, a small multiple of the cluster kernels' time per state-pair.
#include "common.cuh"

namespace hmmb200 {

constexpr int XL_NC = 64;       // output states per CTA
constexpr int XL_NS = 8;        // sequences per CTA
constexpr int XL_KQ = 4;        // thread groups splitting the reduction
constexpr int XL_THREADS = XL_NC * XL_KQ;

struct XlParams {
    const float *emis;      // [B,T,K]
    int mode;
    float eps;
    const float *M;         // [K,K]: forward P, backward P^T, Viterbi log P  (row = source state)
    const float *init;      // [K] probabilities (forward) / log (Viterbi); unused backward
    const float *rowmax;    // [B,T] or null
    int B, T, K;
    int dir;                // 0 forward, 1 backward
    int t;                  // sweep position of this launch; frame = t (forward / Viterbi) or T-1-t (backward)
    const float *v_prev;    // [B,K] previous vector (fb) -- Viterbi reads delta[t-1]
    float *v_cur;           // [B,K]
    float *ws_out;          // fb: ws_a (forward) / ws_b (backward) [B,T,K];  Viterbi: delta [B,T,K]
    float *ws_l;            // [B,T] integer exponents as float bits (fb)
    unsigned *mx;           // [3][B] running maxima of the vectors (float bits, values >= 0), slot = sweep position % 3
    int *ksum;              // [B] running exponent
};

template <bool VIT>
__global__ void __launch_bounds__(XL_THREADS) xl_step_kernel(XlParams p) {
    __shared__ __align__(16) float v_s[64][XL_NS];          // a 64-state slice of the previous vectors, [state][sequence]
    __shared__ float part[XL_KQ][XL_NS][XL_NC];
    const int K = p.K, T = p.T, B = p.B;
    const int jl = threadIdx.x % XL_NC, kq = threadIdx.x / XL_NC;
    const int j = blockIdx.x * XL_NC + jl;
    const int b0 = blockIdx.y * XL_NS;
    const int t = p.t, f = p.dir ? T - 1 - t : t;
    const float NEUTRAL = VIT ? -INFINITY : 0.f;
    float acc[XL_NS];
#pragma unroll
    for (int s = 0; s < XL_NS; ++s) acc[s] = NEUTRAL;
    if (t > 0) {
        const float *vp = VIT ? p.ws_out : p.v_prev;        // Viterbi: the previous delta row
        for (int i0 = 0; i0 < K; i0 += 64) {
            __syncthreads();
            for (int e = threadIdx.x; e < 64 * XL_NS; e += XL_THREADS) {
                const int s = e / 64, ii = e % 64, i = i0 + ii, b = b0 + s;
                float x = NEUTRAL;
                if (i < K && b < B) x = VIT ? vp[((size_t)b * T + (t - 1)) * K + i] : vp[(size_t)b * K + i];
                v_s[ii][s] = x;
            }
            __syncthreads();
            // this thread's quarter of the 64 source states
            float mv[64 / XL_KQ];                            // all of the thread's matrix elements of this slice in flight at once
#pragma unroll
            for (int q = 0; q < 64 / XL_KQ; ++q) {
                const int i = i0 + kq + XL_KQ * q;
                mv[q] = (i < K && j < K) ? __ldg(p.M + (size_t)i * K + j) : NEUTRAL;
            }
#pragma unroll
            for (int q = 0; q < 64 / XL_KQ; ++q) {
                const int ii = kq + XL_KQ * q;
                const float m = mv[q];
                const float4 va = *reinterpret_cast<const float4 *>(&v_s[ii][0]), vb = *reinterpret_cast<const float4 *>(&v_s[ii][4]);
                const float vv[XL_NS] = {va.x, va.y, va.z, va.w, vb.x, vb.y, vb.z, vb.w};
#pragma unroll
                for (int s = 0; s < XL_NS; ++s) {
                    if (VIT) acc[s] = fmaxf(acc[s], __fadd_rn(vv[s], m));
                    else acc[s] = fmaf(vv[s], m, acc[s]);
                }
            }
        }
    }
#pragma unroll
    for (int s = 0; s < XL_NS; ++s) part[kq][s][jl] = acc[s];
    __syncthreads();
    // finals: thread (jl, kq) finishes sequences kq and kq + 4
    for (int s = kq; s < XL_NS; s += XL_KQ) {
        const int b = b0 + s;
        const bool ok = b < B && j < K;
        float a;
        if (VIT) a = fmaxf(fmaxf(part[0][s][jl], part[1][s][jl]), fmaxf(part[2][s][jl], part[3][s][jl]));
        else a = (part[0][s][jl] + part[1][s][jl]) + (part[2][s][jl] + part[3][s][jl]);      // fixed order: deterministic
        const float raw = ok ? p.emis[((size_t)b * T + f) * K + j] : 0.f;
        const float mf = (p.rowmax != nullptr && b < B) ? p.rowmax[(size_t)b * T + f] : 0.f;
        if (VIT) {
            float lb;
            if (p.mode == HMMB200_EMIS_LOG) lb = raw;
            else if (p.mode == HMMB200_EMIS_PROB_FLOOR) lb = logf(raw + p.eps);
            else if (p.mode == HMMB200_EMIS_LOG_EXP_FLOOR) lb = logf(expf(raw) + p.eps);
            else lb = logf(expf(raw - mf) + p.eps);
            if (t == 0) a = ok ? p.init[j] : -INFINITY;
            if (ok) p.ws_out[((size_t)b * T + t) * K + j] = __fadd_rn(a, lb);       // delta_t = max_i(..) + log b_t  (hmm.py:168)
        } else {
            float bq;
            if (p.mode == HMMB200_EMIS_PROB_FLOOR) bq = raw + p.eps;
            else if (p.mode == HMMB200_EMIS_LOG_EXP_FLOOR) bq = expf(raw) + p.eps;
            else bq = expf(raw - mf) + ((p.mode == HMMB200_EMIS_LOG_NORM_FLOOR) ? p.eps : 0.f);
            float r = 1.f;
            int de = 0;
            if (t == 0) {
                a = (p.dir == 0) ? (ok ? p.init[j] : 0.f) : (ok ? 1.f : 0.f);
            } else if (b < B) {                              // power-of-two normaliser from the previous vector's largest entry
                const unsigned eb = p.mx[((t - 1) % 3) * B + b] >> 23;
                if (eb != 0u && eb != 255u) { de = (int)eb - 127; r = __uint_as_float((254u - eb) << 23); }
            }
            const float pre = a * r, wv = ok ? a * (bq * r) : 0.f;
            if (ok) {
                p.v_cur[(size_t)b * K + j] = wv;
                p.ws_out[((size_t)b * T + f) * K + j] = (p.dir == 0) ? wv : pre;
            }
            // the 32 lanes of a warp finish the same sequence: one atomic per warp (a per-thread atomic put 1024 same-address
            // atomics per sequence and step through the L2)
            const unsigned wmax = __reduce_max_sync(FULL_MASK, __float_as_uint(wv));      // wv >= 0
            if ((threadIdx.x & 31) == 0 && b < B) atomicMax(p.mx + (t % 3) * B + b, wmax);
            if (blockIdx.x == 0 && jl == 0 && b < B) {       // one thread per sequence keeps the running exponent
                const int ks = ((t == 0) ? 0 : p.ksum[b]) + de;
                p.ksum[b] = ks;
                p.ws_l[(size_t)b * T + f] = __int_as_float(ks);
            }
        }
    }
    if (!VIT && blockIdx.x == 0 && blockIdx.y == 0)          // the slot the NEXT step accumulates into
        for (int e = threadIdx.x; e < B; e += XL_THREADS) p.mx[((t + 1) % 3) * B + e] = 0u;
}

// log of the sum of the last forward vector (the log scale of the last frame is added by lk_logscale_kernel)
__global__ void __launch_bounds__(128) xl_loglik_kernel(const float *v, int B, int K, float *loglik) {
    const int b = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (b >= B) return;
    float tot = 0.f;
    for (int k = lane; k < K; k += 32) tot += v[(size_t)b * K + k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(FULL_MASK, tot, o);
    if (lane == 0) loglik[b] = logf(tot);
}

// extra workspace of this path behind the cluster kernels' layout: two [B,K] vectors, P^T, maxima slots, exponents
size_t xlk_extra_bytes(int B, int K) {
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    return 2 * al((size_t)B * K * sizeof(float)) + al((size_t)K * K * sizeof(float)) + al(3 * (size_t)B * sizeof(unsigned)) + al((size_t)B * sizeof(int));
}

int xlk_sweep(int mode3, const float *emis, int emis_mode, float eps, const float *M, const float *init, const float *rowmax,
              int B, int T, int K, float *ws_out, float *ws_l, void *extra, float *loglik, cudaStream_t s) {
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    uint8_t *w = (uint8_t *)extra;
    float *v0 = (float *)w; w += al((size_t)B * K * sizeof(float));
    float *v1 = (float *)w; w += al((size_t)B * K * sizeof(float));
    w += al((size_t)K * K * sizeof(float));                 // (P^T: owned by the caller)
    unsigned *mx = (unsigned *)w; w += al(3 * (size_t)B * sizeof(unsigned));
    int *ksum = (int *)w;
    XlParams p;
    p.emis = emis; p.mode = emis_mode; p.eps = eps; p.M = M; p.init = init; p.rowmax = rowmax; p.B = B; p.T = T; p.K = K;
    p.dir = (mode3 == 1) ? 1 : 0; p.ws_out = ws_out; p.ws_l = ws_l; p.mx = mx; p.ksum = ksum;
    cudaMemsetAsync(mx, 0, 3 * (size_t)B * sizeof(unsigned), s);
    dim3 grid((K + XL_NC - 1) / XL_NC, (B + XL_NS - 1) / XL_NS);
    for (int t = 0; t < T; ++t) {
        p.t = t; p.v_prev = (t & 1) ? v0 : v1; p.v_cur = (t & 1) ? v1 : v0;
        if (mode3 == 2) xl_step_kernel<true><<<grid, XL_THREADS, 0, s>>>(p);
        else xl_step_kernel<false><<<grid, XL_THREADS, 0, s>>>(p);
    }
    if (int rc = check_launch("xl_step_kernel")) return rc;
    if (mode3 == 0 && loglik != nullptr) {
        xl_loglik_kernel<<<(B + 3) / 4, 128, 0, s>>>(((T - 1) & 1) ? v1 : v0, B, K, loglik);
        return check_launch("xl_loglik_kernel");
    }
    return HMMB200_OK;
}

}  // namespace hmmb200
