// recursion_tv.cu -- recursions with TIME-VARYING transitions (K <= 32): the NeuralHMM form of the path, sm_100a.
//
//   replaces  NeuralHMM._forward_algorithm / _backward_algorithm   pytorch_hmm/neural.py:403-461
//             NeuralHMM.viterbi_decode (recursion + backtrack)      pytorch_hmm/neural.py:463-511
// The reference steps  log_forward[t] = logsumexp_i(log_forward[t-1][i] + log_trans[:, t-1][i][j]) + log_obs[t][j]  through T ATen
// launches with a [B,K,K] temporary each; the transition matrix of step t-1 -> t is its own [K,K] slice of a [B,T,K,K] tensor (a
// network output), log-emissions are used as they are (no floor).  Here: one warp per sequence, lane = state.
//   * the K x K slice of the NEXT steps is prefetched into registers two steps ahead (the step's HBM traffic, K*K*4 bytes per frame,
//     is 12x the emission traffic at K = 12: this recursion is bandwidth-heavy as well as latency-bound);
//   * forward / backward run in scaled-probability space like the fixed-transition sweeps (recursion_smallk.cuh): per-frame maxima
//     divided out of the emissions, a power-of-two normaliser taken from the previous step's largest entry (one REDUX, exact,
//     integer exponent bookkeeping), log-scales accumulated in double off the chain; the two sweeps are blockIdx.y = 0 / 1 of one
//     launch and leave scaled vectors + log-scales in the SAME workspace layout, so the posterior kernel is shared;
//   * Viterbi is the reference's two roundings (max of fp32 sums, then + log b) with torch.max's first-index tie rule, uint8
//     backpointers in shared memory (global when T is too long), traceback by the same warp.
#include "recursion_smallk.cuh"

#include <stdlib.h>

namespace hmmb200 {

struct TvParams {
    const float *logb;     // [B,T,K] log-emissions
    const float *trans;    // [B,T,K,K]: probabilities (forward-backward) or log-probabilities (Viterbi); slice t = step t -> t+1
    const float *init;     // [K]: probabilities / log-probabilities
    int B, T, K;
    float *ws_a, *ws_b, *ws_la, *ws_lb, *loglik;          // forward-backward scratch (layout of hmmb200_forward_backward_f32)
    float *delta; uint8_t *psi_out; int64_t *states; float *score; uint8_t *psi_ws; int psi_in_smem;   // Viterbi
    int chf;               // bulk-staged form: frames per staged chunk (0: register prefetch)
};

// Bulk-staged feed (K % 4 == 0, 16-byte aligned tensors): the transition slices and emission rows of the next `chf` steps are two
// contiguous byte ranges, fetched with cp.async.bulk into the warp's other staging buffer while the current chunk is consumed from
// shared memory.  [The register prefetch below reaches two steps ahead -- ~0.1 us against an HBM latency of ~1 us: 600 ns per step at
// K = 12, B = 256, T = 2000 (tools/bench_next_rows.py), ten times the fixed-transition recursion.]
struct TvStage {
    float *buf[2];
    uint64_t *bar;
    int chf, per_tr, per_e;
    __device__ TvStage(float *sm, uint64_t *b, int chf_, int K) : bar(b), chf(chf_), per_tr(K * K), per_e(K) {
        buf[0] = sm; buf[1] = sm + (size_t)chf_ * (K * K + K);
    }
    __device__ void init() const {
        if ((threadIdx.x & 31) == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32_tv(bar)) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32_tv(bar + 1)) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
    }
    static __device__ uint32_t smem_u32_tv(const void *q) { return (uint32_t)__cvta_generic_to_shared(q); }
    // lane 0: chunk c = `cnt` slices from tr_src and `cnt` emission rows from e_src
    __device__ void issue(int c, const float *tr_src, const float *e_src, int cnt) const {
        float *dst = buf[c & 1];
        const uint32_t bt = (uint32_t)cnt * per_tr * 4u, be = (uint32_t)cnt * per_e * 4u;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");          // the buffer's earlier generic reads come first
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32_tv(bar + (c & 1))), "r"(bt + be) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32_tv(dst)), "l"(tr_src), "r"(bt), "r"(smem_u32_tv(bar + (c & 1))) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32_tv(dst + (size_t)chf * per_tr)), "l"(e_src), "r"(be), "r"(smem_u32_tv(bar + (c & 1))) : "memory");
    }
    __device__ void wait(int c) const {
        const uint32_t a = smem_u32_tv(bar + (c & 1)), parity = (uint32_t)((c >> 1) & 1);
        asm volatile("{\n\t.reg .pred p;\n\tTVW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t@p bra TVD_%=;\n\tbra TVW_%=;\n\tTVD_%=:\n\t}"
                     ::"r"(a), "r"(parity), "r"(1000000u) : "memory");
    }
    __device__ const float *tr(int c) const { return buf[c & 1]; }
    __device__ const float *em(int c) const { return buf[c & 1] + (size_t)chf * per_tr; }
};

constexpr int TV_PF = 2;           // transition slices in flight ahead of the step that uses them

// DIR 0: alpha_t(j) = (sum_i alpha_{t-1}(i) P_{t-1}(i,j)) b_t(j)            lane j holds COLUMN j of the slice   (neural.py:424-431)
// DIR 1: beta_t(i)  = sum_j P_t(i,j) b_{t+1}(j) beta_{t+1}(j)               lane i holds ROW i of the slice      (neural.py:448-459)
template <int KP, int DIR, bool BULK>
__device__ __forceinline__ void tv_sweep(const TvParams &p, float *stage_sm, uint64_t *stage_bar) {
    const int b = blockIdx.x, lane = threadIdx.x & 31;
    const int K = p.K, T = p.T;
    const bool ok = lane < K;
    const float *lb = p.logb + (size_t)b * T * K;
    const float *tr = p.trans + (size_t)b * T * K * K;
    float *ws = (DIR == 0 ? p.ws_a : p.ws_b) + (size_t)b * T * K;
    float *wl = (DIR == 0 ? p.ws_la : p.ws_lb) + (size_t)b * T;

    // the slice used by the step that PRODUCES frame f: DIR 0 -> slice f-1 (f >= 1); DIR 1 -> slice f (f <= T-2)
    auto load_slice = [&](int f, float (&c)[KP]) {
        const int sl = (DIR == 0) ? f - 1 : f;
        const bool v = ok && f >= 0 && f < T && sl >= 0 && sl < T - 1;
        const float *s = tr + (size_t)(v ? sl : 0) * K * K;
#pragma unroll
        for (int i = 0; i < KP; ++i) c[i] = (v && i < K) ? __ldg((DIR == 0) ? s + i * K + lane : s + lane * K + i) : 0.f;
    };
    auto frame = [&](int n) { return DIR == 0 ? n : T - 1 - n; };      // n-th frame in sweep order

    float col[TV_PF][KP];
    float eraw[TV_PF];
    TvStage stg(stage_sm, stage_bar, p.chf, K);
    // chunk c of the bulk-staged feed: steps n0 .. n0+cnt-1; its lowest frame and the slices / emission rows it needs are contiguous
    const int n_chunks = BULK ? (T - 1 + p.chf - 1) / p.chf : 0;
    auto chunk_issue = [&](int c) {
        const int n0 = 1 + c * p.chf, cnt = min(p.chf, T - n0);
        const int f_lo = (DIR == 0) ? n0 : T - n0 - cnt;
        stg.issue(c, tr + (size_t)((DIR == 0) ? n0 - 1 : f_lo) * K * K, lb + (size_t)f_lo * K, cnt);
    };
    if (BULK) {
        stg.init();
        if (lane == 0 && n_chunks > 0) chunk_issue(0);
    } else {
#pragma unroll
        for (int q = 0; q < TV_PF; ++q) {
            const int n = 1 + q;
            load_slice(n < T ? frame(n) : -1, col[q]);
            eraw[q] = (ok && n < T) ? __ldg(lb + (size_t)frame(n) * K + lane) : -INFINITY;
        }
    }
    // frame 0 of the sweep
    double msum = 0.0;                 // sum of the per-frame maxima divided out so far (sweep order)
    int ksum = 0;
    float r_cur = 1.f; int k_cur = 0;
    auto set_scale = [&](float w) {
        const unsigned eb = __reduce_max_sync(FULL_MASK, ok ? __float_as_uint(w) : 0u) >> 23;
        k_cur = (int)eb - 127;
        r_cur = __uint_as_float((254u - eb) << 23);
        if (eb == 0u || eb >= 254u) { k_cur = 0; r_cur = 1.f; }      // all-zero (impossible frame) or non-finite vector: leave it alone
    };
    // maximum over the warp with ONE integer reduction on an order-preserving key instead of five dependent shuffle rounds (a single warp
    // issues in order: the rounds' latencies were step time)
    auto row_max = [&](float e) {
        const unsigned u = __float_as_uint(e);
        const unsigned key = ok ? ((u & 0x80000000u) ? ~u : (u | 0x80000000u)) : 0u;      // NaN-free inputs; absent lanes lowest
        const unsigned mx = __reduce_max_sync(FULL_MASK, key);
        const float m = __uint_as_float((mx & 0x80000000u) ? (mx & 0x7fffffffu) : ~mx);
        return (mx != 0u && m > -INFINITY) ? m : 0.f;
    };
    float w;                            // DIR 0: scaled alpha_t(j);  DIR 1: scaled beta_t(j) * b~_t(j)
    {
        const int f0 = frame(0);
        const float e0 = ok ? __ldg(lb + (size_t)f0 * K + lane) : -INFINITY;
        const float m0 = row_max(e0);
        const float bq = ok ? __expf(e0 - m0) : 0.f;
        if (DIR == 0) {
            w = ok ? __ldg(p.init + lane) * bq : 0.f;
            msum = (double)m0;
            if (ok) ws[(size_t)f0 * K + lane] = w;
            if (lane == 0) wl[f0] = (float)msum;
        } else {
            w = bq;                                                   // beta_{T-1} = 1
            if (ok) ws[(size_t)f0 * K + lane] = 1.f;
            if (lane == 0) wl[f0] = 0.f;                              // exclusive: nothing divided out of beta_{T-1}
            msum = (double)m0;
        }
        set_scale(w);
    }
    // the step proper: acc = sum_i prev(i) * slice(i) (two accumulators), emission, scaling, outputs
    // (m, bq) = the frame's maximum log-emission and this lane's exp(e - m): computed here from e, or handed in by the bulk-staged
    // form, which prepares a whole chunk at once (one lane per frame) so that neither the 5-round shuffle maximum nor the exponential
    // sits in the serial instruction stream of the step -- a single warp issues in order, every exposed latency is step time
    auto step_core = [&](int n, const float (&cv)[KP], float e, bool prepared, float m_in) {
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int i = 0; i < KP; i += 2) {
            a0 = fmaf(__shfl_sync(FULL_MASK, w, i), cv[i], a0);
            a1 = fmaf(__shfl_sync(FULL_MASK, w, i + 1), cv[i + 1], a1);
        }
        const float acc = a0 + a1;
        const int f = frame(n);
        const float m = prepared ? m_in : row_max(e);
        const float bq = prepared ? e : (ok ? __expf(e - m) : 0.f);
        ksum += k_cur;
        const float beta = acc * r_cur;                               // DIR 1: scaled beta_f
        w = acc * (bq * r_cur);
        if (DIR == 0) {
            msum += (double)m;
            if (ok) ws[(size_t)f * K + lane] = w;
            if (lane == 0) wl[f] = (float)(msum + 0.69314718055994530942 * (double)ksum);
        } else {
            if (ok) ws[(size_t)f * K + lane] = beta;
            if (lane == 0) wl[f] = (float)(msum + 0.69314718055994530942 * (double)ksum);   // maxima of frames f+1 .. T-1 only
            msum += (double)m;
        }
        set_scale(w);
    };
    auto step = [&](int n, auto slot_tag) {
        constexpr int slot = decltype(slot_tag)::value;     // compile-time: the prefetch slots stay in registers
        float cv[KP];
#pragma unroll
        for (int i = 0; i < KP; ++i) cv[i] = col[slot][i];
        const float e = eraw[slot];
        {   // refill the slot two steps ahead (off the chain)
            const int n2 = n + TV_PF;
            load_slice(n2 < T ? frame(n2) : -1, col[slot]);
            eraw[slot] = (ok && n2 < T) ? __ldg(lb + (size_t)frame(n2) * K + lane) : -INFINITY;
        }
        step_core(n, cv, e, false, 0.f);
    };
    static_assert(TV_PF == 2, "the unrolled step pairs below assume two prefetch slots");
    if (BULK) {
        for (int c = 0; c < n_chunks; ++c) {
            __syncwarp();                                    // every lane is done with the other buffer (chunk c - 1)
            if (lane == 0 && c + 1 < n_chunks) chunk_issue(c + 1);
            stg.wait(c);
            const int n0 = 1 + c * p.chf, cnt = min(p.chf, T - n0);
            const float *S = stg.tr(c), *E = stg.em(c);
#pragma unroll 2
            for (int q = 0; q < cnt; ++q) {
                const int li = (DIR == 0) ? q : cnt - 1 - q;
                const float *sl = S + (size_t)li * K * K;
                float cv[KP];
                if (DIR == 0) {                              // lane j: column j
#pragma unroll
                    for (int i = 0; i < KP; ++i) cv[i] = (ok && i < K) ? sl[i * K + lane] : 0.f;
                } else {                                     // lane i: row i (16-byte pieces: K % 4 == 0)
#pragma unroll
                    for (int i4 = 0; i4 < KP / 4; ++i4) {
                        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (ok && 4 * i4 < K) v = reinterpret_cast<const float4 *>(sl + lane * K)[i4];
                        cv[4 * i4] = v.x; cv[4 * i4 + 1] = v.y; cv[4 * i4 + 2] = v.z; cv[4 * i4 + 3] = v.w;
                    }
                }
                step_core(n0 + q, cv, ok ? E[li * K + lane] : -INFINITY, false, 0.f);
            }
        }
    } else
    for (int n = 1; n < T; n += TV_PF) {                    // frame n uses slot (n - 1) % TV_PF
        step(n, std::integral_constant<int, 0>{});
        if (n + 1 < T) step(n + 1, std::integral_constant<int, 1>{});
    }
    if (DIR == 0 && p.loglik != nullptr) {
        float tot = ok ? w : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(FULL_MASK, tot, o);
        if (lane == 0) p.loglik[b] = (float)(msum + 0.69314718055994530942 * (double)ksum + (double)logf(tot));
    }
}

template <int KP, bool BULK>
__global__ void __launch_bounds__(32) tv_sweep_kernel(TvParams p) {
    extern __shared__ __align__(16) float tv_stage_sm[];
    __shared__ uint64_t tv_stage_bar[2];
    if (blockIdx.y == 0) tv_sweep<KP, 0, BULK>(p, tv_stage_sm, tv_stage_bar);
    else tv_sweep<KP, 1, BULK>(p, tv_stage_sm, tv_stage_bar);
}

// delta_t(j) = max_i(delta_{t-1}(i) + logP_{t-1}(i,j)) + log b_t(j), first-index ties (torch.max), psi_0 = 0      (neural.py:487-499)
template <int KP, bool BULK>
__global__ void __launch_bounds__(32) tv_viterbi_kernel(TvParams p) {
    extern __shared__ __align__(16) uint8_t tv_smem[];
    __shared__ uint64_t tv_vbar[2];
    const int b = blockIdx.x, lane = threadIdx.x;
    const int K = p.K, T = p.T;
    const bool ok = lane < K;
    const float *lb = p.logb + (size_t)b * T * K;
    const float *tr = p.trans + (size_t)b * T * K * K;
    // dynamic shared memory: [staging buffers (bulk-staged feed)][backpointers (when they fit)]
    const size_t stage_bytes = BULK ? (size_t)2 * p.chf * (K * K + K) * sizeof(float) : 0;
    uint8_t *psi = p.psi_in_smem ? tv_smem + stage_bytes : p.psi_ws + (size_t)b * T * K;
    TvStage stg(reinterpret_cast<float *>(tv_smem), tv_vbar, p.chf, K);
    const int n_chunks = BULK ? (T - 1 + p.chf - 1) / p.chf : 0;
    auto chunk_issue = [&](int c) {                                    // steps t0 .. t0+cnt-1 use slices t0-1 .. and emission rows t0 ..
        const int t0 = 1 + c * p.chf, cnt = min(p.chf, T - t0);
        stg.issue(c, tr + (size_t)(t0 - 1) * K * K, lb + (size_t)t0 * K, cnt);
    };
    if (BULK) {
        stg.init();
        if (lane == 0 && n_chunks > 0) chunk_issue(0);
    }
    auto load_slice = [&](int t, float (&c)[KP]) {                     // slice t-1 feeds frame t
        const bool v = ok && t >= 1 && t < T;
        const float *s = tr + (size_t)(v ? t - 1 : 0) * K * K;
#pragma unroll
        for (int i = 0; i < KP; ++i) c[i] = (v && i < K) ? __ldg(s + i * K + lane) : -INFINITY;
    };
    float col[TV_PF][KP], eraw[TV_PF];
    if (!BULK) {
#pragma unroll
        for (int q = 0; q < TV_PF; ++q) {
            load_slice(1 + q, col[q]);
            eraw[q] = (ok && 1 + q < T) ? __ldg(lb + (size_t)(1 + q) * K + lane) : -INFINITY;
        }
    }
    float d = ok ? __fadd_rn(__ldg(p.init + lane), __ldg(lb + lane)) : -INFINITY;
    if (ok && p.delta) p.delta[(size_t)b * T * K + lane] = d;
    if (ok) psi[lane] = 0;
    if (ok && p.psi_out) p.psi_out[(size_t)b * T * K + lane] = 0;
    auto step_core = [&](int t, const float (&cv)[KP], float e) {
        float best = -INFINITY;
        int arg = 0;
#pragma unroll
        for (int i = 0; i < KP; ++i) {
            const float c = __fadd_rn(__shfl_sync(FULL_MASK, d, i), cv[i]);
            if (c > best) { best = c; arg = i; }                        // strict '>': the lowest index wins a tie
        }
        d = ok ? __fadd_rn(best, e) : -INFINITY;
        if (ok) {
            if (p.delta) p.delta[((size_t)b * T + t) * K + lane] = d;
            psi[(size_t)t * K + lane] = (uint8_t)arg;
            if (p.psi_out) p.psi_out[((size_t)b * T + t) * K + lane] = (uint8_t)arg;
        }
    };
    auto step = [&](int t, auto slot_tag) {
        constexpr int slot = decltype(slot_tag)::value;
        float cv[KP];
#pragma unroll
        for (int i = 0; i < KP; ++i) cv[i] = col[slot][i];
        const float e = eraw[slot];
        load_slice(t + TV_PF, col[slot]);
        eraw[slot] = (ok && t + TV_PF < T) ? __ldg(lb + (size_t)(t + TV_PF) * K + lane) : -INFINITY;
        step_core(t, cv, e);
    };
    if (BULK) {
        for (int c = 0; c < n_chunks; ++c) {
            __syncwarp();                                              // every lane is done with the other buffer
            if (lane == 0 && c + 1 < n_chunks) chunk_issue(c + 1);
            stg.wait(c);
            const int t0 = 1 + c * p.chf, cnt = min(p.chf, T - t0);
            const float *S = stg.tr(c), *E = stg.em(c);
#pragma unroll 2
            for (int q = 0; q < cnt; ++q) {
                const float *sl = S + (size_t)q * K * K;
                float cv[KP];
#pragma unroll
                for (int i = 0; i < KP; ++i) cv[i] = (ok && i < K) ? sl[i * K + lane] : -INFINITY;
                step_core(t0 + q, cv, ok ? E[q * K + lane] : -INFINITY);
            }
        }
    } else
    for (int t = 1; t < T; t += TV_PF) {
        step(t, std::integral_constant<int, 0>{});
        if (t + 1 < T) step(t + 1, std::integral_constant<int, 1>{});
    }
    // final state: first index of the maximum (neural.py:503), then the backtrack (neural.py:505-506)
    float bv = d; int bi = lane;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(FULL_MASK, bv, o);
        const int oi = __shfl_xor_sync(FULL_MASK, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (!p.psi_in_smem) __threadfence_block();
    __syncwarp();
    if (lane == 0) {
        if (p.score) p.score[b] = bv;
        int s = (bi < K) ? bi : 0;
        int64_t *st = p.states + (size_t)b * T;
        st[T - 1] = s;
        for (int t = T - 1; t >= 1; --t) { s = psi[(size_t)t * K + s]; st[t - 1] = s; }
    }
}

#define TV_DISPATCH(FN, ...)                                  \
    do {                                                      \
        const int kp_ = pad4(K);                              \
        if (kp_ <= 4) { FN<4, false> __VA_ARGS__; }           \
        else if (kp_ <= 8) { FN<8, false> __VA_ARGS__; }      \
        else if (kp_ <= 12) { FN<12, false> __VA_ARGS__; }    \
        else if (kp_ <= 16) { FN<16, false> __VA_ARGS__; }    \
        else if (kp_ <= 24) { FN<24, false> __VA_ARGS__; }    \
        else { FN<32, false> __VA_ARGS__; }                   \
    } while (0)

// frames per staged chunk of the bulk feed (two buffers of <= 32 KB), or 0 when the tensors do not qualify (K % 4, alignment)
static int tv_chunk_frames(const float *emis, const float *trans, int K) {
    if (K % 4 != 0 || ((((uintptr_t)emis) | ((uintptr_t)trans)) & 15) != 0) return 0;
#ifdef HMMB200_DEBUG_HOOKS
    if (getenv("HMMB200_TV_NO_BULK")) return 0;
#endif
    int chf = 8192 / (K * K + K);
    return chf > 32 ? 32 : (chf < 2 ? 2 : chf);
}

}  // namespace hmmb200

using namespace hmmb200;

HMMB200_EXPORT size_t hmmb200_tv_viterbi_workspace_bytes(int B, int T, int K) {
    if (B <= 0 || T <= 0 || K <= 0 || K > 32) return 0;
    return ((size_t)T * K <= 96 * 1024) ? 0 : (size_t)B * T * K;
}

HMMB200_EXPORT int hmmb200_tv_forward_backward_f32(const float *log_emis, const float *trans_prob, const float *init_prob,
                                                   int B, int T, int K, float *gamma, float *fwd_prob, float *bwd_prob,
                                                   float *log_alpha, float *log_beta, float *loglik,
                                                   void *workspace, size_t workspace_bytes, void *stream) {
    if (B < 0 || T < 0 || K <= 0) return set_error(HMMB200_EINVAL, "tv_forward_backward: bad shape B=%d T=%d K=%d", B, T, K);
    if (B == 0 || T == 0) return HMMB200_OK;
    if (K > 32) return set_error(HMMB200_EUNSUPPORTED, "tv_forward_backward: K <= 32 states supported (got %d)", K);
    if (!log_emis || !trans_prob || !init_prob) return set_error(HMMB200_EINVAL, "tv_forward_backward: null input");
    const size_t need = hmmb200_fb_workspace_bytes(B, T, K);
    if (!workspace || workspace_bytes < need) return set_error(HMMB200_EWORKSPACE, "tv_forward_backward: workspace %zu < %zu bytes", workspace_bytes, need);
    if (int rc = require_sm100()) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    const size_t n = (size_t)B * T;
    uint8_t *w = (uint8_t *)workspace;
    TvParams p = {};
    p.logb = log_emis; p.trans = trans_prob; p.init = init_prob; p.B = B; p.T = T; p.K = K;
    p.ws_a = (float *)w;  w += align256(n * K * sizeof(float));
    p.ws_b = (float *)w;  w += align256(n * K * sizeof(float));
    p.ws_la = (float *)w; w += align256(n * sizeof(float));
    p.ws_lb = (float *)w;
    p.loglik = loglik;
    const bool both = gamma || fwd_prob || bwd_prob || log_alpha || log_beta;
    dim3 grid((unsigned)B, both ? 2 : 1);
    p.chf = tv_chunk_frames(log_emis, trans_prob, K);
    if (p.chf > 0) {
        const size_t smem = (size_t)2 * p.chf * (K * K + K) * sizeof(float);
        const int kp = pad4(K);
#define TV_FB_CASE(N)                                                                                                                 \
        if (kp <= N) {                                                                                                                \
            if (smem > 48 * 1024) cudaFuncSetAttribute(tv_sweep_kernel<N, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
            tv_sweep_kernel<N, true><<<grid, 32, smem, s>>>(p);                                                                        \
        } else
        TV_FB_CASE(4) TV_FB_CASE(8) TV_FB_CASE(12) TV_FB_CASE(16) TV_FB_CASE(24) TV_FB_CASE(32) { }
#undef TV_FB_CASE
    } else {
        TV_DISPATCH(tv_sweep_kernel, <<<grid, 32, 0, s>>>(p));
    }
    if (int rc = check_launch("tv_sweep_kernel")) return rc;
    if (both) {
        CombineParams c;
        c.ws_a = p.ws_a; c.ws_b = p.ws_b; c.ws_la = p.ws_la; c.ws_lb = p.ws_lb;
        c.n_frames = (int64_t)n; c.K = K;
        c.gamma = gamma; c.fwd = fwd_prob; c.bwd = bwd_prob; c.log_alpha = log_alpha; c.log_beta = log_beta; c.bf16 = 0;
        return launch_combine(c, s);
    }
    return HMMB200_OK;
}

HMMB200_EXPORT int hmmb200_tv_viterbi_f32(const float *log_emis, const float *log_trans, const float *log_init,
                                          int B, int T, int K, float *delta, void *psi_out, int64_t *states, float *score,
                                          void *workspace, size_t workspace_bytes, void *stream) {
    if (B < 0 || T < 0 || K <= 0) return set_error(HMMB200_EINVAL, "tv_viterbi: bad shape B=%d T=%d K=%d", B, T, K);
    if (B == 0 || T == 0) return HMMB200_OK;
    if (K > 32) return set_error(HMMB200_EUNSUPPORTED, "tv_viterbi: K <= 32 states supported (got %d)", K);
    if (!log_emis || !log_trans || !log_init || !states) return set_error(HMMB200_EINVAL, "tv_viterbi: null argument");
    const size_t need = hmmb200_tv_viterbi_workspace_bytes(B, T, K);
    if (need && (!workspace || workspace_bytes < need)) return set_error(HMMB200_EWORKSPACE, "tv_viterbi: workspace %zu < %zu bytes", workspace_bytes, need);
    if (int rc = require_sm100()) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    TvParams p = {};
    p.logb = log_emis; p.trans = log_trans; p.init = log_init; p.B = B; p.T = T; p.K = K;
    p.delta = delta; p.psi_out = (uint8_t *)psi_out; p.states = states; p.score = score;
    p.psi_ws = (uint8_t *)workspace; p.psi_in_smem = need == 0 ? 1 : 0;
    p.chf = tv_chunk_frames(log_emis, log_trans, K);
    const size_t stage = p.chf > 0 ? (size_t)2 * p.chf * (K * K + K) * sizeof(float) : 0;
    const size_t smem = stage + (p.psi_in_smem ? (size_t)T * K : 0);
    const int kp = pad4(K);
#define TV_VIT_CASE(N)                                                                                                         \
    if (kp <= N) {                                                                                                             \
        if (p.chf > 0) {                                                                                                       \
            if (smem > 48 * 1024) cudaFuncSetAttribute(tv_viterbi_kernel<N, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
            tv_viterbi_kernel<N, true><<<(unsigned)B, 32, smem, s>>>(p);                                                       \
        } else {                                                                                                               \
            if (smem > 48 * 1024) cudaFuncSetAttribute(tv_viterbi_kernel<N, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
            tv_viterbi_kernel<N, false><<<(unsigned)B, 32, smem, s>>>(p);                                                      \
        }                                                                                                                      \
        return check_launch("tv_viterbi_kernel");                                                                              \
    }
    TV_VIT_CASE(4) TV_VIT_CASE(8) TV_VIT_CASE(12) TV_VIT_CASE(16) TV_VIT_CASE(24)
    TV_VIT_CASE(32)
#undef TV_VIT_CASE
    return HMMB200_OK;
}
