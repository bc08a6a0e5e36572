// recursion_largek.cu -- HMM recursions for 32 < K <= 512 states on sm_100a: one thread-block CLUSTER per group of
// sequences, the transition matrix resident in registers, the state vector exchanged through distributed shared memory.
//
//   lk_sweep_kernel<FWD/BWD>  scaled-probability forward / backward sweeps      (pytorch_hmm/hmm.py:95-117)
//   lk_sweep_kernel<VIT>      max-plus recursion, bit-identical delta           (hmm.py:159-168)
//   lk_traceback_kernel       arg-max + backpointers recomputed on the path     (hmm.py:167, :174-178)
//   lk_combine_kernel         posterior / exp(log alpha) / exp(log beta)        (hmm.py:120-128)
//   lk_rowmax_kernel          per-frame max_k of the log-emissions (scaling of the LOG emission modes)
//   lk_logscale_kernel        integer scale exponents of a sweep -> log units (double), log-likelihood
//
// Why a cluster: one step is a [n_seq, K] x [K, K] product followed by a dependency on ALL K results, T times.  A K = 512
// fp32 matrix is 1 MB: it fits in no single SM, and re-reading it from L2 every step (1 MB x 4000 steps x 64 sequences)
// would make the sweep L2-bound.  So the CS CTAs of a cluster each own NC = 64 output states: their [K, 64] slab of P lives
// in the REGISTERS of 8 compute warps (128 per thread: warp w holds source states 64w..64w+63, each half-warp 32 of them,
// each lane 4 output states), read from HBM exactly once.  Per step a compute warp multiplies its slice of the previous
// vector (broadcast LDS.128 from shared memory) into the slab with packed FFMA2; 8 final warps add the 16 partial sums,
// apply emission and scaling, and push the CTA's new 64-state block to every CTA of the cluster with ONE bulk DSMEM copy
// per destination (cp.async.bulk + mbarrier complete_tx): the receiver's mbarrier flips exactly when all CS blocks have
// landed, so there is no cluster-wide barrier on the critical path, only the data's own arrival.  The vector is
// double-buffered; a sender can never run more than one step ahead of a receiver because it needs the receiver's block to
// do so.  Two independent groups of sequences alternate on the slab so that one group's finals and exchange overlap the
// other group's product (see lk_sweep_kernel).
//
// Scaling (forward/backward): every step is multiplied by 2^-k, k = exponent of the previous vector's largest entry
// (each CTA ships its local maxima with its slice), exact power-of-two scaling with integer bookkeeping -- the same
// scheme as the small-K kernels (recursion_smallk.cu), so alpha_t = w_t * 2^ksum_t * exp(sum m).
// Viterbi: the max over source states is exact and order-independent, then ONE fp32 add of log b -- the reference's two
// roundings (hmm.py:164-168) -- so delta is bit-identical; backpointers are not stored for all K states (K x T x B bytes
// nobody reads): the traceback kernel recomputes argmax_i(delta_{t-1}(i) + logP(i, s_t)) with the same fp32 add and the
// lowest-index tie rule only for the states on the path.
#include "common.cuh"

#include <cooperative_groups.h>
#include <stdio.h>
#include <stdlib.h>

namespace cg = cooperative_groups;

namespace hmmb200 {

constexpr int LK_NSQ_MAX = 4;                   // sequences per group: 4, or 3 when that spreads one wave of clusters over more SMs
constexpr int LK_NG = 4;                        // independent sequence groups per cluster at most (software-pipelined); LkParams::ngr are used
constexpr int LK_NC = 64;                       // output states per CTA
constexpr int LK_KS = 64;                       // source states per compute warp (k-slice) = one CTA's block of the vector
constexpr int LK_NWC = 8;                       // compute warps per CTA (warps 0-7); warps 8-15 are the final warps
constexpr int LK_KH = LK_KS / 2;                // source states per half-warp
constexpr int LK_NSL = 2 * LK_NWC;              // k-slices of partial sums: one per compute half-warp
constexpr int LK_NW = 2 * LK_NWC;               // warps per CTA
constexpr int LK_THREADS = LK_NW * 32;          // 512
constexpr int LK_KMAX = LK_NWC * LK_KS;         // 512
constexpr int LK_REGS_COMPUTE = 184;            // setmaxnreg: 256 x 184 + 256 x 72 = 64 K registers
constexpr int LK_REGS_FINAL = 72;
constexpr int LK_BAR_PART = 1;                  // named barriers 1 + g: partial sums of group g are in shared memory
constexpr int LK_BAR_FINAL = 1 + LK_NG;         // the 256 final threads, between staging and the push
constexpr int LK_BLK = LK_NC + 4;               // floats per (CTA, sequence) block: 64 states + 2 local maxima + pad (272 B)
constexpr int LK_CSMAX = LK_KMAX / LK_NC;       // 8 CTAs per cluster at most (portable cluster size)
constexpr int LK_FINAL = LK_NSQ_MAX * LK_NC;    // 256 final threads: (sequence, output state); the last 64 idle when NSQ = 3
constexpr int LK_PF = 8;                        // emission prefetch distance in steps

enum { LK_FWD = 0, LK_BWD = 1, LK_VIT = 2 };

struct LkParams {
    const float *emis;     // [B,T,K]
    int mode;
    float eps;
    int add_rowmax;
    const float *trans;    // [K,K]  fb: effective probabilities; viterbi: log transitions
    const float *init;     // [K]    fb: probabilities; viterbi: log
    const float *rowmax;   // [B,T] per-frame max of the log-emissions (LOG emission modes) or null
    int B, T, K, CS;
    int ngr;               // sequence groups per cluster in this launch (2 for a sweep on its own; more when several sweeps share the GPU)
    float *ws_a, *ws_b;    // [B,T,K] scaled alpha / beta (fb)
    float *ws_la, *ws_lb;  // [B,T]
    float *loglik;         // [B] or null
    float *delta;          // [B,T,K] (viterbi)
    int *err;              // device flag: set to 1 if an exchange wait times out (never in a correct run)
    int trace;
};

__device__ __forceinline__ uint32_t lk_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float2 lk_ffma2(float2 a, float2 b, float2 c) {
    unsigned long long ra = *reinterpret_cast<unsigned long long *>(&a), rb = *reinterpret_cast<unsigned long long *>(&b);
    unsigned long long rc = *reinterpret_cast<unsigned long long *>(&c), rd;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2 *>(&rd);
}
__device__ __forceinline__ float2 lk_fadd2(float2 a, float2 b) {
    unsigned long long ra = *reinterpret_cast<unsigned long long *>(&a), rb = *reinterpret_cast<unsigned long long *>(&b), rd;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    return *reinterpret_cast<float2 *>(&rd);
}
__device__ __forceinline__ float lk_fmax3(float a, float b, float c) {
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
__device__ __forceinline__ uint32_t lk_mapa(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
// bulk copy shared::cta -> (remote) shared::cluster, completion as transaction bytes on the receiver's mbarrier
__device__ __forceinline__ void lk_bulk_push(uint32_t rdst, uint32_t src, uint32_t bytes, uint32_t rbar) {
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(rdst), "r"(src), "r"(bytes), "r"(rbar) : "memory");
}
__device__ __forceinline__ void lk_mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(lk_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void lk_mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(lk_smem_u32(bar)), "r"(bytes) : "memory");
}
// returns false if the phase did not complete within ~4 s (a protocol bug; never in a correct run)
__device__ __forceinline__ bool lk_mbar_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t a = lk_smem_u32(bar);
    uint32_t done = 0;
    for (int spin = 0; spin < (1 << 22); ++spin) {
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(a), "r"(parity), "r"(1000u) : "memory");
        if (done) return true;
    }
    return false;
}
__device__ __forceinline__ void lk_cp_async4(float *dst_smem, const float *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(lk_smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void lk_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void lk_cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// A group holds NSQ x NP sequences: the product and the finals take them in NP passes of NSQ (the register accumulators of a pass
// are the limit), the exchange moves the whole group at once.  With NP = 2 there are half as many groups (same footprint).
template <int NSQ, int NP>
struct LkSmem {
    static constexpr int NSQT = NSQ * NP, NGM = LK_NG / NP;
    float vec[NGM][2][LK_CSMAX][NSQT][LK_BLK];    // the exchanged state vector: one block per source CTA, double-buffered
    float part[NGM][LK_NSL][NSQT][LK_NC];         // per-k-slice partial sums, one buffer per group
    float stage[NGM][2][NSQT][LK_BLK];            // this CTA's new block before it is pushed (double-buffered)
    float eraw[NGM][LK_PF][NP][LK_FINAL];         // prefetched emissions, one slot per (pass, final thread)
    float mraw[NGM][LK_PF][NP][LK_FINAL];         // prefetched per-frame max
    uint64_t bar[NGM][2];
};

// timing trace (debug builds only, -DHMMB200_DEBUG_HOOKS; tools/lk_trace.py): clock64 at the phase boundaries of steps 64..71 of
// CTA 0, thread 0, group 0.  Release builds carry no device globals, no environment look-ups and no debug exports.
#ifdef HMMB200_DEBUG_HOOKS
__device__ long long lk_trace_buf[8 * 8];
#define LK_TRACE(slot)                                                                          \
    do {                                                                                        \
        if (p.trace && g == 0 && blockIdx.x == 0 && (tid & (LK_FINAL - 1)) == 0 && t >= 64 && t < 72) lk_trace_buf[(t - 64) * 8 + (slot)] = clock64(); \
    } while (0)
#else
#define LK_TRACE(slot) do { } while (0)
#endif

// One cluster = CS CTAs x LK_NG groups of NSQ (3 or 4) sequences.  The groups are independent recursions that share the CTA's
// register-resident slab of P.  Warp roles (register budgets moved between them with setmaxnreg):
//   warps 0-7  COMPUTE: warp w keeps P[64w..64w+63][the CTA's 64 output states] (128 registers per thread: 4 output
//              states x 32 source states per lane) and, per
//              (step, group), multiplies CTA w's block of the previous vector into it -- nothing else;
//   warps 8-15 FINAL: one thread per (sequence, output state) adds the 16 k-slice partials, applies emission and scaling,
//              stages the CTA's new block and pushes it to the cluster.
// While the final warps finish group g, the compute warps are already on group g+1, and group g's exchange flies during
// both: the FMA-bound phase and the latency-bound phase overlap instead of alternating.  (The first version used all 16
// warps for the product and half of them for each group's finals: 1 800-cycle finals serialised with 1 150-cycle products.)
__device__ __forceinline__ void lk_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void lk_bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

template <int MODE, int NSQ, int NP>
__device__ __forceinline__ void lk_sweep_body(const LkParams &p, const int cluster_id) {
    extern __shared__ __align__(16) uint8_t lk_smem_raw[];
    using Smem = LkSmem<NSQ, NP>;
    constexpr int NSQT = Smem::NSQT, NGM = Smem::NGM;
    Smem &sm = *reinterpret_cast<Smem *>(lk_smem_raw);
    constexpr bool VIT = (MODE == LK_VIT);
    constexpr int DIR = (MODE == LK_BWD) ? 1 : 0;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int K = p.K, T = p.T, B = p.B, CS = p.CS;
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int seq0 = cluster_id * (NSQT * p.ngr);
    const int col0 = rank * LK_NC;
    const float PADV = VIT ? -INFINITY : 0.f;
    const int n_groups = min(min(p.ngr, NGM), (B - seq0 + NSQT - 1) / NSQT);   // groups of this cluster that hold sequences

    for (int i = tid; i < NGM * 2 * LK_CSMAX * NSQT * LK_BLK; i += LK_THREADS) (&sm.vec[0][0][0][0][0])[i] = 0.f;   // unused blocks stay 0
    for (int i = tid; i < NGM * 2 * NSQT * LK_BLK; i += LK_THREADS) (&sm.stage[0][0][0][0])[i] = 0.f;
    for (int i = tid; i < NGM * LK_NSL * NSQT * LK_NC; i += LK_THREADS) (&sm.part[0][0][0][0])[i] = PADV;   // idle k-slices: neutral element
    // bytes every receiver gets per step and group: one [NSQ][BLK] block from each of the CS CTAs
    constexpr uint32_t BLOCK_BYTES = NSQT * LK_BLK * sizeof(float);
    const uint32_t tx_bytes = (uint32_t)(CS - 1) * BLOCK_BYTES;   // (the CTA's own block is written in place by its final warps)
    if (tid == 0) {
#pragma unroll
        // a phase = the arming thread's expect_tx + the 8 final warps of this CTA (own block) + the bytes of the other CTAs' blocks
        for (int g = 0; g < NGM; ++g) { lk_mbar_init(&sm.bar[g][0], 1 + LK_NWC); lk_mbar_init(&sm.bar[g][1], 1 + LK_NWC); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        // step 0 is armed here: the final warps push it without waiting for the compute warps (no product at t = 0),
        // so nothing else orders this expect_tx before their complete_tx
#pragma unroll
        for (int g = 0; g < NGM; ++g)
            if (g < n_groups) lk_mbar_expect_tx(&sm.bar[g][0], tx_bytes);
    }
    cluster.sync();                                          // every CTA's mbarriers exist (and are armed) before anyone pushes
    bool ok = true;

    if (warp < LK_NWC) {
        // ================================ COMPUTE warps ================================================================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(LK_REGS_COMPUTE));
        // forward / viterbi: out(j) = sum_i v(i) M(i,j)  -> M[i][j];   backward: out(i) = sum_j M(i,j) v(j) -> M[out][src]
        // Lane (hw, c): FOUR output states 4c..4c+3 x the 32 source states of half hw of the warp's 64 (128 registers).
        // One LDS.128 of the vector then feeds 8 packed FMAs -- with two output states per lane the product was bound
        // by the shared-memory return path (a broadcast LDS.128 still delivers 512 B per warp), not by the FMA pipe.
        const int hw = lane >> 4, c4 = (lane & 15) * 4;
        float2 Pj[4][LK_KH / 2];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int out = col0 + c4 + j;
#pragma unroll
            for (int i = 0; i < LK_KH / 2; ++i) {
                const int s0 = warp * LK_KS + hw * LK_KH + 2 * i, s1 = s0 + 1;
                auto ld = [&](int src) -> float {
                    if (src >= K || out >= K) return PADV;
                    return (DIR == 0) ? __ldg(p.trans + (size_t)src * K + out) : __ldg(p.trans + (size_t)out * K + src);
                };
                Pj[j][i] = make_float2(ld(s0), ld(s1));
            }
        }
        const bool active = warp < CS;                       // source states 64w.. exist
        for (int t = 0; t < T; ++t) {
            const int cur = t & 1, prv = cur ^ 1;
#pragma unroll 1                                             // (one copy of the product code: the groups differ by shared-memory offsets only)
            for (int g = 0; g < n_groups; ++g) {
                LK_TRACE(0);
                if (t == 0) continue;                        // step 0 has no product (and was armed at set-up)
                if (tid == 0) lk_mbar_expect_tx(&sm.bar[g][cur], tx_bytes);
                // (idle warps wait too: nobody may arrive on the named barrier twice within one of its phases)
                if (ok) ok = lk_mbar_wait(&sm.bar[g][prv], ((t - 1) >> 1) & 1);   // after a time-out: drain without waiting
                LK_TRACE(1);
                if (active) {
#pragma unroll 1                                             // (one copy of the product code: the passes differ by shared-memory offsets only)
                  for (int h = 0; h < NP; ++h) {
                    const float *v = &sm.vec[g][prv][warp][h * NSQ][hw * LK_KH];       // states 64w.. = CTA w's block
                    float *po = &sm.part[g][2 * warp + hw][h * NSQ][c4];
                    if (!VIT) {
                        float2 acc[NSQ][4];
#pragma unroll
                        for (int s = 0; s < NSQ; ++s)
#pragma unroll
                            for (int j = 0; j < 4; ++j) acc[s][j] = make_float2(0.f, 0.f);
#pragma unroll
                        for (int kk = 0; kk < LK_KH; kk += 4) {
#pragma unroll
                            for (int s = 0; s < NSQ; ++s) {
                                const float4 x = *reinterpret_cast<const float4 *>(v + s * LK_BLK + kk);
#pragma unroll
                                for (int j = 0; j < 4; ++j) acc[s][j] = lk_ffma2(make_float2(x.x, x.y), Pj[j][kk / 2], acc[s][j]);
#pragma unroll
                                for (int j = 0; j < 4; ++j) acc[s][j] = lk_ffma2(make_float2(x.z, x.w), Pj[j][kk / 2 + 1], acc[s][j]);
                            }
                        }
#pragma unroll
                        for (int s = 0; s < NSQ; ++s)
                            *reinterpret_cast<float4 *>(po + s * LK_NC) = make_float4(acc[s][0].x + acc[s][0].y, acc[s][1].x + acc[s][1].y,
                                                                                      acc[s][2].x + acc[s][2].y, acc[s][3].x + acc[s][3].y);
                    } else {
                        float m[NSQ][4];
#pragma unroll
                        for (int s = 0; s < NSQ; ++s)
#pragma unroll
                            for (int j = 0; j < 4; ++j) m[s][j] = -INFINITY;
#pragma unroll
                        for (int kk = 0; kk < LK_KH; kk += 4) {
#pragma unroll
                            for (int s = 0; s < NSQ; ++s) {
                                const float4 x = *reinterpret_cast<const float4 *>(v + s * LK_BLK + kk);
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    const float2 c0 = lk_fadd2(make_float2(x.x, x.y), Pj[j][kk / 2]);
                                    const float2 c1 = lk_fadd2(make_float2(x.z, x.w), Pj[j][kk / 2 + 1]);
                                    m[s][j] = lk_fmax3(m[s][j], c0.x, c0.y);
                                    m[s][j] = lk_fmax3(m[s][j], c1.x, c1.y);
                                }
                            }
                        }
#pragma unroll
                        for (int s = 0; s < NSQ; ++s)
                            *reinterpret_cast<float4 *>(po + s * LK_NC) = make_float4(m[s][0], m[s][1], m[s][2], m[s][3]);
                    }
                  }
                    LK_TRACE(2);
                }
                // (part[g] is free again by then: group g's next exchange, which the wait above needs, is pushed after its finals)
                lk_bar_arrive(LK_BAR_PART + g, LK_THREADS);
            }
        }
    } else {
        // ================================ FINAL warps: thread = (sequence fs of the group, output state fc) ============
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(LK_REGS_FINAL));
        const int ft = tid - LK_FINAL, fwarp = ft >> 5;
        const bool valid = ft < NSQ * LK_NC;                 // (whole warps: NSQ = 3 leaves final warps 6, 7 with the barriers only)
        const int fs = valid ? ft / LK_NC : 0, fc = ft % LK_NC;
        const int gcol = col0 + fc;
        const bool need_m = (p.rowmax != nullptr);
        auto frame_of = [&](int t) { return (DIR == 0) ? t : T - 1 - t; };
        auto prefetch = [&](int t) {
            if (t < T && valid) {
                const int f = frame_of(t);
#pragma unroll
                for (int g = 0; g < NGM; ++g) {
                    if (g >= n_groups) continue;
#pragma unroll
                    for (int h = 0; h < NP; ++h) {
                        const int sq = seq0 + g * NSQT + h * NSQ + fs;
                        const bool okk = sq < B && gcol < K;
                        lk_cp_async4(&sm.eraw[g][t % LK_PF][h][ft], p.emis + ((size_t)(okk ? sq : 0) * T + f) * K + (okk ? gcol : 0));
                        if (need_m) lk_cp_async4(&sm.mraw[g][t % LK_PF][h][ft], p.rowmax + (size_t)(sq < B ? sq : 0) * T + f);
                    }
                }
            }
            lk_cp_commit();
        };
        for (int t = 0; t < LK_PF - 1; ++t) prefetch(t);
        int ks0 = 0, ks1 = 0, ks2 = 0, ks3 = 0;              // running power-of-two exponent per (group, pass) (scalars: the loops are rolled)
        static_assert(LK_NG == 4, "four exponent scalars");
        float *ws_l = (DIR == 0) ? p.ws_la : p.ws_lb;

        for (int t = 0; t < T; ++t) {
            const int cur = t & 1, prv = cur ^ 1;
            prefetch(t + LK_PF - 1);
            lk_cp_wait<LK_PF - 1>();                         // this step's emissions have landed (own slots only)
            const int f = frame_of(t);
#pragma unroll 1
            for (int g = 0; g < n_groups; ++g) {
                if (t > 0) {
                    lk_bar_sync(LK_BAR_PART + g, LK_THREADS);                           // all k-slice partials are in part[g]
                    if (ok) ok = lk_mbar_wait(&sm.bar[g][prv], ((t - 1) >> 1) & 1);     // the previous vector (its maxima) is visible
                }
                LK_TRACE(4);
                float wv_h[NP], pre_h[NP];                   // wv: the value pushed to the cluster
#pragma unroll
                for (int h = 0; h < NP; ++h) {
                    const int gs = h * NSQ + fs;             // sequence within the group
                    const int fseq = seq0 + g * NSQT + gs;
                    const bool f_ok = valid && fseq < B && gcol < K;
                    const float raw = sm.eraw[g][t % LK_PF][h][ft];
                    const float mf = need_m ? sm.mraw[g][t % LK_PF][h][ft] : 0.f;
                    float bq = 0.f, lb = 0.f;                // emission: probability form (fb) / log form (viterbi)
                    if (!VIT) {
                        if (p.mode == HMMB200_EMIS_PROB_FLOOR) bq = raw + p.eps;
                        else if (p.mode == HMMB200_EMIS_LOG_EXP_FLOOR) bq = expf(raw) + p.eps;
                        else bq = expf(raw - mf) + ((p.mode == HMMB200_EMIS_LOG_NORM_FLOOR) ? p.eps : 0.f);
                        if (!f_ok) bq = 0.f;
                    } else {                                 // the reference's formula per input kind
                        if (p.mode == HMMB200_EMIS_LOG) lb = raw;
                        else if (p.mode == HMMB200_EMIS_PROB_FLOOR) lb = logf(raw + p.eps);
                        else if (p.mode == HMMB200_EMIS_LOG_EXP_FLOOR) lb = logf(expf(raw) + p.eps);
                        else lb = logf(expf(raw - mf) + p.eps);
                    }
                    float wv, pre = 0.f;
                    if (!VIT) {
                        float acc, r = 1.f;
                        if (t == 0) {
                            acc = (DIR == 0) ? (f_ok ? __ldg(p.init + gcol) : 0.f) : (f_ok ? 1.f : 0.f);
                        } else {
                            float a4[4];                     // fixed summation order: deterministic
#pragma unroll
                            for (int w = 0; w < LK_NSL; ++w) {
                                const float x = sm.part[g][w][gs][fc];
                                a4[w & 3] = (w < 4) ? x : a4[w & 3] + x;
                            }
                            acc = (a4[0] + a4[1]) + (a4[2] + a4[3]);
                            // power-of-two normaliser from the largest entry of the previous vector (all CTAs' local maxima)
                            float m = 0.f;
#pragma unroll
                            for (int q = 0; q < LK_CSMAX; ++q) {
                                const float2 y = *reinterpret_cast<const float2 *>(&sm.vec[g][prv][q][gs][LK_NC]);
                                m = lk_fmax3(m, y.x, y.y);
                            }
                            const unsigned eb = __float_as_uint(m) >> 23;
                            const int de = (int)eb - 127;
                            const int ki = g * NP + h;
                            if (ki == 0) ks0 += de; else if (ki == 1) ks1 += de; else if (ki == 2) ks2 += de; else ks3 += de;
                            r = __uint_as_float((254u - eb) << 23);
                        }
                        pre = acc * r;                       // beta_t (scaled) for the backward sweep
                        wv = acc * (bq * r);
                    } else {
                        float acc;
                        if (t == 0) {
                            acc = f_ok ? __ldg(p.init + gcol) : -INFINITY;
                        } else {
                            float a4[4];
#pragma unroll
                            for (int w = 0; w < LK_NSL; ++w) {
                                const float x = sm.part[g][w][gs][fc];
                                a4[w & 3] = (w < 4) ? x : fmaxf(a4[w & 3], x);
                            }
                            acc = fmaxf(fmaxf(a4[0], a4[1]), fmaxf(a4[2], a4[3]));
                        }
                        wv = f_ok ? __fadd_rn(acc, lb) : -INFINITY;   // delta_t = max_i(..) + log b_t  (hmm.py:168)
                    }
                    wv_h[h] = wv; pre_h[h] = pre;
                    // ---- stage the block in shared memory for the other CTAs; the CTA's own copy is written in place
                    if (valid) { sm.stage[g][cur][gs][fc] = wv; sm.vec[g][cur][rank][gs][fc] = wv; }
                    if (!VIT && valid) {
                        const float lmax = __uint_as_float(__reduce_max_sync(FULL_MASK, __float_as_uint(wv)));   // wv >= 0
                        if (lane == 0) { sm.stage[g][cur][gs][LK_NC + (fwarp & 1)] = lmax; sm.vec[g][cur][rank][gs][LK_NC + (fwarp & 1)] = lmax; }
                    }
                }
                __syncwarp();
                if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(lk_smem_u32(&sm.bar[g][cur])) : "memory");
                // ---- push: ONE bulk DSMEM copy per OTHER CTA of the cluster for the whole group.  These shared::cta -> shared::cluster
                // copies are what a step costs: they move ~3 bytes per cycle per SM (~27 B/cycle for the whole cluster, all CTAs
                // sending at once) -- 2 600 cycles per group of 4 sequences against 1 100 - 1 400 cycles of product, i.e. 5 400 /
                // 7 900 / 10 100 cycles per step with 2 / 3 / 4 one-pass groups (tools/lk_trace.py).  Four other exchanges were
                // measured and lost (profiles/r02_largek_exchange_experiments.txt).
                LK_TRACE(5);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> async-proxy reads
                lk_bar_sync(LK_BAR_FINAL, LK_FINAL);
                LK_TRACE(6);
                // (the bulk copy is a per-warp uniform-datapath instruction: one destination per final warp)
                if (lane == 0 && fwarp < CS && fwarp != rank)
                    lk_bulk_push(lk_mapa(lk_smem_u32(&sm.vec[g][cur][rank][0][0]), fwarp), lk_smem_u32(&sm.stage[g][cur][0][0]),
                                 BLOCK_BYTES, lk_mapa(lk_smem_u32(&sm.bar[g][cur]), fwarp));
                LK_TRACE(7);
                // ---- results to HBM: after the push, off the step's critical path.  The log scale goes out as the integer
                // exponent; lk_logscale_kernel turns it into log units (and adds the per-frame maxima) in double afterwards.
#pragma unroll
                for (int h = 0; h < NP; ++h) {
                    const int fseq = seq0 + g * NSQT + h * NSQ + fs;
                    const bool f_ok = valid && fseq < B && gcol < K;
                    if (f_ok) {
                        const size_t o = ((size_t)fseq * T + f) * K + gcol;
                        if (VIT) p.delta[o] = wv_h[h];
                        else if (DIR == 0) p.ws_a[o] = wv_h[h];
                        else p.ws_b[o] = pre_h[h];
                    }
                    if (!VIT && rank == 0 && fc == 0 && valid && fseq < B) {
                        const int ki = g * NP + h;
                        ws_l[(size_t)fseq * T + f] = __int_as_float((ki == 0) ? ks0 : ((ki == 1) ? ks1 : ((ki == 2) ? ks2 : ks3)));
                    }
                }
            }
        }
    }

    // ---- epilogue: everybody waits for the last vectors (no CTA may exit while blocks are still in flight to it) ------
#pragma unroll
    for (int g = 0; g < NGM; ++g)
        if (g < n_groups && ok) ok = lk_mbar_wait(&sm.bar[g][(T - 1) & 1], ((T - 1) >> 1) & 1);
    __syncthreads();
    if (MODE == LK_FWD && p.loglik != nullptr && rank == 0 && warp < n_groups * NSQT) {
        const int g = warp / NSQT, s = warp % NSQT, sq = seq0 + warp;
        if (sq < B) {
            float tot = 0.f;
            for (int k = lane; k < K; k += 32) tot += sm.vec[g][(T - 1) & 1][k / LK_NC][s][k % LK_NC];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(FULL_MASK, tot, o);
            if (lane == 0) p.loglik[sq] = logf(tot);         // lk_logscale_kernel adds the last frame's log scale
        }
    }
    if (!ok && p.err != nullptr) atomicExch(p.err, 1);
    cluster.sync();
}

template <int MODE, int NSQ, int NP>
__global__ void __launch_bounds__(LK_THREADS, 1) lk_sweep_kernel(LkParams p) {
    lk_sweep_body<MODE, NSQ, NP>(p, (int)blockIdx.x / p.CS);
}

// Several sweeps of the same batch in ONE launch (forward + backward, or forward + backward + Viterbi): clusters [0, ncl) run the
// forward sweep, [ncl, 2 ncl) the backward sweep, [2 ncl, 3 ncl) the Viterbi recursion.  A B200 holds 15 clusters of 8 CTAs; on its
// own a sweep of B = 64 sequences takes 11 of them, so the sweeps of one call used to run one after another.  With more sequences
// per cluster (LkParams::ngr groups, NP passes per group) the sweeps fit side by side.
template <int NSQ, int NP>
__global__ void __launch_bounds__(LK_THREADS, 1) lk_sweep_multi_kernel(LkParams pf, LkParams pb, LkParams pv, int ncl) {
    const int c = (int)blockIdx.x / pf.CS;
    if (c < ncl) lk_sweep_body<LK_FWD, NSQ, NP>(pf, c);
    else if (c < 2 * ncl) lk_sweep_body<LK_BWD, NSQ, NP>(pb, c - ncl);
    else lk_sweep_body<LK_VIT, NSQ, NP>(pv, c - 2 * ncl);
}

// ----------------------------------------------------------------------------------------------------------------------
// log scales (one warp per sequence): the sweeps leave the running power-of-two exponent k_t as an integer; this turns it
// into log units, la_t = k_t ln 2 + sum of the per-frame maxima divided out so far (inclusive for the forward sweep,
// exclusive for the backward one), in double, and adds the last forward scale to the log-likelihood.
// ----------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) lk_logscale_kernel(float *ws_l, const float *rowmax, int T, int dir, int add_m, float *loglik) {
    const int sq = blockIdx.x, lane = threadIdx.x;
    float *l = ws_l + (size_t)sq * T;
    const float *rm = rowmax + (size_t)sq * T;
    double carry = 0.0;
    for (int t0 = 0; t0 < T; t0 += 32) {
        const int t = t0 + lane;
        const int f = dir ? T - 1 - t : t;
        const double m = (add_m && t < T) ? (double)rm[f] : 0.0;
        double x = m;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double y = __shfl_up_sync(FULL_MASK, x, o);
            if (lane >= o) x += y;
        }
        if (t < T) {
            const double ms = carry + (dir ? x - m : x);
            const float v = (float)(ms + 0.69314718055994530942 * (double)__float_as_int(l[f]));
            l[f] = v;
            if (loglik != nullptr && t == T - 1) loglik[sq] += v;
        }
        carry += __shfl_sync(FULL_MASK, x, 31);
    }
}

// ----------------------------------------------------------------------------------------------------------------------
// per-frame max over K of the log-emissions (one warp per frame)
// ----------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) lk_rowmax_kernel(const float *emis, int64_t n_frames, int K, float *rowmax) {
    const int64_t fr = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (fr >= n_frames) return;
    const float *row = emis + fr * K;
    float m = -INFINITY;
    for (int k = lane; k < K; k += 32) m = fmaxf(m, __ldg(row + k));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(FULL_MASK, m, o));
    if (!(m > -INFINITY)) m = 0.f;                           // all states impossible: keep the frame finite
    if (lane == 0) rowmax[fr] = m;
}

// ----------------------------------------------------------------------------------------------------------------------
// combine (one warp per frame): gamma = a.*b / sum, fwd = a*exp(la), bwd = b*exp(lb)            (hmm.py:120-128)
// ----------------------------------------------------------------------------------------------------------------------
struct LkCombineParams {
    const float *ws_a, *ws_b, *ws_la, *ws_lb;
    int64_t n_frames;
    int K;
    float *gamma, *fwd, *bwd, *log_alpha, *log_beta;
};

__global__ void __launch_bounds__(256) lk_combine_kernel(LkCombineParams p) {
    const int64_t fr = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (fr >= p.n_frames) return;
    const int K = p.K;
    const float *a = p.ws_a + fr * K, *b = p.ws_b + fr * K;
    float Z = 0.f;
    for (int k = lane; k < K; k += 32) Z += a[k] * b[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) Z += __shfl_xor_sync(FULL_MASK, Z, o);
    const float inv = 1.f / Z;
    const float la = p.ws_la[fr], lb = p.ws_lb[fr];
    const float ea = expf(la), eb = expf(lb);
    for (int k = lane; k < K; k += 32) {
        const float x = a[k], y = b[k];
        if (p.gamma) p.gamma[fr * K + k] = x * y * inv;
        if (p.fwd) p.fwd[fr * K + k] = x * ea;
        if (p.bwd) p.bwd[fr * K + k] = y * eb;
        if (p.log_alpha) p.log_alpha[fr * K + k] = logf(x) + la;
        if (p.log_beta) p.log_beta[fr * K + k] = logf(y) + lb;
    }
}

// ----------------------------------------------------------------------------------------------------------------------
// traceback (one warp per sequence): s_{T-1} = first argmax delta_{T-1}; s_{t-1} = first argmax_i(delta_{t-1}(i) +
// logP(i, s_t)) -- the backpointer psi_t[s_t] of hmm.py:167 recomputed with the same fp32 add, lowest index on ties.
// logPT is the transposed log-transition matrix so that the needed column is a contiguous row.
// ----------------------------------------------------------------------------------------------------------------------
__global__ void lk_transpose_kernel(const float *src, int K, float *dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < K * K) dst[(size_t)(i % K) * K + i / K] = src[i];
}

// first arg-max over the warp of (v, i), v never NaN: two warp reductions (REDUX) on an order-preserving integer key instead of five
// shuffle rounds -- the traceback's step is a dependency chain and this is a third of it
__device__ __forceinline__ void lk_warp_argmax(float &v, int &i) {
    const unsigned u = __float_as_uint(v + 0.f);                       // (-0 -> +0: equal values must get equal keys)
    const unsigned key = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    const unsigned mx = __reduce_max_sync(FULL_MASK, key);
    i = (int)__reduce_min_sync(FULL_MASK, (key == mx) ? (unsigned)i : 0xffffffffu);
    v = __uint_as_float((mx & 0x80000000u) ? (mx & 0x7fffffffu) : ~mx);
}

constexpr int LK_TB_PF = 6;                                  // delta rows prefetched into L2 ahead of the traceback
// MAXR: delta-row registers per lane (K <= 32 MAXR); ALIGNED: K is a multiple of 32, so a 32-state slice is either whole or absent
// (uniform predicates instead of per-lane index clamps: the step is a dependency chain and its instruction count is its time)
template <int MAXR, bool ALIGNED>
__global__ void __launch_bounds__(128) lk_traceback_kernel(const float *delta, const float *logPT, int B, int T, int K,
                                                           int64_t *states, float *score) {
    const int sq = blockIdx.x * 4 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (sq >= B) return;
    const float *d = delta + (size_t)sq * T * K;
    float bv = -INFINITY;
    int bi = K;
    for (int k = lane; k < K; k += 32) {
        const float v = d[(size_t)(T - 1) * K + k];
        if (v > bv) { bv = v; bi = k; }
    }
    if (bi == K) bi = lane < K ? lane : 0;                   // all -inf / NaN row: still a valid index
    lk_warp_argmax(bv, bi);
    int s = bi;
    if (lane == 0) {
        states[(size_t)sq * T + (T - 1)] = s;
        if (score) score[sq] = bv;
    }
    const int mk = K >> 5;                                   // whole slices
    auto have = [&](int m) { return ALIGNED ? (m < mk) : (lane + 32 * m < K); };
    auto at = [&](int m) { return ALIGNED ? lane + 32 * m : min(lane + 32 * m, K - 1); };
    float nxt[MAXR];                                         // delta row of the next step; -inf where the lane has no state
    const float *dn = d + (size_t)(T >= 2 ? T - 2 : 0) * K;
#pragma unroll
    for (int m = 0; m < MAXR; ++m) nxt[m] = (T >= 2 && have(m)) ? dn[at(m)] : -INFINITY;
    int64_t *st = states + (size_t)sq * T;
    for (int t = T - 1; t >= 1; --t) {
        const float *col = logPT + (size_t)s * K;
        // The loads of the logPT row first -- they are the step's critical path -- and none inside a per-lane branch: with the load
        // under `if (i < K)` the compiler issued them one by one, each behind the previous compare -- 16 L2 round trips per step
        // instead of one (ncu source page: 4.4 ms for T = 4000, 90 % of it on the 16 dependent FADDs; 2.3 ms with the loads batched)
        float cv[MAXR];
#pragma unroll
        for (int m = 0; m < MAXR; ++m) cv[m] = (!ALIGNED || m < mk) ? __ldg(col + at(m)) : 0.f;
        float cur[MAXR];
#pragma unroll
        for (int m = 0; m < MAXR; ++m) cur[m] = nxt[m];
        dn -= K;
        if (t >= 2) {                                        // the next delta row does not depend on s: fetch it behind them
#pragma unroll
            for (int m = 0; m < MAXR; ++m) nxt[m] = have(m) ? __ldcs(dn + at(m)) : -INFINITY;
        }
        // ... and pull the rows after it from HBM into L2 ahead of time (a step is shorter than an HBM round trip)
        if (t >= 2 + LK_TB_PF && lane * 32 < K) asm volatile("prefetch.global.L2 [%0];" ::"l"(dn - (size_t)LK_TB_PF * K + lane * 32));
        // the lane's first arg-max over its MAXR candidates as a tree over the slice number (ties keep the lower index; a NaN
        // candidate never wins, as in a sequential `c > v` scan; absent states are -inf through `cur`)
        float cc[MAXR];
        int cm[MAXR];
#pragma unroll
        for (int m = 0; m < MAXR; ++m) {
            const float c = __fadd_rn(cur[m], cv[m]);
            cc[m] = (c > -INFINITY) ? c : -INFINITY;
            cm[m] = m;
        }
#pragma unroll
        for (int w = 1; w < MAXR; w <<= 1)
#pragma unroll
            for (int m = 0; m + w < MAXR; m += 2 * w)
                if (cc[m + w] > cc[m]) { cc[m] = cc[m + w]; cm[m] = cm[m + w]; }
        float v = cc[0];
        int vi = (v > -INFINITY) ? lane + 32 * cm[0] : (lane < K ? lane : 0);
        lk_warp_argmax(v, vi);
        s = vi;
        if (lane == 0) st[t - 1] = s;
    }
}

// Full backpointer table on request (the traceback does not need it): psi_t[j] = first argmax_i(delta_{t-1}(i) + logP(i,j)),
// psi_0 = 0 (hmm.py:156, :167).  One warp per (sequence, frame, group of 32 target states): lane = target state j, the
// delta_{t-1} row is broadcast through shuffles, logPT rows are read coalesced.  uint8 for K <= 256, uint16 above.
template <typename PsiT>
__global__ void __launch_bounds__(256) lk_psi_kernel(const float *delta, const float *log_trans, int B, int T, int K, PsiT *psi) {
    const int64_t wid = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    const int jg = (K + 31) / 32;
    if (wid >= (int64_t)B * T * jg) return;
    const int g = (int)(wid % jg);
    const int64_t fr = wid / jg;                             // b * T + t
    const int t = (int)(fr % T);
    const int j = g * 32 + lane;
    int best_i = 0;
    if (t > 0) {
        const float *dprev = delta + (fr - 1) * K;
        float best = -INFINITY;
        for (int i0 = 0; i0 < K; i0 += 32) {
            const float dv = (i0 + lane < K) ? dprev[i0 + lane] : -INFINITY;
            const int n = min(32, K - i0);
            for (int ii = 0; ii < n; ++ii) {
                const float di = __shfl_sync(FULL_MASK, dv, ii);
                if (j < K) {
                    const float c = __fadd_rn(di, __ldg(log_trans + (size_t)(i0 + ii) * K + j));
                    if (c > best) { best = c; best_i = i0 + ii; }
                }
            }
        }
    }
    if (j < K) psi[fr * K + j] = (PsiT)best_i;
}

// ----------------------------------------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------------------------------------
static size_t lk_align256(size_t x) { return (x + 255) & ~(size_t)255; }

constexpr int XL_KMAX = 2048;                   // above LK_KMAX: one launch per time step (recursion_xlk.cu)
size_t xlk_extra_bytes(int B, int K);
int xlk_sweep(int mode3, const float *emis, int emis_mode, float eps, const float *M, const float *init, const float *rowmax,
              int B, int T, int K, float *ws_out, float *ws_l, void *extra, float *loglik, cudaStream_t s);

bool largek_shape_ok(int K) { return K > 32 && K <= XL_KMAX; }

// fb workspace: ws_a, ws_b [B,T,K]; la, lb, rowmax [B,T]; err flag
size_t largek_fb_workspace_bytes(int B, int T, int K) {
    const size_t n = (size_t)B * T;
    return 2 * lk_align256(n * K * sizeof(float)) + 3 * lk_align256(n * sizeof(float)) + 256 + (K > LK_KMAX ? xlk_extra_bytes(B, K) : 0);
}
// viterbi workspace: delta [B,T,K] (used when the caller does not want delta), logPT [K,K], rowmax [B,T], err flag
size_t largek_viterbi_workspace_bytes(int B, int T, int K) {
    const size_t n = (size_t)B * T;
    return lk_align256(n * K * sizeof(float)) + lk_align256((size_t)K * K * sizeof(float)) + lk_align256(n * sizeof(float)) + 256 +
           (K > LK_KMAX ? xlk_extra_bytes(B, K) : 0);
}

static int lk_cluster_size(int K) {
    const int need = (K + LK_NC - 1) / LK_NC;
    int cs = 1;
    while (cs < need) cs <<= 1;
    return cs;                                               // 1, 2, 4 or 8
}

// Clusters of `cs` CTAs that are co-resident on the device (every variant of the sweep kernels takes one SM per CTA).
static int lk_max_clusters(int cs) {
    static int cache[LK_CSMAX + 1] = {0};
    if (cache[cs] > 0) return cache[cs];
    auto kern = lk_sweep_kernel<LK_FWD, 4, 1>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LkSmem<4, 1>));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(cs * 32), 1, 1);
    cfg.blockDim = dim3(LK_THREADS, 1, 1);
    cfg.dynamicSmemBytes = sizeof(LkSmem<4, 1>);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n < 1) { cudaGetLastError(); n = 1; }
    return cache[cs] = n;
}

// An exchange wait that timed out (never in a correct run) leaves the sweep's outputs undefined: make that visible to a caller
// who does not look at the flag -- NaN log-likelihoods / scores and a -1 first state per sequence.
__global__ void lk_poison_kernel(const int *err, int B, int T, float *loglik, float *score, int64_t *states) {
    if (*err == 0) return;
    const float nan = __int_as_float(0x7fc00000);
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < B; b += gridDim.x * blockDim.x) {
        if (loglik) loglik[b] = nan;
        if (score) score[b] = nan;
        if (states) states[(size_t)b * T] = -1;
    }
}

// n_modes sweeps (1: mode0; 2: forward + backward; 3: + Viterbi) in one launch with NSQ x NP sequences per group
template <int NSQ, int NP>
static int lk_launch_variant(int mode0, const LkParams &pf, const LkParams &pb, const LkParams &pv, int n_modes, cudaStream_t s) {
    const size_t smem = sizeof(LkSmem<NSQ, NP>);
    const int ncl = (pf.B + NSQ * NP * pf.ngr - 1) / (NSQ * NP * pf.ngr);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(n_modes * ncl * pf.CS), 1, 1);
    cfg.blockDim = dim3(LK_THREADS, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)pf.CS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    auto launch = [&](auto kern, auto... args) -> int {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return set_error(HMMB200_ELAUNCH, "large-K smem opt-in: %s", cudaGetErrorString(e));
        e = cudaLaunchKernelEx(&cfg, kern, args...);
        if (e != cudaSuccess) return set_error(HMMB200_ELAUNCH, "large-K sweep launch: %s", cudaGetErrorString(e));
        return check_launch("lk_sweep_kernel");
    };
    if (n_modes > 1) return launch(lk_sweep_multi_kernel<NSQ, NP>, pf, pb, pv, ncl);
    if (mode0 == LK_FWD) return launch(lk_sweep_kernel<LK_FWD, NSQ, NP>, pf);
    if (mode0 == LK_BWD) return launch(lk_sweep_kernel<LK_BWD, NSQ, NP>, pf);
    return launch(lk_sweep_kernel<LK_VIT, NSQ, NP>, pf);
}

// How a sweep's batch is cut into clusters: `ngr` groups per cluster, NSQ sequences per register pass, NP passes per group.
// A step of a cluster (tools/lk_trace.py, profiles/r02_largek_exchange_experiments.txt; cycles) is the longest of
//   the products of all its sequences        ngr x NP x 370 NSQ                    (the compute warps do nothing else),
//   one group's dependency chain             NP x (370 NSQ + 600) + its exchange + 300   (product, finals, exchange),
//   the exchange of all its sequences        82 x CS per sequence                  (DSMEM bandwidth: what usually decides),
// and the launch takes ceil(clusters / co-resident clusters) waves of T such steps.  The cheapest estimate wins: as many clusters
// as are co-resident, and among equals the fewest groups (two-pass groups keep the finals and the barriers per sequence down).
struct LkShape { int ngr, nsq, np; };

static LkShape lk_choose_shape(int B, int cs, int n_modes) {
    const int maxc = lk_max_clusters(cs);
    LkShape best = {2, 4, 1};
    double best_cost = 1e300;
    for (int np = 1; np <= 2; ++np)
        for (int ngr = 2; ngr <= LK_NG / np; ++ngr)
            for (int nsq = 3; nsq <= 4; ++nsq) {
                const int per = nsq * np * ngr;
                const int ncl = (B + per - 1) / per;
                const int waves = (n_modes * ncl + maxc - 1) / maxc;
                const double prod = 370.0 * nsq, xchg = 82.0 * cs * nsq * np;
                double step = ngr * np * prod;
                const double chain = np * (prod + 600.0) + xchg + 300.0, queue = ngr * xchg + 150.0 * ngr;
                if (chain > step) step = chain;
                if (queue > step) step = queue;
                const double cost = waves * step;
                if (cost < best_cost) { best_cost = cost; best = {ngr, nsq, np}; }
            }
#ifdef HMMB200_DEBUG_HOOKS
    if (const char *f = getenv("HMMB200_LK_SHAPE")) {            // "ngr,nsq,np": debug builds can force a shape
        int a = 0, b = 0, c = 0;
        if (sscanf(f, "%d,%d,%d", &a, &b, &c) == 3 && a >= 1 && a * c <= LK_NG && (b == 3 || b == 4) && (c == 1 || c == 2)) best = {a, b, c};
    }
#endif
    return best;
}

static int lk_launch_shaped(int mode0, LkParams pf, LkParams pb, LkParams pv, int n_modes, cudaStream_t s) {
    const LkShape sh = lk_choose_shape(pf.B, pf.CS, n_modes);
    pf.ngr = pb.ngr = pv.ngr = sh.ngr;
    if (sh.np == 1) return (sh.nsq == 3) ? lk_launch_variant<3, 1>(mode0, pf, pb, pv, n_modes, s) : lk_launch_variant<4, 1>(mode0, pf, pb, pv, n_modes, s);
    return (sh.nsq == 3) ? lk_launch_variant<3, 2>(mode0, pf, pb, pv, n_modes, s) : lk_launch_variant<4, 2>(mode0, pf, pb, pv, n_modes, s);
}

// n_modes sweeps side by side in one launch
static int lk_launch_multi(const LkParams &pf, const LkParams &pb, const LkParams &pv, int n_modes, cudaStream_t s) {
#ifdef HMMB200_DEBUG_HOOKS
    if (getenv("HMMB200_LK_NO_MULTI")) return 1;
#endif
    return lk_launch_shaped(LK_FWD, pf, pb, pv, n_modes, s);
}

template <int MODE>
static int lk_launch(const LkParams &p, cudaStream_t s) { return lk_launch_shaped(MODE, p, p, p, 1, s); }

// The two halves of a large-K call, split so that the sweeps of a forward-backward pass and of a Viterbi pass can share one launch:
// *_prepare fills the sweep parameters and runs the small preparation kernels, *_finish turns the sweeps' raw outputs into results.
struct LkFbCall {
    LkParams p;            // forward sweep (the backward sweep is the same with loglik = null)
    float *rowmax;
    void *xl_extra;        // K > 512: scratch of the step-per-launch path
    int add_m;
    bool both;             // posteriors / backward values wanted: the backward sweep runs too
    float *gamma, *fwd_prob, *bwd_prob, *log_alpha, *log_beta, *loglik;
};

static int lk_fb_prepare(LkFbCall &c, const float *emis, int emis_mode, float floor_eps, int add_rowmax, const float *trans_prob,
                         const float *init_prob, int B, int T, int K, float *gamma, float *fwd_prob, float *bwd_prob,
                         float *log_alpha, float *log_beta, float *loglik, void *workspace, cudaStream_t s) {
    const size_t n = (size_t)B * T;
    uint8_t *w = (uint8_t *)workspace;
    LkParams &p = c.p;
    p = LkParams{};
    p.emis = emis; p.mode = emis_mode; p.eps = floor_eps; p.add_rowmax = add_rowmax;
    p.trans = trans_prob; p.init = init_prob; p.B = B; p.T = T; p.K = K; p.CS = lk_cluster_size(K); p.ngr = 2;
    p.ws_a = (float *)w;  w += lk_align256(n * K * sizeof(float));
    p.ws_b = (float *)w;  w += lk_align256(n * K * sizeof(float));
    p.ws_la = (float *)w; w += lk_align256(n * sizeof(float));
    p.ws_lb = (float *)w; w += lk_align256(n * sizeof(float));
    c.rowmax = (float *)w; w += lk_align256(n * sizeof(float));
    p.err = (int *)w;
    c.xl_extra = w + 256;
    cudaMemsetAsync(p.err, 0, sizeof(int), s);
#ifdef HMMB200_DEBUG_HOOKS
    p.trace = getenv("HMMB200_LK_TRACE") != nullptr;
#endif
    p.loglik = loglik;
    if (emis_mode == HMMB200_EMIS_LOG || emis_mode == HMMB200_EMIS_LOG_NORM_FLOOR) {
        lk_rowmax_kernel<<<(unsigned)((n + 7) / 8), 256, 0, s>>>(emis, (int64_t)n, K, c.rowmax);
        if (int rc = check_launch("lk_rowmax_kernel")) return rc;
        p.rowmax = c.rowmax;
    }
    c.add_m = (emis_mode == HMMB200_EMIS_LOG) || (emis_mode == HMMB200_EMIS_LOG_NORM_FLOOR && add_rowmax);
    c.both = gamma || fwd_prob || bwd_prob || log_alpha || log_beta;
    c.gamma = gamma; c.fwd_prob = fwd_prob; c.bwd_prob = bwd_prob; c.log_alpha = log_alpha; c.log_beta = log_beta; c.loglik = loglik;
    return HMMB200_OK;
}

static int lk_fb_finish(const LkFbCall &c, cudaStream_t s) {
    const LkParams &p = c.p;
    const size_t n = (size_t)p.B * p.T;
    lk_logscale_kernel<<<(unsigned)p.B, 32, 0, s>>>(p.ws_la, c.rowmax, p.T, 0, c.add_m, c.loglik);
    if (int rc = check_launch("lk_logscale_kernel")) return rc;
    if (c.both) {
        lk_logscale_kernel<<<(unsigned)p.B, 32, 0, s>>>(p.ws_lb, c.rowmax, p.T, 1, c.add_m, nullptr);
        if (int rc = check_launch("lk_logscale_kernel")) return rc;
        LkCombineParams cc;
        cc.ws_a = p.ws_a; cc.ws_b = p.ws_b; cc.ws_la = p.ws_la; cc.ws_lb = p.ws_lb; cc.n_frames = (int64_t)n; cc.K = p.K;
        cc.gamma = c.gamma; cc.fwd = c.fwd_prob; cc.bwd = c.bwd_prob; cc.log_alpha = c.log_alpha; cc.log_beta = c.log_beta;
        lk_combine_kernel<<<(unsigned)((n + 7) / 8), 256, 0, s>>>(cc);
        if (int rc = check_launch("lk_combine_kernel")) return rc;
    }
    if (c.loglik != nullptr) {
        lk_poison_kernel<<<1, 256, 0, s>>>(p.err, p.B, p.T, c.loglik, nullptr, nullptr);
        if (int rc = check_launch("lk_poison_kernel")) return rc;
    }
    return HMMB200_OK;
}

struct LkVitCall {
    LkParams p;
    float *logPT;
    void *xl_extra;
    const float *log_trans;
    void *psi;
    int64_t *states;
    float *score;
};

static int lk_vit_prepare(LkVitCall &c, const float *emis, int emis_mode, float floor_eps, const float *log_trans, const float *log_init,
                          int B, int T, int K, float *delta, void *psi, int64_t *states, float *score, void *workspace, cudaStream_t s) {
    const size_t n = (size_t)B * T;
    uint8_t *w = (uint8_t *)workspace;
    float *ws_delta = (float *)w; w += lk_align256(n * K * sizeof(float));
    c.logPT = (float *)w;         w += lk_align256((size_t)K * K * sizeof(float));
    float *rowmax = (float *)w;   w += lk_align256(n * sizeof(float));
    LkParams &p = c.p;
    p = LkParams{};
    p.emis = emis; p.mode = emis_mode; p.eps = floor_eps; p.trans = log_trans; p.init = log_init;
    p.B = B; p.T = T; p.K = K; p.CS = lk_cluster_size(K); p.ngr = 2;
    p.delta = delta ? delta : ws_delta;
    p.err = (int *)w;
    c.xl_extra = w + 256;
    cudaMemsetAsync(p.err, 0, sizeof(int), s);
#ifdef HMMB200_DEBUG_HOOKS
    p.trace = getenv("HMMB200_LK_TRACE") != nullptr;
#endif
    if (emis_mode == HMMB200_EMIS_LOG_NORM_FLOOR) {
        lk_rowmax_kernel<<<(unsigned)((n + 7) / 8), 256, 0, s>>>(emis, (int64_t)n, K, rowmax);
        if (int rc = check_launch("lk_rowmax_kernel")) return rc;
        p.rowmax = rowmax;
    }
    lk_transpose_kernel<<<(K * K + 255) / 256, 256, 0, s>>>(log_trans, K, c.logPT);
    if (int rc = check_launch("lk_transpose_kernel")) return rc;
    c.log_trans = log_trans; c.psi = psi; c.states = states; c.score = score;
    return HMMB200_OK;
}

static int lk_vit_finish(const LkVitCall &c, cudaStream_t s) {
    const LkParams &p = c.p;
    const size_t n = (size_t)p.B * p.T;
    const int B = p.B, T = p.T, K = p.K;
    const dim3 tb_grid((B + 3) / 4);
    if (K <= LK_KMAX && K % 32 == 0) lk_traceback_kernel<LK_KMAX / 32, true><<<tb_grid, 128, 0, s>>>(p.delta, c.logPT, B, T, K, c.states, c.score);
    else if (K <= LK_KMAX) lk_traceback_kernel<LK_KMAX / 32, false><<<tb_grid, 128, 0, s>>>(p.delta, c.logPT, B, T, K, c.states, c.score);
    else lk_traceback_kernel<XL_KMAX / 32, false><<<tb_grid, 128, 0, s>>>(p.delta, c.logPT, B, T, K, c.states, c.score);
    if (int rc = check_launch("lk_traceback_kernel")) return rc;
    if (c.psi != nullptr) {
        const int64_t n_warps = (int64_t)n * ((K + 31) / 32);
        const unsigned blocks = (unsigned)((n_warps + 7) / 8);
        if (K <= 256) lk_psi_kernel<uint8_t><<<blocks, 256, 0, s>>>(p.delta, c.log_trans, B, T, K, (uint8_t *)c.psi);
        else lk_psi_kernel<uint16_t><<<blocks, 256, 0, s>>>(p.delta, c.log_trans, B, T, K, (uint16_t *)c.psi);
        if (int rc = check_launch("lk_psi_kernel")) return rc;
    }
    lk_poison_kernel<<<1, 256, 0, s>>>(p.err, B, T, nullptr, c.score, c.states);
    return check_launch("lk_poison_kernel");
}

int largek_forward_backward(const float *emis, int emis_mode, float floor_eps, int add_rowmax, const float *trans_prob,
                            const float *init_prob, int B, int T, int K, float *gamma, float *fwd_prob, float *bwd_prob,
                            float *log_alpha, float *log_beta, float *loglik, void *workspace, cudaStream_t s) {
    LkFbCall c;
    if (int rc = lk_fb_prepare(c, emis, emis_mode, floor_eps, add_rowmax, trans_prob, init_prob, B, T, K, gamma, fwd_prob, bwd_prob,
                               log_alpha, log_beta, loglik, workspace, s)) return rc;
    if (K > LK_KMAX) {                                          // step-per-launch path (recursion_xlk.cu)
        const LkParams &p = c.p;
        if (int r = xlk_sweep(0, emis, emis_mode, floor_eps, trans_prob, init_prob, p.rowmax, B, T, K, p.ws_a, p.ws_la, c.xl_extra, loglik, s)) return r;
        if (c.both) {
            auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
            float *PT = (float *)((uint8_t *)c.xl_extra + 2 * al((size_t)B * K * sizeof(float)));
            lk_transpose_kernel<<<(K * K + 255) / 256, 256, 0, s>>>(trans_prob, K, PT);
            if (int r = check_launch("lk_transpose_kernel")) return r;
            if (int r = xlk_sweep(1, emis, emis_mode, floor_eps, PT, init_prob, p.rowmax, B, T, K, p.ws_b, p.ws_lb, c.xl_extra, nullptr, s)) return r;
        }
        return lk_fb_finish(c, s);
    }
    LkParams q = c.p;
    q.loglik = nullptr;
    int rc = c.both ? lk_launch_multi(c.p, q, q, 2, s) : 1;    // forward and backward side by side when their clusters fit
    if (rc < 0) return rc;
    if (rc == 1) {
        if (int r2 = lk_launch<LK_FWD>(c.p, s)) return r2;
        if (c.both) if (int r2 = lk_launch<LK_BWD>(q, s)) return r2;
    }
    return lk_fb_finish(c, s);
}

int largek_viterbi(const float *emis, int emis_mode, float floor_eps, const float *log_trans, const float *log_init,
                   int B, int T, int K, float *delta, void *psi, int64_t *states, float *score, void *workspace, cudaStream_t s) {
    LkVitCall c;
    if (int rc = lk_vit_prepare(c, emis, emis_mode, floor_eps, log_trans, log_init, B, T, K, delta, psi, states, score, workspace, s)) return rc;
    if (K > LK_KMAX) {
        if (int r = xlk_sweep(2, emis, emis_mode, floor_eps, log_trans, log_init, c.p.rowmax, B, T, K, c.p.delta, nullptr, c.xl_extra, nullptr, s)) return r;
        return lk_vit_finish(c, s);
    }
    if (int rc = lk_launch<LK_VIT>(c.p, s)) return rc;
    return lk_vit_finish(c, s);
}

// forward + backward + Viterbi of the same batch: the three sweeps in one launch when their clusters are co-resident
// (fb_workspace / vit_workspace: the two calls' own workspaces)
int largek_fb_viterbi(const float *emis, int fb_mode, int vit_mode, float floor_eps, int add_rowmax, const float *trans_prob,
                      const float *init_prob, const float *log_trans, const float *log_init, int B, int T, int K,
                      float *gamma, float *fwd_prob, float *bwd_prob, float *log_alpha, float *log_beta, float *loglik,
                      float *delta, void *psi, int64_t *states, float *score, void *fb_workspace, void *vit_workspace, cudaStream_t s) {
    if (K > LK_KMAX) {
        if (int rc = largek_forward_backward(emis, fb_mode, floor_eps, add_rowmax, trans_prob, init_prob, B, T, K, gamma, fwd_prob, bwd_prob,
                                             log_alpha, log_beta, loglik, fb_workspace, s)) return rc;
        return largek_viterbi(emis, vit_mode, floor_eps, log_trans, log_init, B, T, K, delta, psi, states, score, vit_workspace, s);
    }
    LkFbCall f;
    LkVitCall v;
    if (int rc = lk_fb_prepare(f, emis, fb_mode, floor_eps, add_rowmax, trans_prob, init_prob, B, T, K, gamma, fwd_prob, bwd_prob,
                               log_alpha, log_beta, loglik, fb_workspace, s)) return rc;
    if (int rc = lk_vit_prepare(v, emis, vit_mode, floor_eps, log_trans, log_init, B, T, K, delta, psi, states, score, vit_workspace, s)) return rc;
    LkParams q = f.p;
    q.loglik = nullptr;
    int rc = f.both ? lk_launch_multi(f.p, q, v.p, 3, s) : 1;
    if (rc < 0) return rc;
    if (rc == 1) {                                               // not co-resident: forward + backward together if possible, then Viterbi
        int r2 = f.both ? lk_launch_multi(f.p, q, q, 2, s) : 1;
        if (r2 < 0) return r2;
        if (r2 == 1) {
            if (int r3 = lk_launch<LK_FWD>(f.p, s)) return r3;
            if (f.both) if (int r3 = lk_launch<LK_BWD>(q, s)) return r3;
        }
        if (int r3 = lk_launch<LK_VIT>(v.p, s)) return r3;
    }
    if (int r4 = lk_fb_finish(f, s)) return r4;
    return lk_vit_finish(v, s);
}

}  // namespace hmmb200

#ifdef HMMB200_DEBUG_HOOKS
HMMB200_EXPORT int hmmb200_debug_lk_max_clusters(int cs) {
    if (cs < 1 || cs > hmmb200::LK_CSMAX) return -1;
    return hmmb200::lk_max_clusters(cs);
}

HMMB200_EXPORT int hmmb200_debug_lk_trace(long long *host64) {
    return (int)cudaMemcpyFromSymbol(host64, hmmb200::lk_trace_buf, sizeof(long long) * 64);
}
#endif
