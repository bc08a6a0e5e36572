// recursion_smallk.cuh -- device code of the small-K (K <= 32) recursions, shared by the stand-alone kernels
// (recursion_smallk.cu) and the fused forward + backward + Viterbi kernel (recursion_fused.cu).
#pragma once
// (design notes) recursion_smallk.cu -- small-K (K <= 32) HMM recursions for sm_100a, warp-specialised.
//
//   fb_sweep_kernel     forward and backward sweeps (blockIdx.y = direction), scaled-probability space.
//                       Replaces the per-time-step ATen launches of pytorch_hmm/hmm.py:95-117.
//   fb_combine_warp_kernel / fb_combine_kernel   posterior / exp(log alpha) / exp(log beta) from the two scaled sweeps
//                       (hmm.py:120-128): coalesced warp-per-32-frames form for K % 4 == 0, scalar form otherwise.
//   viterbi_kernel      max-plus recursion, packed uint8 backpointers in shared memory, chunk-parallel on-device
//                       traceback (hmm.py:159-178; mixture_gaussian.py:312-336).
//
// The time recursion is a latency problem (T dependent steps), so each CTA splits roles:
//   * ONE consumer warp carries the dependent chain and nothing else.  A warp holds NS = 32/G sequences; lane
//     (sub, j) owns state j of sequence `sub` and keeps its column (or row) of the transition matrix in registers.
//     Per step it reads the previous K-vector back from a shared-memory ring with K/4 broadcast LDS.128, does K
//     FMAs (or K add + max), one multiply/add with the emission term and one STS.  No global access, no
//     transcendental, no integer division on the chain.
//   * helper warps stream emissions from HBM a chunk of CH frames ahead, convert them to what the consumer needs
//     (floor, max-normalise, exp) into a shared-memory ring, and trail behind the consumer draining its results to
//     HBM with coalesced stores; for Viterbi they also recompute the backpointers from the stored delta vectors
//     (same fp32 adds -> bit-identical), which takes the arg-max off the chain entirely.
//   * hand-off is two named barriers per ring buffer (FULL: helpers -> consumer, DONE: consumer -> helpers).
// Scaling: the forward/backward vectors are renormalised every step by a power of two taken from the exponent of the
// previous step's largest entry (one REDUX; exact; integer exponent bookkeeping), so alpha = a * 2^ksum * exp(sum m).
// A lone warp issues roughly one instruction every ~4-5 cycles on a dependent chain, so the consumer's instruction
// count IS its latency: packed fp32x2 FMAs/adds, 3-input max, no per-step branches, 4x unrolled (fits the L0 I-cache).
#include "common.cuh"

#include <stdlib.h>
#include <type_traits>

namespace hmmb200 {

constexpr int CH = 64;          // frames per pipeline chunk
constexpr int NB = 2;           // ring buffers
constexpr int BT_PITCH = 33;    // floats per frame row of the emission ring (32 lanes + 1 pad)
constexpr int MAXNS = 8;        // sequences per warp at G = 4
constexpr int MR_BUFS = 2 * NB; // log-scale ring depth: loaders run up to NB chunks ahead of the drainer's read
// Named barriers of one pipeline (consumer + its loaders + its drainers): ids full0 + b / done0 + b, n participating threads.
// (0 is __syncthreads; a CTA hosting several pipelines gives each its own id range.)
struct PipeBars { int full0, done0, n; };
constexpr int BAR_FULL = 1;
constexpr int BAR_DONE = 1 + NB;
enum { SC_FULL = 0, SC_APPLY = 1, SC_ESTIMATE = 2 };   // fb consumer: what a step does about the power-of-two normaliser

__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }
// Programmatic dependent launch: a kernel launched with the stream-serialisation attribute may start while its predecessor is
// still running; everything before this call (shared-memory carve-up, parameters into registers) overlaps the predecessor's
// tail, everything after it sees the predecessor's memory.  A no-op when the kernel was launched the ordinary way.
__device__ __forceinline__ void grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---- raw emission stage: chunks of the [B,T,K] emission tensor land in shared memory by bulk asynchronous copies ----------------
// A chunk of CH frames of one sequence is CH*K*4 contiguous bytes in HBM.  Reading it with per-lane loads (lane = frame) costs one
// L1 wavefront per lane and per value -- ~770 wavefronts per loader warp and chunk, and the SM's single load/store FIFO is what the
// latency-critical consumer warps' shared-memory round trips queue in.  One cp.async.bulk per (sequence, chunk) moves the same
// bytes without occupying that FIFO at all; its arrival is counted on an mbarrier (transaction bytes); NR chunks are in flight.
constexpr int NR = 2;
__device__ __forceinline__ uint32_t sk_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sk_mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sk_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void sk_mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(sk_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void sk_mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sk_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sk_mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "SKWAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"      // suspends: a waiting warp burns no issue slots
        "@p bra SKDONE_%=;\n\t"
        "bra SKWAIT_%=;\n\t"
        "SKDONE_%=:\n\t"
        "}" ::"r"(sk_smem_u32(bar)), "r"(parity), "r"(1000000u) : "memory");
}
__device__ __forceinline__ void sk_bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(sk_smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(sk_smem_u32(bar)) : "memory");
}
// floats of one (buffer, sequence) block: CH*K values + 16 floats of padding so that the blocks of the two sequences of a warp sit
// 16 banks apart; 16-byte aligned for any K
__host__ __device__ inline int raw_seq_floats(int K) { return CH * K + 16 + ((4 - (CH * K) % 4) % 4); }
__host__ __device__ inline size_t raw_stage_bytes(int K, int NS) { return 64 + (size_t)NR * NS * raw_seq_floats(K) * sizeof(float); }
struct RawStage {
    uint64_t *full, *empty;   // [NR] each: bytes of the chunk have landed / every reader warp has copied its frames out
    float *buf;               // [NR][NS][raw_seq_floats(K)]
    int seq_floats;
    __device__ RawStage() : full(nullptr), empty(nullptr), buf(nullptr), seq_floats(0) {}
    __device__ RawStage(uint8_t *smem, int K, int /*bulk*/) {
        full = reinterpret_cast<uint64_t *>(smem);
        empty = full + NR;
        buf = reinterpret_cast<float *>(smem + 64);
        seq_floats = raw_seq_floats(K);
    }
    __device__ void init(int reader_warps) const {               // ONE thread, before a CTA-wide barrier
        for (int i = 0; i < NR; ++i) { sk_mbar_init(full + i, 1); sk_mbar_init(empty + i, reader_warps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
};

// Blackwell packed fp32 pairs (one issue slot for two IEEE round-to-nearest operations) and 3-input max.
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    unsigned long long ra = *reinterpret_cast<unsigned long long *>(&a), rb = *reinterpret_cast<unsigned long long *>(&b);
    unsigned long long rc = *reinterpret_cast<unsigned long long *>(&c), rd;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2 *>(&rd);
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
    unsigned long long ra = *reinterpret_cast<unsigned long long *>(&a), rb = *reinterpret_cast<unsigned long long *>(&b), rd;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    return *reinterpret_cast<float2 *>(&rd);
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
// exact maximum of N values as a 3-ary tree (order-independent, so bit-identical to any other order)
template <int N>
__device__ __forceinline__ float max_tree(float (&x)[N]) {
    if constexpr (N == 1) {
        return x[0];
    } else if constexpr (N == 2) {
        return fmaxf(x[0], x[1]);
    } else {
        constexpr int M = (N + 2) / 3;
        float y[M];
#pragma unroll
        for (int i = 0; i < M; ++i) {
            if (3 * i + 2 < N) y[i] = fmax3(x[3 * i], x[3 * i + 1], x[3 * i + 2]);
            else if (3 * i + 1 < N) y[i] = fmaxf(x[3 * i], x[3 * i + 1]);
            else y[i] = x[3 * i];
        }
        return max_tree<M>(y);
    }
}

// ----------------------------------------------------------------------------------------------------------
// per-frame emission transforms (helper side; one lane holds the K values of one frame)
// ----------------------------------------------------------------------------------------------------------
// scaled-probability form for forward/backward: b~ and the log-scale m divided out of the frame
template <int KP>
__device__ __forceinline__ void row_to_scaled(int mode, float eps, int K, float (&e)[KP], float &m) {
    m = 0.f;
    if (mode == HMMB200_EMIS_PROB_FLOOR) {
#pragma unroll
        for (int k = 0; k < KP; ++k) e[k] = (k < K) ? e[k] + eps : 0.f;
    } else if (mode == HMMB200_EMIS_LOG_EXP_FLOOR) {
#pragma unroll
        for (int k = 0; k < KP; ++k) e[k] = (k < K) ? expf(e[k]) + eps : 0.f;
    } else {
        float mx = -INFINITY;
#pragma unroll
        for (int k = 0; k < KP; ++k) if (k < K) mx = fmaxf(mx, e[k]);
        if (!(mx > -INFINITY)) mx = 0.f;                     // all states impossible: keep the frame finite
        const float add = (mode == HMMB200_EMIS_LOG_NORM_FLOOR) ? eps : 0.f;
#pragma unroll
        for (int k = 0; k < KP; ++k) e[k] = (k < K) ? expf(e[k] - mx) + add : 0.f;
        m = mx;
    }
}

// log form for Viterbi (fp32, the reference's formula for each input kind)
template <int KP>
__device__ __forceinline__ void row_to_log(int mode, float eps, int K, float (&e)[KP]) {
    if (mode == HMMB200_EMIS_LOG) return;
    if (mode == HMMB200_EMIS_PROB_FLOOR) {
#pragma unroll
        for (int k = 0; k < KP; ++k) e[k] = (k < K) ? logf(e[k] + eps) : 0.f;
    } else if (mode == HMMB200_EMIS_LOG_EXP_FLOOR) {
#pragma unroll
        for (int k = 0; k < KP; ++k) e[k] = (k < K) ? logf(expf(e[k]) + eps) : 0.f;
    } else {
        float mx = -INFINITY;
#pragma unroll
        for (int k = 0; k < KP; ++k) if (k < K) mx = fmaxf(mx, e[k]);
#pragma unroll
        for (int k = 0; k < KP; ++k) e[k] = (k < K) ? logf(expf(e[k] - mx) + eps) : 0.f;
    }
}

// scaled-probability form with the hardware exponential (ex2.approx, relative error 2^-22: far inside the 1e-4 contract of the
// posteriors, and a fifth of the instructions of expf -- the helper warps' instruction footprint is what the consumers' instruction
// fetches compete with, see the bulk loader below)
__device__ __forceinline__ float fast_exp(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x * 1.4426950408889634f));
    return y;
}
template <int KP>
__device__ __forceinline__ void row_to_scaled_fast(int mode, float eps, int K, float (&e)[KP], float &m) {
    m = 0.f;
    if (mode == HMMB200_EMIS_PROB_FLOOR) {
#pragma unroll
        for (int k = 0; k < KP; ++k) e[k] = e[k] + eps;
    } else if (mode == HMMB200_EMIS_LOG_EXP_FLOOR) {
#pragma unroll
        for (int k = 0; k < KP; ++k) e[k] = fast_exp(e[k]) + eps;
    } else {
        float t[KP];
#pragma unroll
        for (int k = 0; k < KP; ++k) t[k] = (k < K) ? e[k] : -INFINITY;
        float mx = max_tree<KP>(t);
        if (!(mx > -INFINITY)) mx = 0.f;                     // all states impossible: keep the frame finite
        const float add = (mode == HMMB200_EMIS_LOG_NORM_FLOOR) ? eps : 0.f;
#pragma unroll
        for (int k = 0; k < KP; ++k) e[k] = fast_exp(e[k] - mx) + add;
        m = mx;
    }
}

// Loader-warp loop: stream this CTA's NS sequences from HBM one chunk (CH frames) ahead of the consumer, transform each
// frame (scaled-probability form for forward/backward, log form for Viterbi) and publish the per-lane values in bt[b]
// (and the per-frame log-scale in mraw[b]).  NL loader warps split the frames; lane (sub, q) of loader `lw` handles
// FL = CH/G/NL consecutive frames of sequence `sub`.  Raw values for chunk c+1 are already in flight (registers)
// while chunk c is converted, so no HBM latency is exposed.
template <int G, int KP, int DIR, bool SCALED, int NL>
__device__ __forceinline__ void loader_loop(const float *emis, int mode, float eps, int B, int T, int K, int lw,
                                            float *bt, float *mraw, PipeBars pb) {
    constexpr int NS = 32 / G, FPL = CH / G, FL = FPL / NL;
    static_assert(FPL % NL == 0 && FL >= 1, "loader split");
    const int lane = threadIdx.x & 31;
    const int sub = lane / G, q = lane % G;
    const int seq = blockIdx.x * NS + sub;
    const bool seq_ok = seq < B;
    const int nch = (T + CH - 1) / CH;
    const int u_base = q * FPL + lw * FL;

    float cur[FL][KP], nxt[FL][KP];
    auto fetch = [&](int c, float (&dst)[FL][KP]) {
#pragma unroll
        for (int i = 0; i < FL; ++i) {
            const int n = c * CH + u_base + i;
            const bool ok = seq_ok && n < T;
            const int f = (DIR == 0) ? n : T - 1 - n;
            const float *row = emis + ((size_t)(ok ? seq : 0) * T + (ok ? f : 0)) * K;
#pragma unroll
            // (ld.global.cg, not the read-only path: an invariant load could be hoisted above griddepcontrol.wait by the compiler)
            for (int k = 0; k < KP; ++k) dst[i][k] = (ok && k < K) ? __ldcg(row + k) : 0.f;
        }
    };
    fetch(0, cur);
    for (int c = 0; c < nch; ++c) {
        const int b = c % NB;
        if (c + 1 < nch) fetch(c + 1, nxt);
        if (c >= NB) bar_sync(pb.done0 + b, pb.n);           // consumer is done reading bt[b] (chunk c - NB)
        float *btb = bt + (size_t)b * CH * BT_PITCH;
#pragma unroll
        for (int i = 0; i < FL; ++i) {
            const int u = u_base + i;
            const bool ok = seq_ok && (c * CH + u) < T;
            float m = 0.f;
            if (SCALED) row_to_scaled_fast<KP>(mode, eps, K, cur[i], m);   // (the same arithmetic as the bulk-copy feed: results do not depend on the tensor's alignment)
            else row_to_log<KP>(mode, eps, K, cur[i]);
            float *dst = btb + u * BT_PITCH + sub * G;
#pragma unroll
            for (int k = 0; k < G; ++k) {
                float v = 0.f;
                if (k < KP) v = (ok && k < K) ? cur[i][k] : 0.f;
                dst[k] = v;
            }
            if (SCALED) mraw[((size_t)(c % MR_BUFS) * CH + u) * MAXNS + sub] = ok ? m : 0.f;
        }
        bar_arrive(pb.full0 + b, pb.n);
#pragma unroll
        for (int i = 0; i < FL; ++i)
#pragma unroll
            for (int k = 0; k < KP; ++k) cur[i][k] = nxt[i][k];
    }
    for (int c = max(0, nch - NB); c < nch; ++c) bar_sync(pb.done0 + (c % NB), pb.n);
}

// Same role with the chunks arriving in the raw stage by bulk copies (lane 0 of loader 0 issues them NR chunks ahead; needs every
// chunk start and size to be a multiple of 16 bytes: (T*K) % 4 == 0 and a 16-byte aligned tensor).  Lane (sub, q) of loader `lw`
// takes the frames u = q + G*(lw*FL + i) of the chunk: neighbouring lanes read neighbouring frame rows (conflict-free 16-byte
// shared-memory loads at K = 12) and write neighbouring rows of the emission ring.
// Written for a SMALL instruction footprint (one frame per trip of a rolled loop, hardware exponential, `dir` a run-time value so
// that the forward and backward pipelines of a fused CTA run the same code): the three consumer warps of a fused CTA lose a third
// of their issue slots to instruction-cache misses when the helper warps around them stream tens of KB of unrolled code
// (ncu: stall_no_inst on the consumers' loops).
template <int G, int KP, bool SCALED, int NL>
__device__ __noinline__ void loader_loop_bulk(const float *emis, int mode, float eps, int B, int T, int K, int lw, int dir,
                                              float *bt, float *mraw, PipeBars pb, RawStage rs) {
    constexpr int NS = 32 / G, FPL = CH / G, FL = FPL / NL;
    static_assert(FPL % NL == 0 && FL >= 1, "loader split");
    const int lane = threadIdx.x & 31;
    const int sub = lane / G, q = lane % G;
    const int seq0 = blockIdx.x * NS;
    const bool seq_ok = seq0 + sub < B;
    const int nch = (T + CH - 1) / CH;
    const bool issuer = (lw == 0 && lane == 0);
    const int n_seq = min(NS, B - seq0);
    auto issue = [&](int c) {                                        // chunk c (sweep order) -> raw buffer c % NR
        const int rb = c % NR;
        const int nf = min(CH, T - c * CH);
        const int f_lo = (dir == 0) ? c * CH : T - c * CH - nf;      // lowest frame of the chunk: the block is frames f_lo .. f_lo+nf-1
        const uint32_t bytes = (uint32_t)nf * K * sizeof(float);
        sk_mbar_expect_tx(rs.full + rb, bytes * n_seq);
        for (int s = 0; s < n_seq; ++s)
            sk_bulk_g2s(rs.buf + (size_t)(rb * NS + s) * rs.seq_floats, emis + ((size_t)(seq0 + s) * T + f_lo) * K, bytes, rs.full + rb);
    };
    if (issuer) for (int c = 0; c < min(NR, nch); ++c) issue(c);
    // the ring's columns K .. G-1 (lanes that own no state) are zero for the whole sweep: written once here, never again
    // (each loader clears the rows it owns -- u in [lw*CH/NL, (lw+1)*CH/NL) of every ring buffer -- so no hand-off between loaders)
    for (int b = 0; b < NB; ++b)
        for (int e = lane; e < (CH / NL) * BT_PITCH; e += 32) bt[((size_t)b * CH + lw * (CH / NL)) * BT_PITCH + e] = 0.f;
    __syncwarp();
#pragma unroll 1
    for (int c = 0; c < nch; ++c) {
        const int b = c % NB, rb = c % NR;
        const int nf = min(CH, T - c * CH);
        sk_mbar_wait(rs.full + rb, (c / NR) & 1);
        if (c >= NB) bar_sync(pb.done0 + b, pb.n);                   // consumer is done reading bt[b] (chunk c - NB)
        float *btb = bt + (size_t)b * CH * BT_PITCH;
#pragma unroll 1
        for (int i = 0; i < FL; ++i) {
            const int u = q + G * (lw * FL + i);
            const bool ok = seq_ok && u < nf;
            const int row = ok ? ((dir == 0) ? u : nf - 1 - u) : 0;
            const float *src = rs.buf + (size_t)(rb * NS + sub) * rs.seq_floats + (size_t)row * K;
            float cur[KP];
            if ((KP % 4) == 0 && (K % 4) == 0) {
#pragma unroll
                for (int k4 = 0; k4 < KP / 4; ++k4) {
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (4 * k4 < K) v = *reinterpret_cast<const float4 *>(src + 4 * k4);
                    cur[4 * k4] = v.x; cur[4 * k4 + 1] = v.y; cur[4 * k4 + 2] = v.z; cur[4 * k4 + 3] = v.w;
                }
            } else {
#pragma unroll
                for (int k = 0; k < KP; ++k) cur[k] = (k < K) ? src[k] : 0.f;
            }
            float m = 0.f;
            if (SCALED) row_to_scaled_fast<KP>(mode, eps, K, cur, m);
            else row_to_log<KP>(mode, eps, K, cur);
            float *dst = btb + u * BT_PITCH + sub * G;
#pragma unroll
            for (int k = 0; k < KP; ++k) dst[k] = (ok && k < K) ? cur[k] : 0.f;
            if (SCALED) mraw[((size_t)(c % MR_BUFS) * CH + u) * MAXNS + sub] = ok ? m : 0.f;
        }
        // Release raw[rb].  Every load of the buffer has RETURNED by now (the stores above consumed the values and a warp issues in
        // order); a reader that let the issuer refill the buffer while its loads were still queued read the next chunk's rows (seen
        // as wrong paths next to a co-resident kernel).  The proxy fence orders these generic-proxy reads before the bulk copy's
        // asynchronous-proxy writes.
        __syncwarp();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if (lane == 0) sk_mbar_arrive(rs.empty + rb);
        if (issuer && c + NR < nch) {
            sk_mbar_wait(rs.empty + rb, (c / NR) & 1);               // ... and so have the other loader warps: refill the buffer
            issue(c + NR);
        }
        bar_arrive(pb.full0 + b, pb.n);
    }
    for (int c = max(0, nch - NB); c < nch; ++c) bar_sync(pb.done0 + (c % NB), pb.n);
}

// ----------------------------------------------------------------------------------------------------------
// forward / backward sweeps
// ----------------------------------------------------------------------------------------------------------
struct FbParams {
    const float *emis;
    int mode;
    float eps;
    int add_rowmax;
    const float *trans;   // [K,K] effective probabilities
    const float *init;    // [K]
    int B, T, K;
    float *ws_a, *ws_b;   // [B,T,K] scaled alpha / beta
    float *ws_la, *ws_lb; // [B,T]   log scale: alpha = a * exp(la), beta = b * exp(lb)
    float *loglik;        // [B] or null
    int pdl;              // launched with programmatic stream serialisation: wait for the previous kernel before touching its data
    int bulk;             // emission chunks arrive by bulk asynchronous copies (16-byte aligned tensor, (T*K) % 4 == 0)
};

constexpr int FB_NL = 2;                                   // loader warps
constexpr int FB_ND = 2;                                   // drainer warps
constexpr int FB_THREADS = 32 * (1 + FB_NL + FB_ND);
constexpr size_t FB_SMEM_BT = (size_t)NB * CH * BT_PITCH * sizeof(float);
constexpr size_t FB_SMEM_WR = (size_t)NB * CH * 32 * sizeof(float);
constexpr size_t FB_SMEM_ER = (size_t)NB * CH * MAXNS * sizeof(int);
constexpr size_t FB_SMEM_MR = (size_t)MR_BUFS * CH * MAXNS * sizeof(float);
constexpr size_t FB_SMEM_BYTES = FB_SMEM_MR + FB_SMEM_BT + 2 * FB_SMEM_WR + FB_SMEM_ER;   // backward sweep (the larger of the two)

// DIR 0: alpha_t(j) = (sum_i alpha_{t-1}(i) P(i,j)) b_t(j)                         (hmm.py:98-101)
// DIR 1: beta_t(i)  = sum_j P(i,j) b_{t+1}(j) beta_{t+1}(j)                         (hmm.py:113-117)
// Both are  w <- (sum_i w_prev(i) * M[i]) * b~ * r  on w = alpha (DIR 0) or w = beta .* b~ (DIR 1), M = the lane's
// column (DIR 0) / row (DIR 1) of P, r = 2^-k from the previous step's sum.
template <int G, int KP, int DIR, bool PAD>
__device__ __forceinline__ void fb_consumer(const FbParams &p, const float *bt, float *wr, float *br, int *er, PipeBars pb) {
    constexpr int NS = 32 / G;
    const int lane = threadIdx.x & 31;
    const int sub = lane / G, j = lane % G;
    const int seq = blockIdx.x * NS + sub;
    const int K = p.K, T = p.T;
    const bool lane_ok = seq < p.B && j < K;
    // PAD (KP < G): the last lane of each group owns no state and its ring slot is never read back by the recursion;
    // it publishes the running exponent there, so the bookkeeping costs no extra store.  Otherwise lane 0 writes
    // it to the `er` ring.
    const bool pad_lane = PAD && (j == G - 1);

    float2 M2[KP / 2];
#pragma unroll
    for (int i = 0; i < KP / 2; ++i) {
        float v0 = 0.f, v1 = 0.f;
        if (lane_ok && 2 * i < K) v0 = (DIR == 0) ? __ldg(p.trans + (2 * i) * K + j) : __ldg(p.trans + j * K + 2 * i);
        if (lane_ok && 2 * i + 1 < K) v1 = (DIR == 0) ? __ldg(p.trans + (2 * i + 1) * K + j) : __ldg(p.trans + j * K + 2 * i + 1);
        M2[i] = make_float2(v0, v1);
    }
    const float pi = (DIR == 0 && lane_ok) ? __ldg(p.init + j) : 0.f;

    float r_cur = 1.f;
    int k_cur = 0, ksum = 0;
    const float4 *prev = reinterpret_cast<const float4 *>(wr + sub * G);
    float *wp = wr, *bp2 = br;
    int *ep = er;

    // Power-of-two normaliser r = 2^-k for the NEXT step, from the exponent of the largest entry (exact scaling,
    // integer bookkeeping).  One sequence per warp: a single REDUX over the new vector.  Several sequences per warp:
    // a REDUX with per-group masks takes a slow divergent path, so each lane instead reduces the previous vector it
    // has just read back (3-input max tree, no cross-lane traffic; one more step of lag, which is harmless).
    auto set_scale = [&](unsigned mx) {
        const unsigned eb = mx >> 23;
        k_cur = (int)eb - 127;
        r_cur = __uint_as_float((254u - eb) << 23);
    };
    // one time step: w <- (sum_i prev[i] * M[i]) * (b~ * r) ; the pad lane carries (float)ksum instead.
    // The normaliser costs a third of the step's instructions (max tree, exponent arithmetic), and the step's
    // instruction count is its latency, so inside the 4x unrolled loop only every other step derives one:
    //   SC_APPLY     applies the pending scale, derives none;
    //   SC_ESTIMATE  applies none (r = 1 folds away), derives the next one from the vector it has just read back;
    //   SC_FULL      both (loop remainders).
    // Powers of two are exact, so posteriors do not depend on where the scales fall; three steps of lag are far from
    // underflow even on floored frames (1e-8 per step).
    auto step = [&](auto sc_tag, int u, float bqv) {
        constexpr int SC = decltype(sc_tag)::value;
        constexpr bool APPLY = (SC != SC_ESTIMATE), EST = (SC != SC_APPLY);
        float2 acc_a = make_float2(0.f, 0.f), acc_b = make_float2(0.f, 0.f);
        float v[KP];
#pragma unroll
        for (int i4 = 0; i4 < KP / 4; ++i4) {
            const float4 t = prev[i4];
            acc_a = ffma2(make_float2(t.x, t.y), M2[2 * i4], acc_a);
            acc_b = ffma2(make_float2(t.z, t.w), M2[2 * i4 + 1], acc_b);
            v[4 * i4 + 0] = t.x; v[4 * i4 + 1] = t.y; v[4 * i4 + 2] = t.z; v[4 * i4 + 3] = t.w;
        }
        const float2 s2 = fadd2(acc_a, acc_b);
        const float acc = s2.x + s2.y;
        const float mb = APPLY ? bqv * r_cur : bqv;
        if (APPLY) ksum += k_cur;
        const float padf = pad_lane ? (float)ksum : 0.f;
        const float w = fmaf(acc, mb, padf);
        if (DIR == 1) bp2[u * 32] = APPLY ? acc * r_cur : acc;
        wp[u * 32] = w;
        if (!PAD && j == 0) ep[u * MAXNS] = ksum;
        if (EST) {
            if (NS == 1) {
                set_scale(__reduce_max_sync(FULL_MASK, lane_ok ? __float_as_uint(w) : 0u));
            } else {
                // scale the new vector by what the previous one needed, times the growth r_cur already applied to it
                const float vm = APPLY ? max_tree<KP>(v) * r_cur : max_tree<KP>(v);
                set_scale(__float_as_uint(vm));
            }
        }
        prev = reinterpret_cast<const float4 *>(wp + u * 32 - j);
        __syncwarp();
    };
    using ScFull = std::integral_constant<int, SC_FULL>;
    using ScApply = std::integral_constant<int, SC_APPLY>;
    using ScEstimate = std::integral_constant<int, SC_ESTIMATE>;

    const int nch = (T + CH - 1) / CH;
    for (int c = 0; c < nch; ++c) {
        const int b = c % NB;
        bar_sync(pb.full0 + b, pb.n);
        const int nf = min(CH, T - c * CH);
        const float *btb = bt + (size_t)b * CH * BT_PITCH + lane;
        wp = wr + (size_t)b * CH * 32 + lane;
        bp2 = br + (size_t)b * CH * 32 + lane;
        ep = er + (size_t)b * CH * MAXNS + sub;
        int u = 0;
        if (c == 0) {
            // step 0: alpha_0 = p0 .* b_0 (hmm.py:92)  /  beta_{T-1} = 1 (hmm.py:107)
            const float b0 = btb[0];
            const float w = (DIR == 0) ? pi * b0 : (lane_ok ? b0 : 0.f);
            if (DIR == 1) bp2[0] = lane_ok ? 1.f : 0.f;
            wp[0] = w;                                       // pad lane: ksum = 0
            if (!PAD && j == 0) ep[0] = 0;
            if (NS == 1) set_scale(__reduce_max_sync(FULL_MASK, lane_ok ? __float_as_uint(w) : 0u));
            prev = reinterpret_cast<const float4 *>(wp - j);
            __syncwarp();
            u = 1;
        }
        for (; u + 4 <= nf; u += 4) {
            float bq[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) bq[i] = btb[(u + i) * BT_PITCH];
            step(ScApply{}, u, bq[0]);
            step(ScEstimate{}, u + 1, bq[1]);
            step(ScApply{}, u + 2, bq[2]);
            step(ScEstimate{}, u + 3, bq[3]);
        }
        for (; u < nf; ++u) step(ScFull{}, u, btb[u * BT_PITCH]);
        bar_arrive(pb.done0 + b, pb.n);
    }
}

// Drainer warp: trails the consumer by NB chunks; copies the scaled vectors from the ring to HBM with coalesced
// stores and turns the exponent / log-scale bookkeeping into the per-frame log scale la (alpha = a * exp(la)).
template <int G, int KP, int DIR, bool PAD>
__device__ __forceinline__ void fb_drainer(const FbParams &p, const float *wr, const float *br, const int *er,
                                           const float *mraw, int dw, PipeBars pb) {
    constexpr int NS = 32 / G;
    static_assert(CH == 64, "the log-scale scan assumes two frames per lane");
    const int lane = threadIdx.x & 31;
    const int K = p.K, T = p.T, B = p.B;
    const bool add_m = (p.mode == HMMB200_EMIS_LOG) || (p.mode == HMMB200_EMIS_LOG_NORM_FLOOR && p.add_rowmax);
    const int nch = (T + CH - 1) / CH;
    float *ws = (DIR == 0) ? p.ws_a : p.ws_b;
    float *wsl = (DIR == 0) ? p.ws_la : p.ws_lb;
    const float *src_ring = (DIR == 0) ? wr : br;
    double carry[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) carry[s] = 0.0;
    // running exponent published by the consumer: pad slot of the w ring (PAD) or the er ring
    auto ksum_at = [&](int b, int u, int s) -> double {
        return PAD ? (double)wr[((size_t)b * CH + u) * 32 + s * G + (G - 1)] : (double)er[((size_t)b * CH + u) * MAXNS + s];
    };

    auto drain = [&](int c, int b) {
        const int nf = min(CH, T - c * CH);
        const int n_lo = c * CH;
        const int f_lo = (DIR == 0) ? n_lo : T - n_lo - nf;          // lowest frame index of the chunk
        constexpr int R = 32 / G;                                    // frame rows covered per warp iteration
        const int rr = lane / G, k = lane % G;
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            const int sq = blockIdx.x * NS + s;
            if (sq < B) {
                const float *src = src_ring + (size_t)b * CH * 32 + s * G;
                float *dst = ws + ((size_t)sq * T + f_lo) * K;
                // rows are interleaved over the FB_ND drainer warps
#pragma unroll 4
                for (int fl0 = dw * R; fl0 < nf; fl0 += R * FB_ND) {
                    const int fl = fl0 + rr;
                    if (fl < nf && k < K) {
                        const int u = (DIR == 0) ? fl : nf - 1 - fl;
                        dst[(size_t)fl * K + k] = src[u * 32 + k];
                    }
                }
                if ((s % FB_ND) != dw) continue;                     // the log-scale of sequence s belongs to one drainer
                // log scale: prefix of the per-frame m (inclusive for alpha, exclusive for beta) + ln2 * exponent
                const int u0 = 2 * lane, u1 = 2 * lane + 1;
                const double m0 = add_m ? (double)mraw[((size_t)(c % MR_BUFS) * CH + u0) * MAXNS + s] : 0.0;
                const double m1 = add_m ? (double)mraw[((size_t)(c % MR_BUFS) * CH + u1) * MAXNS + s] : 0.0;
                double x = m0 + m1;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const double y = __shfl_up_sync(FULL_MASK, x, o);
                    if (lane >= o) x += y;
                }
                const double pre0 = carry[s] + (x - m1);             // inclusive prefix at u0
                const double pre1 = carry[s] + x;                    // inclusive prefix at u1
                carry[s] += __shfl_sync(FULL_MASK, x, 31);
                const double l0 = ((DIR == 0) ? pre0 : pre0 - m0) + 0.69314718055994530942 * ksum_at(b, u0, s);
                const double l1 = ((DIR == 0) ? pre1 : pre1 - m1) + 0.69314718055994530942 * ksum_at(b, u1, s);
                float *dl = wsl + (size_t)sq * T;
                if (u0 < nf) dl[(DIR == 0) ? n_lo + u0 : T - 1 - (n_lo + u0)] = (float)l0;
                if (u1 < nf) dl[(DIR == 0) ? n_lo + u1 : T - 1 - (n_lo + u1)] = (float)l1;
                if (DIR == 0 && p.loglik != nullptr && c == nch - 1) {
                    const int ul = (T - 1) - n_lo;                   // last frame: loglik = la + log(sum_k a)
                    if (u0 == ul || u1 == ul) {
                        const float *v = wr + ((size_t)b * CH + ul) * 32 + s * G;
                        float tot = 0.f;
                        for (int kk = 0; kk < K; ++kk) tot += v[kk];
                        p.loglik[sq] = (float)(((u0 == ul) ? l0 : l1) + (double)logf(tot));
                    }
                }
            }
        }
    };

#pragma unroll 1
    for (int c = 0; c < nch + NB; ++c) {                             // (one call site of the drain body: instruction footprint)
        const int b = c % NB;
        if (c >= NB) {
            bar_sync(pb.done0 + b, pb.n);
            drain(c - NB, b);
        }
        if (c < nch) bar_arrive(pb.full0 + b, pb.n);                 // ring buffer b drained: consumer may overwrite it
    }
}

// One sweep's shared-memory carve-up (the forward sweep has no `br` ring: DIR 0 never touches it; with a pad lane -- K < group
// width -- the running exponent travels in the w ring and there is no `er` ring either).
template <int DIR, bool PAD = false>
struct FbSmem {
    float *mraw, *bt, *wr, *br;
    int *er;
    static constexpr size_t BYTES = FB_SMEM_MR + FB_SMEM_BT + FB_SMEM_WR + (DIR == 1 ? FB_SMEM_WR : 0) + (PAD ? 0 : FB_SMEM_ER);
    __device__ explicit FbSmem(uint8_t *smem) {
        mraw = reinterpret_cast<float *>(smem);
        bt = reinterpret_cast<float *>(smem + FB_SMEM_MR);
        wr = reinterpret_cast<float *>(smem + FB_SMEM_MR + FB_SMEM_BT);
        br = (DIR == 1) ? reinterpret_cast<float *>(smem + FB_SMEM_MR + FB_SMEM_BT + FB_SMEM_WR) : wr;
        er = reinterpret_cast<int *>(smem + FB_SMEM_MR + FB_SMEM_BT + FB_SMEM_WR + (DIR == 1 ? FB_SMEM_WR : 0));
    }
};

// role: 0 consumer, 1 .. FB_NL loaders, FB_NL + 1 .. FB_NL + FB_ND drainers
template <int G, int KP, int DIR, bool PAD>
__device__ __forceinline__ void fb_roles(const FbParams &p, uint8_t *smem, int role, PipeBars pb, RawStage rs) {
    FbSmem<DIR, PAD> s(smem);
    if (role == 0) {
        fb_consumer<G, KP, DIR, PAD>(p, s.bt, s.wr, s.br, s.er, pb);       // (touches no global memory a predecessor writes)
    } else if (role <= FB_NL) {
        if (p.pdl) grid_dependency_wait();
        if (p.bulk) loader_loop_bulk<G, KP, true, FB_NL>(p.emis, p.mode, p.eps, p.B, p.T, p.K, role - 1, DIR, s.bt, s.mraw, pb, rs);
        else loader_loop<G, KP, DIR, true, FB_NL>(p.emis, p.mode, p.eps, p.B, p.T, p.K, role - 1, s.bt, s.mraw, pb);
    } else {
        if (p.pdl) grid_dependency_wait();
        fb_drainer<G, KP, DIR, PAD>(p, s.wr, s.br, s.er, s.mraw, role - 1 - FB_NL, pb);
    }
}

template <int G, int KP>
__global__ void __launch_bounds__(FB_THREADS) fb_sweep_kernel(FbParams p) {
    extern __shared__ __align__(16) uint8_t smem[];
    constexpr bool PAD = KP < G;
    const PipeBars pb = {BAR_FULL, BAR_DONE, FB_THREADS};
    const RawStage rs(smem + FB_SMEM_BYTES, p.K, p.bulk);                  // (bulk-copy feed only; unused bytes otherwise)
    if (p.bulk) {
        if (threadIdx.x == 0) rs.init(FB_NL);
        __syncthreads();
    }
    if (blockIdx.y == 0) fb_roles<G, KP, 0, PAD>(p, smem, threadIdx.x >> 5, pb, rs);
    else fb_roles<G, KP, 1, PAD>(p, smem, threadIdx.x >> 5, pb, rs);
}

// ----------------------------------------------------------------------------------------------------------
// combine: gamma = a.*b / sum, fwd = a*exp(la), bwd = b*exp(lb)           (hmm.py:120-128)
// ----------------------------------------------------------------------------------------------------------
struct CombineParams {
    const float *ws_a, *ws_b, *ws_la, *ws_lb;
    int64_t n_frames;
    int K;
    float *gamma, *fwd, *bwd, *log_alpha, *log_beta;
    int bf16;             // gamma / fwd / bwd are bfloat16 arrays (round to nearest even); log_alpha / log_beta must then be null
};

// two floats -> packed bf16 pair (first argument in the low half)
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    uint32_t d;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}

// Scalar form (K not a multiple of 4, or unaligned outputs): one thread per frame, COMBINE_FPT frames per thread.
// Explicit fused / rounded operations in a fixed order: a frame's result must not depend on how the batch was sharded.
constexpr int COMBINE_FPT = 2;

static __global__ void __launch_bounds__(256) fb_combine_kernel(CombineParams p) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t idx0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int K = p.K;
    for (int f = 0; f < COMBINE_FPT; ++f) {
        const int64_t idx = idx0 + f * stride;
        if (idx >= p.n_frames) continue;
        const float *a = p.ws_a + idx * K, *b = p.ws_b + idx * K;
        const float la = p.ws_la[idx], lb = p.ws_lb[idx];
        const float ea = expf(la), eb = expf(lb);
        float Z = 0.f;
        for (int k = 0; k < K; ++k) Z = fmaf(a[k], b[k], Z);
        const float inv = __fdiv_rn(1.f, Z);
        for (int k = 0; k < K; ++k) {
            float x = a[k], y = b[k];
            if (p.bf16) {
                unsigned short *g16 = reinterpret_cast<unsigned short *>(p.gamma), *f16 = reinterpret_cast<unsigned short *>(p.fwd);
                unsigned short *b16 = reinterpret_cast<unsigned short *>(p.bwd);
                if (g16) g16[idx * K + k] = (unsigned short)(pack_bf16x2(__fmul_rn(__fmul_rn(x, y), inv), 0.f) & 0xffffu);
                if (f16) f16[idx * K + k] = (unsigned short)(pack_bf16x2(x * ea, 0.f) & 0xffffu);
                if (b16) b16[idx * K + k] = (unsigned short)(pack_bf16x2(y * eb, 0.f) & 0xffffu);
                continue;
            }
            if (p.gamma) p.gamma[idx * K + k] = __fmul_rn(__fmul_rn(x, y), inv);
            if (p.fwd) p.fwd[idx * K + k] = x * ea;
            if (p.bwd) p.bwd[idx * K + k] = y * eb;
            if (p.log_alpha) p.log_alpha[idx * K + k] = logf(x) + la;
            if (p.log_beta) p.log_beta[idx * K + k] = logf(y) + lb;
        }
    }
}

// Coalesced form for K % 4 == 0: a warp finishes 32 consecutive frames = 32 * KV float4 pieces, lane l taking pieces
// l, l + 32, ... -- every load and store instruction covers 512 contiguous bytes (with one thread per frame the 32 lanes
// of an instruction sit K * 4 bytes apart: three times the L1 wavefronts and partial-sector stores at K = 12).  The
// per-piece dot products meet in a warp-private shared-memory row; each frame's normaliser is their sum in piece order,
// so a frame's posterior does not depend on its position in the batch (sharding-independent).
template <int KV>
__global__ void __launch_bounds__(256) fb_combine_warp_kernel(CombineParams p) {
    __shared__ float zs[8][32 * KV];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t n_blocks = (p.n_frames + 31) / 32;
    float *z = zs[wib];
    for (int64_t blk = (int64_t)blockIdx.x * 8 + wib; blk < n_blocks; blk += (int64_t)gridDim.x * 8) {
        const int64_t f0 = blk * 32;
        const int nf = (int)((p.n_frames - f0 < 32) ? (p.n_frames - f0) : 32);
        const int npieces = nf * KV;
        const float4 *pa = reinterpret_cast<const float4 *>(p.ws_a) + f0 * KV;
        const float4 *pb = reinterpret_cast<const float4 *>(p.ws_b) + f0 * KV;
        float4 xa[KV], xb[KV];
#pragma unroll
        for (int j = 0; j < KV; ++j) {
            const int idx = lane + 32 * j;
            const bool ok = idx < npieces;
            xa[j] = ok ? __ldcs(pa + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
            xb[j] = ok ? __ldcs(pb + idx) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        const float la_l = (lane < nf) ? __ldcs(p.ws_la + f0 + lane) : 0.f;
        const float lb_l = (lane < nf) ? __ldcs(p.ws_lb + f0 + lane) : 0.f;
#pragma unroll
        for (int j = 0; j < KV; ++j) {
            float d = __fmul_rn(xa[j].x, xb[j].x);
            d = __fmaf_rn(xa[j].y, xb[j].y, d);
            d = __fmaf_rn(xa[j].z, xb[j].z, d);
            d = __fmaf_rn(xa[j].w, xb[j].w, d);
            z[lane + 32 * j] = d;
        }
        __syncwarp();
        const float ea_l = expf(la_l), eb_l = expf(lb_l);
#pragma unroll
        for (int j = 0; j < KV; ++j) {
            const int idx = lane + 32 * j;
            const bool ok = idx < npieces;
            const int fr = ok ? idx / KV : 0;
            float Z = z[fr * KV];
#pragma unroll
            for (int q = 1; q < KV; ++q) Z = __fadd_rn(Z, z[fr * KV + q]);
            const float inv = 1.f / Z;
            const float ea = __shfl_sync(FULL_MASK, ea_l, fr), eb = __shfl_sync(FULL_MASK, eb_l, fr);
            const float la = __shfl_sync(FULL_MASK, la_l, fr), lb = __shfl_sync(FULL_MASK, lb_l, fr);
            if (!ok) continue;
            const float4 x = xa[j], y = xb[j];
            const int64_t o = f0 * KV + idx;
            auto g = [&](float u, float v) { return __fmul_rn(__fmul_rn(u, v), inv); };
            if (p.bf16) {                                    // 8-byte pieces: half the bytes of the fp32 outputs
                if (p.gamma) __stcs(reinterpret_cast<uint2 *>(p.gamma) + o, make_uint2(pack_bf16x2(g(x.x, y.x), g(x.y, y.y)), pack_bf16x2(g(x.z, y.z), g(x.w, y.w))));
                if (p.fwd) __stcs(reinterpret_cast<uint2 *>(p.fwd) + o, make_uint2(pack_bf16x2(x.x * ea, x.y * ea), pack_bf16x2(x.z * ea, x.w * ea)));
                if (p.bwd) __stcs(reinterpret_cast<uint2 *>(p.bwd) + o, make_uint2(pack_bf16x2(y.x * eb, y.y * eb), pack_bf16x2(y.z * eb, y.w * eb)));
                continue;
            }
            if (p.gamma) __stcs(reinterpret_cast<float4 *>(p.gamma) + o, make_float4(g(x.x, y.x), g(x.y, y.y), g(x.z, y.z), g(x.w, y.w)));
            if (p.fwd) __stcs(reinterpret_cast<float4 *>(p.fwd) + o, make_float4(x.x * ea, x.y * ea, x.z * ea, x.w * ea));
            if (p.bwd) __stcs(reinterpret_cast<float4 *>(p.bwd) + o, make_float4(y.x * eb, y.y * eb, y.z * eb, y.w * eb));
            if (p.log_alpha) __stcs(reinterpret_cast<float4 *>(p.log_alpha) + o,
                                    make_float4(logf(x.x) + la, logf(x.y) + la, logf(x.z) + la, logf(x.w) + la));
            if (p.log_beta) __stcs(reinterpret_cast<float4 *>(p.log_beta) + o,
                                   make_float4(logf(y.x) + lb, logf(y.y) + lb, logf(y.z) + lb, logf(y.w) + lb));
        }
        __syncwarp();                                        // the row of partial dots is reused by the next block
    }
}

// ----------------------------------------------------------------------------------------------------------
// Viterbi
// ----------------------------------------------------------------------------------------------------------
struct VitParams {
    const float *emis;
    int mode;
    float eps;
    const float *log_trans, *log_init;
    int B, T, K;
    float *delta;        // [B,T,K] or null
    uint8_t *psi_out;    // [B,T,K] or null
    int64_t *states;     // [B,T]
    float *score;        // [B] or null
    uint8_t *psi_ws;     // [B,T,G] global fallback when the backpointers do not fit in shared memory
    int psi_in_smem;
    int chunk;           // traceback chunk length L
    int n_chunks;        // ceil((T-1)/L)
    int pdl;             // launched with programmatic stream serialisation: wait for the previous kernel before touching its data
    int bulk;            // emission chunks arrive by bulk asynchronous copies (16-byte aligned tensor, (T*K) % 4 == 0)
};

constexpr int VIT_NL = 2;                                  // loader warps
constexpr int VIT_ND = 5;                                  // drainer warps (delta store + backpointers)
constexpr int VIT_THREADS = 32 * (1 + VIT_NL + VIT_ND);
constexpr size_t VIT_SMEM_BT = (size_t)NB * CH * BT_PITCH * sizeof(float);
constexpr int DR_PITCH = 36;                               // floats per frame row of the delta ring: 32 lanes + 4, so that the drainers'
                                                           // 16-byte reads of DIFFERENT rows (lane = frame) fall on different banks
constexpr size_t VIT_SMEM_DR = (size_t)NB * CH * DR_PITCH * sizeof(float);
constexpr size_t VIT_SMEM_CARRY = 2 * 32 * sizeof(float);
constexpr size_t VIT_SMEM_MT = 32 * 32 * sizeof(float);    // transposed log-transition table of the drainers (KP x KP used)
constexpr size_t VIT_SMEM_PIPE = VIT_SMEM_BT + VIT_SMEM_DR + VIT_SMEM_CARRY + VIT_SMEM_MT;

// Shared-memory layout (bytes), NS sequences per CTA:
//   [bt ring][delta ring][carry][logP^T][psi: NS*T*G if psi_in_smem][st: NS*T][exit: NS*n_chunks*G][entry: NS*n_chunks][final]
// vit_roles is the whole Viterbi pipeline of one group of NS sequences as seen by ONE warp: `role` 0 is the consumer, 1 .. VIT_NL
// the loaders, above that the drainers; `vtid` in [0, VIT_THREADS) numbers the pipeline's threads for the traceback, and
// `sync_all` synchronises exactly these VIT_THREADS threads (__syncthreads in the stand-alone kernel, a named barrier when the
// pipeline shares its CTA with the forward / backward sweeps).
template <int G, int KP, typename SyncAll>
__device__ __forceinline__ void vit_roles(const VitParams &p, uint8_t *smem, int role, int vtid, PipeBars pb, RawStage rs, SyncAll sync_all) {
    constexpr int NS = 32 / G;
    constexpr bool NIB = G <= 16;                         // backpointers < 16: two per byte in shared memory
    constexpr int PSI_ROW = NIB ? G / 2 : G;              // bytes per (sequence, frame) of the shared-memory table
    const int K = p.K, T = p.T, B = p.B;
    const int lane = threadIdx.x & 31;
    const int seq_base = blockIdx.x * NS;

    float *bt = reinterpret_cast<float *>(smem);
    float *dr = reinterpret_cast<float *>(smem + VIT_SMEM_BT);
    float *carry = reinterpret_cast<float *>(smem + VIT_SMEM_BT + VIT_SMEM_DR);
    float *mt = reinterpret_cast<float *>(smem + VIT_SMEM_BT + VIT_SMEM_DR + VIT_SMEM_CARRY);
    uint8_t *psi_s = smem + VIT_SMEM_PIPE;
    size_t off = VIT_SMEM_PIPE + (p.psi_in_smem ? (size_t)NS * T * PSI_ROW : 0);
    uint8_t *st_s = smem + off;            off += (size_t)NS * T;
    uint8_t *exit_s = smem + off;          off += (size_t)NS * p.n_chunks * G;
    uint8_t *entry_s = smem + off;         off += (size_t)NS * p.n_chunks;
    off = (off + 15) & ~(size_t)15;
    int *final_s = reinterpret_cast<int *>(smem + off);

    const int sub = lane / G, j = lane % G;
    const int seq = seq_base + sub;
    const bool seq_ok = seq < B;
    const bool lane_ok = seq_ok && j < K;
    const int nch = (T + CH - 1) / CH;

    // every warp keeps the lane's transition column: the consumer for the recursion, the helpers for psi.
    // Lanes that own no state hold -inf, so their delta stays -inf without any select on the chain.
    float2 M2[KP / 2];
#pragma unroll
    for (int i = 0; i < KP / 2; ++i) {
        const float v0 = (lane_ok && 2 * i < K) ? __ldg(p.log_trans + (2 * i) * K + j) : -INFINITY;
        const float v1 = (lane_ok && 2 * i + 1 < K) ? __ldg(p.log_trans + (2 * i + 1) * K + j) : -INFINITY;
        M2[i] = make_float2(v0, v1);
    }
    // c[i] = delta_prev[i] + logP[i][j] as packed IEEE adds (bit-identical to scalar adds)
    auto candidates = [&](const float4 *pv, float (&cv)[KP]) {
#pragma unroll
        for (int i4 = 0; i4 < KP / 4; ++i4) {
            const float4 t = pv[i4];
            const float2 lo = fadd2(make_float2(t.x, t.y), M2[2 * i4]);
            const float2 hi = fadd2(make_float2(t.z, t.w), M2[2 * i4 + 1]);
            cv[4 * i4 + 0] = lo.x; cv[4 * i4 + 1] = lo.y; cv[4 * i4 + 2] = hi.x; cv[4 * i4 + 3] = hi.y;
        }
    };

    if (role == 0) {
        // ---------------- consumer: delta_t(j) = max_i(delta_{t-1}(i) + logP(i,j)) + log b_t(j) --------------
        const float li = lane_ok ? __ldg(p.log_init + j) : -INFINITY;
        float d = -INFINITY;
        const float4 *prev = reinterpret_cast<const float4 *>(dr + sub * G);
        float *dp = dr;
        auto step = [&](int u, float eqv) {
            float cv[KP];
            candidates(prev, cv);
            d = __fadd_rn(max_tree<KP>(cv), eqv);
            dp[u * DR_PITCH] = d;
            prev = reinterpret_cast<const float4 *>(dp + u * DR_PITCH - j);
            __syncwarp();
        };
        for (int c = 0; c < nch; ++c) {
            const int b = c % NB;
            bar_sync(pb.full0 + b, pb.n);
            const int nf = min(CH, T - c * CH);
            const float *btb = bt + (size_t)b * CH * BT_PITCH + lane;
            dp = dr + (size_t)b * CH * DR_PITCH + lane;
            int u = 0;
            if (c == 0) {
                d = __fadd_rn(li, btb[0]);                  // delta_0 = log_p0 + log b_0 (hmm.py:159)
                dp[0] = d;
                prev = reinterpret_cast<const float4 *>(dp - j);
                __syncwarp();
                u = 1;
            }
            for (; u + 4 <= nf; u += 4) {
                float eq[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) eq[i] = btb[(u + i) * BT_PITCH];
#pragma unroll
                for (int i = 0; i < 4; ++i) step(u + i, eq[i]);
            }
            for (; u < nf; ++u) step(u, btb[u * BT_PITCH]);
            bar_arrive(pb.done0 + b, pb.n);
        }
        // final state: first index of the maximum (hmm.py:174)
        float bv = lane_ok ? d : -INFINITY;
        int bi = j;
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(FULL_MASK, bv, o, G);
            const int oi = __shfl_xor_sync(FULL_MASK, bi, o, G);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (j == 0) {
            final_s[sub] = (bi < K) ? bi : 0;
            if (seq_ok && p.score) p.score[seq] = bv;
        }
    } else if (role <= VIT_NL) {
        // ---------------- loaders: emission feed -----------------------------------------------------------
        if (p.pdl) grid_dependency_wait();
        if (p.bulk) loader_loop_bulk<G, KP, false, VIT_NL>(p.emis, p.mode, p.eps, B, T, K, role - 1, 0, bt, nullptr, pb, rs);
        else loader_loop<G, KP, 0, false, VIT_NL>(p.emis, p.mode, p.eps, B, T, K, role - 1, bt, nullptr, pb);
    } else {
        // ---------------- drainers: delta store and backpointers, NB chunks behind the consumer ------------
        // One LANE per (frame, sequence): the lane reads the frame's and the previous frame's delta vectors (conflict-free 16-byte
        // loads thanks to DR_PITCH), recomputes all K arg-maxima against the transposed transition table in shared memory
        // (uniform-address loads) with the consumer's own fp32 adds, and leaves with whole rows: K floats of delta (coalesced
        // 16-byte stores), one packed backpointer row.  ~30 warp instructions per frame instead of ~90 with a lane per state.
        const int hw = role - 1 - VIT_NL;
        if (p.pdl) grid_dependency_wait();                    // delta / psi outputs may still be read by the previous kernel's consumers
        // every drainer warp fills the whole table with the same values (no cross-warp hand-off needed): mt[j][i] = logP[i][j]
        for (int e = lane; e < KP * KP; e += 32) {
            const int jj = e / KP, ii = e % KP;
            mt[e] = (ii < K && jj < K) ? __ldg(p.log_trans + ii * K + jj) : -INFINITY;
        }
        __syncwarp();
        const bool delta_vec = (K % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.delta) & 15) == 0);
        auto drain = [&](int c, int b) {
            const int nf = min(CH, T - c * CH);
            const float *drb = dr + (size_t)b * CH * DR_PITCH;
            for (int task = hw * 32 + lane; task < nf * NS; task += VIT_ND * 32) {
                const int u = task / NS, s = task % NS;
                const int n = c * CH + u;
                const int sq = seq_base + s;
                const bool ok = sq < B;
                const float4 *pv = reinterpret_cast<const float4 *>(
                    (u > 0) ? (drb + (u - 1) * DR_PITCH + s * G) : (carry + ((c - 1) & 1) * 32 + s * G));
                float prev[KP];
#pragma unroll
                for (int i4 = 0; i4 < KP / 4; ++i4) {
                    const float4 t = pv[i4];
                    prev[4 * i4] = t.x; prev[4 * i4 + 1] = t.y; prev[4 * i4 + 2] = t.z; prev[4 * i4 + 3] = t.w;
                }
                // backpointer of state j: lowest index attaining max_i(delta_{n-1}(i) + logP(i,j))  (torch.max tie rule, hmm.py:167),
                // from the stored delta vector with the same fp32 adds as the consumer; psi_0 = 0.
                // four states per trip of a rolled loop (small instruction footprint, see loader_loop_bulk): their backpointers leave
                // as one 16-bit (nibbles) or 32-bit (bytes) piece of the row
                uint8_t *row_s = psi_s + ((size_t)s * T + n) * PSI_ROW;
                uint8_t *row_g = p.psi_in_smem ? nullptr : p.psi_ws + ((size_t)(ok ? sq : 0) * T + n) * G;
                uint8_t *row_o = (ok && p.psi_out) ? p.psi_out + ((size_t)sq * T + n) * K : nullptr;
#pragma unroll 1
                for (int j4 = 0; j4 < KP / 4; ++j4) {
                    uint32_t arg4[4];
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const float4 *col = reinterpret_cast<const float4 *>(mt + (4 * j4 + r) * KP);
                        float cv[KP];
#pragma unroll
                        for (int i4 = 0; i4 < KP / 4; ++i4) {
                            const float4 m = col[i4];
                            const float2 lo = fadd2(make_float2(prev[4 * i4], prev[4 * i4 + 1]), make_float2(m.x, m.y));
                            const float2 hi = fadd2(make_float2(prev[4 * i4 + 2], prev[4 * i4 + 3]), make_float2(m.z, m.w));
                            cv[4 * i4 + 0] = lo.x; cv[4 * i4 + 1] = lo.y; cv[4 * i4 + 2] = hi.x; cv[4 * i4 + 3] = hi.y;
                        }
                        const float best = max_tree<KP>(cv);
                        uint32_t arg = KP - 1;                           // (stays in range if everything is NaN)
#pragma unroll
                        for (int i = KP - 2; i >= 0; --i) arg = (cv[i] == best) ? (uint32_t)i : arg;
                        arg4[r] = (n > 0) ? arg : 0u;                    // psi_0 = 0
                    }
                    if (p.psi_in_smem) {
                        if (NIB) *reinterpret_cast<uint16_t *>(row_s + 2 * j4) = (uint16_t)(arg4[0] | (arg4[1] << 4) | (arg4[2] << 8) | (arg4[3] << 12));
                        else *reinterpret_cast<uint32_t *>(row_s + 4 * j4) = arg4[0] | (arg4[1] << 8) | (arg4[2] << 16) | (arg4[3] << 24);
                    } else if (ok) {
                        *reinterpret_cast<uint32_t *>(row_g + 4 * j4) = arg4[0] | (arg4[1] << 8) | (arg4[2] << 16) | (arg4[3] << 24);
                    }
                    if (row_o) {
#pragma unroll
                        for (int r = 0; r < 4; ++r) if (4 * j4 + r < K) row_o[4 * j4 + r] = (uint8_t)arg4[r];
                    }
                }
                // the frame's own delta vector: output row, and the carry for the first frame of the next chunk
                const float4 *cvp = reinterpret_cast<const float4 *>(drb + u * DR_PITCH + s * G);
                float4 cur[KP / 4];
#pragma unroll
                for (int i4 = 0; i4 < KP / 4; ++i4) cur[i4] = cvp[i4];
                if (ok && p.delta) {
                    float *dst = p.delta + ((size_t)sq * T + n) * K;
                    if (delta_vec) {
#pragma unroll
                        for (int i4 = 0; i4 < KP / 4; ++i4) if (4 * i4 < K) reinterpret_cast<float4 *>(dst)[i4] = cur[i4];
                    } else {
#pragma unroll
                        for (int i4 = 0; i4 < KP / 4; ++i4) {
                            const float v[4] = {cur[i4].x, cur[i4].y, cur[i4].z, cur[i4].w};
#pragma unroll
                            for (int q = 0; q < 4; ++q) if (4 * i4 + q < K) dst[4 * i4 + q] = v[q];
                        }
                    }
                }
                if (u == nf - 1) {
                    float4 *cd = reinterpret_cast<float4 *>(carry + (c & 1) * 32 + s * G);
#pragma unroll
                    for (int i4 = 0; i4 < KP / 4; ++i4) cd[i4] = cur[i4];
                }
            }
        };
        for (int c = 0; c < nch + NB; ++c) {                    // (one call site: the drain body is large)
            const int b = c % NB;
            if (c >= NB) {
                bar_sync(pb.done0 + b, pb.n);
                drain(c - NB, b);
            }
            if (c < nch) bar_arrive(pb.full0 + b, pb.n);        // ring buffer b drained: consumer may overwrite it
        }
        if (!p.psi_in_smem) __threadfence_block();
    }
    sync_all();

    // ---------------- chunk-parallel traceback: all threads of the pipeline ----------------------------
    // chunk c covers t in [1 + c*L, min(T-1, (c+1)*L)]; following psi from its top frame to its bottom frame
    // maps the state at t_hi to the state at t_lo - 1.
    const int L = p.chunk, nC = p.n_chunks;
    auto psi_at = [&](int s_sub, int s_seq, int t, int s) -> int {
        if (!p.psi_in_smem) return p.psi_ws[((size_t)s_seq * T + t) * G + s];
        if (NIB) return (psi_s[((size_t)s_sub * T + t) * PSI_ROW + (s >> 1)] >> ((s & 1) * 4)) & 15;
        return psi_s[((size_t)s_sub * T + t) * PSI_ROW + s];
    };
    // phase A: exit state for every (sequence, chunk, entry state)
    for (int task = vtid; task < NS * nC * K; task += VIT_THREADS) {
        const int e = task % K, c = (task / K) % nC, s_sub = task / (K * nC);
        const int s_seq = min(seq_base + s_sub, B - 1);
        const int t_lo = 1 + c * L, t_hi = min(T - 1, t_lo + L - 1);
        int s = e;
        for (int t = t_hi; t >= t_lo; --t) s = psi_at(s_sub, s_seq, t, s);
        exit_s[((size_t)s_sub * nC + c) * G + e] = (uint8_t)s;
    }
    sync_all();
    // phase B: the true entry state of every chunk (serial over chunks, one thread per sequence)
    if (vtid < NS) {
        int s = final_s[vtid];
        st_s[(size_t)vtid * T + (T - 1)] = (uint8_t)s;
        for (int c = nC - 1; c >= 0; --c) {
            entry_s[(size_t)vtid * nC + c] = (uint8_t)s;
            s = exit_s[((size_t)vtid * nC + c) * G + s];
        }
    }
    sync_all();
    // phase C: re-walk every chunk from its true entry state, recording the path
    for (int task = vtid; task < NS * nC; task += VIT_THREADS) {
        const int c = task % nC, s_sub = task / nC;
        const int s_seq = min(seq_base + s_sub, B - 1);
        const int t_lo = 1 + c * L, t_hi = min(T - 1, t_lo + L - 1);
        int s = entry_s[(size_t)s_sub * nC + c];
        for (int t = t_hi; t >= t_lo; --t) {
            s = psi_at(s_sub, s_seq, t, s);
            st_s[(size_t)s_sub * T + (t - 1)] = (uint8_t)s;
        }
    }
    sync_all();
    // phase D: coalesced int64 store
    for (int i = vtid; i < NS * T; i += VIT_THREADS) {
        const int s_sub = i / T, t = i % T;
        const int s_seq = seq_base + s_sub;
        if (s_seq < B) p.states[(size_t)s_seq * T + t] = (int64_t)st_s[i];
    }
}

template <int G, int KP>
__global__ void __launch_bounds__(VIT_THREADS) viterbi_kernel(VitParams p) {
    extern __shared__ __align__(16) uint8_t smem[];
    const PipeBars pb = {BAR_FULL, BAR_DONE, VIT_THREADS};
    const RawStage rs(smem, p.K, p.bulk);                                // the raw stage leads the layout (bulk-copy feed only)
    const size_t raw = p.bulk ? raw_stage_bytes(p.K, 32 / G) : 0;
    if (p.bulk) {
        if (threadIdx.x == 0) rs.init(VIT_NL);
        __syncthreads();
    }
    vit_roles<G, KP>(p, smem + raw, threadIdx.x >> 5, threadIdx.x, pb, rs, [] { __syncthreads(); });
}

// shared-memory bytes of one Viterbi pipeline and its traceback plan (chunk length L, number of chunks, where psi lives)
inline size_t vit_smem_bytes(int NS, int T, int G, int nC, bool psi_in_smem) {
    size_t off = VIT_SMEM_PIPE + (psi_in_smem ? (size_t)NS * T * (G <= 16 ? G / 2 : G) : 0);
    off += (size_t)NS * T + (size_t)NS * nC * G + (size_t)NS * nC;
    off = (off + 15) & ~(size_t)15;
    return off + NS * sizeof(int) * 2;
}

inline void vit_plan(int T, int G, int &L, int &nC, bool &psi_in_smem, size_t &smem, size_t budget = 200 * 1024) {
    const int NS = 32 / G;
    L = 64;
    if (T - 1 > 64 * 1024) L = (T - 1 + 1023) / 1024;
    nC = (T <= 1) ? 0 : (T - 1 + L - 1) / L;
    psi_in_smem = true;
    smem = vit_smem_bytes(NS, T, G, nC, true);
    if (smem > budget) {
        psi_in_smem = false;
        smem = vit_smem_bytes(NS, T, G, nC, false);
    }
}

#define HMMB200_DISPATCH_GK(FN, K, ...)                                                   \
    do {                                                                                  \
        const int kp_ = pad4(K);                                                          \
        if (kp_ <= 4) return FN<4, 4>(__VA_ARGS__);                                       \
        if (kp_ <= 8) return FN<8, 8>(__VA_ARGS__);                                       \
        if (kp_ <= 12) return FN<16, 12>(__VA_ARGS__);                                    \
        if (kp_ <= 16) return FN<16, 16>(__VA_ARGS__);                                    \
        if (kp_ <= 20) return FN<32, 20>(__VA_ARGS__);                                    \
        if (kp_ <= 24) return FN<32, 24>(__VA_ARGS__);                                    \
        if (kp_ <= 28) return FN<32, 28>(__VA_ARGS__);                                    \
        return FN<32, 32>(__VA_ARGS__);                                                   \
    } while (0)

inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }
// every chunk of every sequence starts and ends on a 16-byte boundary: the bulk-copy feed applies
inline bool bulk_feed_ok(const float *emis, int T, int K) { return (((uintptr_t)emis) & 15) == 0 && ((size_t)T * K) % 4 == 0; }
// value of the kernels' `bulk` parameter: 1 bulk-copy feed, 0 per-lane loads
inline int bulk_feed_param(const float *emis, int T, int K) { return bulk_feed_ok(emis, T, K) ? 1 : 0; }

// posterior / exp(log alpha) / exp(log beta) from the scaled sweeps left in the workspace
inline int launch_combine(const CombineParams &c, cudaStream_t s) {
    const size_t n = (size_t)c.n_frames;
    const int K = c.K;
    const int threads = 256;
    const unsigned blocks = (unsigned)((n + (size_t)threads * COMBINE_FPT - 1) / ((size_t)threads * COMBINE_FPT));
    auto al16 = [](const void *q) { return q == nullptr || ((uintptr_t)q & 15) == 0; };
    const bool vec = K % 4 == 0 && al16(c.gamma) && al16(c.fwd) && al16(c.bwd) && al16(c.log_alpha) && al16(c.log_beta);
    const unsigned wblocks = (unsigned)((n + 255) / 256);   // one 32-frame block per warp, 8 warps per CTA
    switch (vec ? K / 4 : 0) {
        case 1: fb_combine_warp_kernel<1><<<wblocks, threads, 0, s>>>(c); break;
        case 2: fb_combine_warp_kernel<2><<<wblocks, threads, 0, s>>>(c); break;
        case 3: fb_combine_warp_kernel<3><<<wblocks, threads, 0, s>>>(c); break;
        case 4: fb_combine_warp_kernel<4><<<wblocks, threads, 0, s>>>(c); break;
        case 5: fb_combine_warp_kernel<5><<<wblocks, threads, 0, s>>>(c); break;
        case 6: fb_combine_warp_kernel<6><<<wblocks, threads, 0, s>>>(c); break;
        case 7: fb_combine_warp_kernel<7><<<wblocks, threads, 0, s>>>(c); break;
        case 8: fb_combine_warp_kernel<8><<<wblocks, threads, 0, s>>>(c); break;
        default: fb_combine_kernel<<<blocks, threads, 0, s>>>(c); break;
    }
    return check_launch("fb_combine_kernel");
}

}  // namespace hmmb200
