// emission_tc.cu -- GMM emission log-likelihoods on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
//   l_kc(x) = const_kc + sum_d z_d * W1_kcd + z_d^2 * W2_kcd,   z = x - center,  W1 = (mu - center)/var,  W2 = -1/(2 var)
// is a dense [frames, 2D] x [2D, K*C] contraction.  fp32-grade accuracy on fp16 tensor-core inputs comes from a 3-term
// split: z = z_hi + z_lo, W = W_hi + W_lo (each fp16, 11 significant bits); all four partial products are accumulated.
// fp16 products are exact in the fp32 accumulator.
//
// One persistent CTA per SM, tile = 128 frames (= the 128 TMEM lanes):
//   warp 0        producer : TMA tensor copies (cp.async.bulk.tensor.2d, 128B swizzle) of the tile's frames, three
//                            [128 rows x 32 floats] boxes per tile, completion by mbarrier transaction bytes
//   warp 1        MMA      : one thread issues 4 * D/16 tcgen05.mma (A from TMEM, B = [W_hi; W_lo] from smem, N = 2*K*C,
//                            D accumulates in TMEM)
//   warps 2-9     transform: two groups of 4 warps; thread = frame row; centre, square, split into fp16 hi/lo pairs,
//                            tcgen05.st into the A buffer (each group takes every other 16-dim chunk)
//   warps 10-13   epilogue : tcgen05.ld the 128 x 2*K*C accumulator rows, add the hi/lo halves + const into a private smem
//                            row, release the accumulator, mixture log-sum-exp, store log b
// A and D are double-buffered in TMEM (2 x 2*DP + 2 x 2*NP columns <= 512) so transform(i+1), MMA(i) and epilogue(i-1)
// overlap; x stages are a 3-deep ring.  All hand-offs are mbarriers (tcgen05.commit for MMA completion).
//
// Range guard: fp16 needs |z| <= 240 (z^2 < 65504).  A frame outside that range is recomputed by its epilogue thread on
// the CUDA cores from the fp32 packed parameters; parameter sets whose W exceed the fp16 range are flagged at pack time
// and routed to the fp32 kernel (emission.cu).
#include "common.cuh"

#include <cuda.h>
#include <stdlib.h>
#include <cuda_fp16.h>

namespace hmmb200 {

constexpr int TC_TILE = 128;
constexpr int TC_STAGES = 3;
constexpr int TC_XF_GROUPS = 2;                                    // transform warp groups (4 warps each) per tile
constexpr int TC_EP_GROUPS = 1;                                    // epilogue warp groups (4 warps each)
constexpr int TC_THREADS = 32 * (2 + 4 * TC_XF_GROUPS + 4 * TC_EP_GROUPS);   // producer, MMA, transform groups, epilogue groups
constexpr float TC_ZMAX = 240.f;

// ---- raw PTX wrappers ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"      // suspends up to %2 ns: waiting warps do not burn issue slots
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity), "r"(1000000u) : "memory");
}
// 2-D TMA tile load: box (32 floats x 128 rows) at element coordinates (c0 = column, c1 = row); rows past the end of
// the tensor are zero-filled by the TMA unit.
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *tmap, int c0, int c1, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst)), "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T, kind::f16 (fp16 inputs, fp32 accumulate), M = 128, K = 16 per instruction
__device__ __forceinline__ void tc_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_st8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void tc_ld8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float fmax3(float a, float b, float c) {
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

// K-major, no-swizzle shared-memory matrix descriptor (sm_100 "version 1"): core matrix = 8 rows x 16 bytes,
//   LBO = byte distance between the two 16-byte K chunks of one K = 16 slice, SBO = byte distance between 8-row groups.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr >> 4) & 0x3fffu) | ((uint64_t)((lbo >> 4) & 0x3fffu) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3fffu) << 32) | (1ull << 46);
}

struct TcParams {
    const float *x;
    int64_t n_frames, n_tiles;
    int D, K, C, KC;
    int DP, NP;               // D and K*C padded to multiples of 16
    const float *tc;          // tensor-core section of the packed buffer (see tc_section_* below)
    const float *packed32;    // fp32 section (for out-of-range rows)
    int NP2;                  // component pairs in the fp32 section
    float *logb;
    float *comp;              // optional [n_frames, K*C] per-component log-likelihoods (Baum-Welch E-step) or null
    int dbg;                  // timing experiments only (HMMB200_TC_DBG): 1 skip MMA, 2 skip transform math, 4 skip epilogue math
};

// tensor-core section layout (floats): [0] usable flag, [4 .. 4+DP) centre, [.. +NP) const, then 4 fp16 matrices
// (W1_hi, W1_lo, W2_hi, W2_lo), each NP x DP halves in UMMA K-major no-swizzle order [n/8][k/8][n%8][k%8].
__host__ __device__ inline size_t tc_off_center() { return 4; }
__host__ __device__ inline size_t tc_off_const(int DP) { return 4 + (size_t)DP; }
__host__ __device__ inline size_t tc_off_w(int DP, int NP) { return (4 + (size_t)DP + NP + 3) & ~(size_t)3; }
__host__ __device__ inline size_t tc_section_floats(int DP, int NP) { return tc_off_w(DP, NP) + 2 * (size_t)NP * DP; }

// reference-private logsumexp over a state's C components read from a staging row (mixture_gaussian.py:141-155)
__device__ __forceinline__ float lse_row(const float *l, int C) {
    if (C == 1) return l[0];
    float m = l[0];
    for (int c = 1; c < C; ++c) m = fmaxf(m, l[c]);
    if (isinf(m)) m = 0.f;
    float s = 0.f;
    for (int c = 0; c < C; ++c) s += expf(l[c] - m);
    return logf(fmaxf(s, 1e-8f)) + m;
}

// timing trace (HMMB200_TC_DBG & 8): CTA 0 records clock64() at role hand-offs for its first 48 tiles
__device__ long long g_tc_trace[6][48][2];
#define TC_TRACE(role, ev) do { if ((p.dbg & 8) && blockIdx.x == 0 && it < 48 && lane == 0) g_tc_trace[role][it][ev] = clock64(); } while (0)

constexpr int TC_BOXW = 32;                                       // floats per TMA box row (= the 128-byte swizzle span)
constexpr int TC_BOX_FLOATS = TC_TILE * TC_BOXW;                  // 16 KB per box

// same, over C register values with the hardware exp2/log2 approximations (relative error ~2^-21; the sum is O(1..C))
template <int C>
__device__ __forceinline__ float lse_fast(const float *l) {
    float m = l[0];
#pragma unroll
    for (int c = 1; c < C; ++c) m = fmaxf(m, l[c]);
    if (isinf(m)) m = 0.f;
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) s += __expf(l[c] - m);
    return __logf(fmaxf(s, 1e-8f)) + m;
}

// WITH_COMP: also write the per-component values log w_kc + log N(x | mu_kc, var_kc) the epilogue holds before the mixture
// log-sum-exp (they are what the Baum-Welch E-step needs for the component responsibilities).
template <bool WITH_COMP>
__global__ void __launch_bounds__(TC_THREADS, 1) gmm_emission_tc_kernel(const __grid_constant__ CUtensorMap tmap, TcParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int D = p.D, DP = p.DP, NP = p.NP, K = p.K, C = p.C, KC = p.KC;
    if (p.tc[0] == 0.f) return;                                   // parameters outside the fp16 range: fp32 kernel runs instead
    const int NBOX = (D + TC_BOXW - 1) / TC_BOXW;                 // TMA boxes per tile
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // ---- carve shared memory ----
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem);          // x_full[S] x_empty[S] a_full[2] a_empty[2] d_full[2] d_empty[2]
    uint64_t *x_full = bars, *x_empty = bars + TC_STAGES, *a_full = bars + 2 * TC_STAGES, *a_empty = a_full + 2;
    uint64_t *d_full = a_empty + 2, *d_empty = d_full + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * TC_STAGES + 8);
    size_t off = 128;
    __half *wsm = reinterpret_cast<__half *>(smem + off);         off += (size_t)4 * NP * DP * sizeof(__half);
    float *cst_s = reinterpret_cast<float *>(smem + off);         off += (size_t)NP * sizeof(float);
    float *ctr_s = reinterpret_cast<float *>(smem + off);         off += (size_t)DP * sizeof(float);
    uint8_t *bad_s = smem + off;                                  off += 4 * TC_XF_GROUPS * TC_TILE;   // 4-deep ring x groups
    off = (off + 15) & ~(size_t)15;
    float *est = reinterpret_cast<float *>(smem + off);           off += (size_t)TC_TILE * (NP + 1) * sizeof(float);
    off += (1024u - ((smem_u32(smem) + (uint32_t)off) & 1023u)) & 1023u;   // 128B-swizzled TMA destinations: 1024-byte aligned ADDRESS
    float *xs = reinterpret_cast<float *>(smem + off);            // [S][NBOX][128][32], chunk c of row r at (c ^ (r & 7))

    // ---- one-time setup ----
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(p.tc + tc_off_w(DP, NP));
        uint4 *dst = reinterpret_cast<uint4 *>(wsm);
        const int n16 = 4 * NP * DP * 2 / 16;
        for (int i = threadIdx.x; i < n16; i += TC_THREADS) dst[i] = __ldg(src + i);
        for (int i = threadIdx.x; i < NP; i += TC_THREADS) cst_s[i] = __ldg(p.tc + tc_off_const(DP) + i);
        for (int i = threadIdx.x; i < DP; i += TC_THREADS) ctr_s[i] = __ldg(p.tc + tc_off_center() + i);
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(x_full + s, 1); mbar_init(x_empty + s, 128 * TC_XF_GROUPS); }
        for (int a = 0; a < 2; ++a) { mbar_init(a_full + a, 128 * TC_XF_GROUPS); mbar_init(a_empty + a, 1); mbar_init(d_full + a, 1); mbar_init(d_empty + a, 128 * TC_EP_GROUPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {                                              // TMEM: 512 columns, allocated (and later freed) by warp 2
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // W tile written with generic stores, read by the MMA (async proxy)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t ACOLS = 2 * DP;                                // columns of one A buffer: 4 segments x DP/2
    const uint32_t a_col0 = 0, d_col0 = 2 * ACOLS;                // [A0][A1][D0][D1]
    const int n_my = (p.n_tiles > blockIdx.x) ? (int)((p.n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;

    if (warp == 0) {
        // ================= producer: one elected thread issues the TMA tile loads =================
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
            for (int it = 0; it < n_my; ++it) {
                const int64_t tile = blockIdx.x + (int64_t)it * gridDim.x;
                const int s = it % TC_STAGES;
                mbar_wait(x_empty + s, ((it / TC_STAGES) & 1) ^ 1);
                TC_TRACE(0, 0);
                mbar_expect_tx(x_full + s, (uint32_t)NBOX * TC_BOX_FLOATS * sizeof(float));
                float *stage = xs + (size_t)s * NBOX * TC_BOX_FLOATS;
                for (int bx = 0; bx < NBOX; ++bx)
                    tma_load_2d(stage + (size_t)bx * TC_BOX_FLOATS, &tmap, bx * TC_BOXW, (int)(tile * TC_TILE), x_full + s);
                TC_TRACE(0, 1);
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (one thread) =================
        // The issue loop is a single thread's dependent instruction stream, so it is kept minimal: descriptor words
        // are precomputed, a K = 16 step advances the B descriptor's address field by 256 B >> 4 and the A column by 8.
        if (lane == 0) {
            const uint32_t idesc = (1u << 4) | ((uint32_t)((2 * NP) >> 3) << 17) | ((uint32_t)(TC_TILE >> 4) << 24);   // f16 x f16 -> f32, M = 128, N = 2*NP, K-major A and B
            const uint32_t lbo = 128, sbo = (uint32_t)(DP / 8) * 128;
            const uint32_t wbytes = (uint32_t)NP * DP * sizeof(__half);
            const uint32_t w_addr = smem_u32(wsm);
            const uint32_t desc_hi = (uint32_t)(make_smem_desc(0, lbo, sbo) >> 32);
            uint32_t wlo[4];
#pragma unroll
            for (int w = 0; w < 4; ++w) wlo[w] = (uint32_t)make_smem_desc(w_addr + w * wbytes, lbo, sbo);
            const int KS = (p.dbg & 1) ? 0 : DP / 16;
            const uint32_t half = DP / 2;
            auto mma = [&](uint32_t d_addr, uint32_t a_addr, uint32_t b_lo, int kk, uint32_t acc) {
                const uint64_t bdesc = ((uint64_t)desc_hi << 32) | (uint64_t)(b_lo + kk * 16);
                tc_mma_ts(d_addr, a_addr + kk * 8, bdesc, idesc, acc);
            };
            for (int it = 0; it < n_my; ++it) {
                const int a = it & 1;
                mbar_wait(a_full + a, (it >> 1) & 1);
                mbar_wait(d_empty + a, ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                TC_TRACE(1, 0);
                const uint32_t a_base = tmem_base + a_col0 + a * ACOLS;
                // One tcgen05.mma costs ~100 cycles here almost independently of N (N <= 96), so W_hi and W_lo are stacked
                // along N: the two matrices are adjacent in shared memory with the same 8-row-group stride, i.e. they ARE
                // one 2*NP-row K-major operand.  D[:, 0:NP] accumulates A*W_hi and D[:, NP:2NP] accumulates A*W_lo (the
                // epilogue adds them); 4 * D/16 instructions per tile instead of 6 * D/16, and the lo*lo term comes free.
                const uint32_t d_addr = tmem_base + d_col0 + (a * 2) * NP;
                // A segments (columns): z_hi [0,DP/2)  z_lo [DP/2,DP)  q_hi [DP,3DP/2)  q_lo [3DP/2,2DP);  W: 0 W1_hi 1 W1_lo 2 W2_hi 3 W2_lo
                for (int kk = 0; kk < KS; ++kk) {
                    mma(d_addr, a_base + 0 * half, wlo[0], kk, kk > 0 ? 1u : 0u);   // z_hi * [W1_hi ; W1_lo]
                    mma(d_addr, a_base + 2 * half, wlo[2], kk, 1u);                 // q_hi * [W2_hi ; W2_lo]
                    mma(d_addr, a_base + 1 * half, wlo[0], kk, 1u);                 // z_lo * [W1_hi ; W1_lo]
                    mma(d_addr, a_base + 3 * half, wlo[2], kk, 1u);                 // q_lo * [W2_hi ; W2_lo]
                }
                tc_commit(a_empty + a);                           // A buffer free once these MMAs retire
                tc_commit(d_full + a);                            // accumulator ready for the epilogue
                TC_TRACE(1, 1);
            }
        }
    } else if (warp < 2 + 4 * TC_XF_GROUPS) {
        // ================= transform: thread = frame row -> fp16 hi/lo pairs of z and z^2 into TMEM =================
        // TC_XF_GROUPS warp groups share a tile: group g converts the 16-dim chunks ch = g, g + TC_XF_GROUPS, ...
        const int q = warp & 3;                                   // TMEM lane quarter this warp may access
        const int grp = (warp - 2) >> 2;
        const int row = q * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        const int sw = row & 7;                                   // 128B swizzle: 16-byte chunk c of row r sits at c ^ (r & 7)
        for (int it = 0; it < n_my; ++it) {
            const int s = it % TC_STAGES, a = it & 1;
            mbar_wait(x_full + s, (it / TC_STAGES) & 1);
            if (q == 2) TC_TRACE(2 + grp, 0);
            mbar_wait(a_empty + a, ((it >> 1) & 1) ^ 1);
            tc_fence_after();
            const float *srow = xs + (size_t)s * NBOX * TC_BOX_FLOATS + row * TC_BOXW;
            const uint32_t a_base = tmem_base + lane_addr + a_col0 + a * ACOLS;
            float qmax = 0.f;
            for (int ch = grp; ch < ((p.dbg & 2) ? 0 : DP / 16); ch += TC_XF_GROUPS) {   // 16 dims -> 8 packed columns per segment
                // rows past the end of x and columns past D were zero-filled by the TMA unit (and the centre is 0 there)
                float z[16];
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    const int cg = ch * 4 + v;                    // global 16-byte chunk index of the row
                    const float4 t = *reinterpret_cast<const float4 *>(srow + (size_t)(cg >> 3) * TC_BOX_FLOATS + (((cg & 7) ^ sw) << 2));
                    const float4 c4 = *reinterpret_cast<const float4 *>(ctr_s + cg * 4);
                    z[4 * v + 0] = t.x - c4.x; z[4 * v + 1] = t.y - c4.y; z[4 * v + 2] = t.z - c4.z; z[4 * v + 3] = t.w - c4.w;
                }
                uint32_t zh[8], zl[8], qh[8], ql[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const float z0 = z[2 * e], z1 = z[2 * e + 1];
                    const __half2 h = __floats2half2_rn(z0, z1);
                    const float2 hf = __half22float2(h);
                    const __half2 l = __floats2half2_rn(z0 - hf.x, z1 - hf.y);
                    const float q0 = z0 * z0, q1 = z1 * z1;
                    qmax = fmax3(qmax, q0, q1);
                    const __half2 g = __floats2half2_rn(q0, q1);
                    const float2 gf = __half22float2(g);
                    const __half2 m = __floats2half2_rn(q0 - gf.x, q1 - gf.y);
                    zh[e] = *reinterpret_cast<const uint32_t *>(&h); zl[e] = *reinterpret_cast<const uint32_t *>(&l);
                    qh[e] = *reinterpret_cast<const uint32_t *>(&g); ql[e] = *reinterpret_cast<const uint32_t *>(&m);
                }
                tc_st8(a_base + 0 * (DP / 2) + ch * 8, zh);
                tc_st8(a_base + 1 * (DP / 2) + ch * 8, zl);
                tc_st8(a_base + 2 * (DP / 2) + ch * 8, qh);
                tc_st8(a_base + 3 * (DP / 2) + ch * 8, ql);
            }
            // a frame with |z| > 240 (z^2 beyond fp16) or NaN is recomputed in fp32 by its epilogue thread
            bad_s[((it & 3) * TC_XF_GROUPS + grp) * TC_TILE + row] = (uint8_t)(!(qmax <= TC_ZMAX * TC_ZMAX) ? 1 : 0);
            tc_wait_st();
            tc_fence_before();
            mbar_arrive(a_full + a);
            mbar_arrive(x_empty + s);
            if (q == 2) TC_TRACE(2 + grp, 1);
        }
    } else {
        // ================= epilogue: accumulator rows -> + const -> mixture log-sum-exp -> log b =================
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        float *my = est + (size_t)row * (NP + 1);
        const bool fast = (16 % C) == 0;                          // every 16-column accumulator chunk holds whole states
        const bool vec_out = (K % 4) == 0;
        for (int it = 0; it < n_my; ++it) {
            const int64_t tile = blockIdx.x + (int64_t)it * gridDim.x;
            const int a = it & 1;
            mbar_wait(d_full + a, (it >> 1) & 1);
            tc_fence_after();
            if (q == 2) TC_TRACE(4, 0);
            const uint32_t d_addr = tmem_base + lane_addr + d_col0 + (a * 2) * NP;      // A*W_hi part; the A*W_lo part follows at + NP
            // phase 1: drain the accumulator row (hi + lo halves + const) into this thread's private smem row and release
            // the TMEM buffer at once, so the next MMA never waits for the log-sum-exp arithmetic
            for (int ch = 0; ch < NP / 16; ++ch) {
                uint32_t v0[16], v1[16];
                tc_ld16(d_addr + ch * 16, v0);
                tc_ld16(d_addr + NP + ch * 16, v1);
                // constants into registers BEFORE the stores: cst_s and the staging row are both shared memory, and the
                // compiler must otherwise serialise every (load const, add, store) triple for fear of aliasing
                float4 c0[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) c0[i] = *reinterpret_cast<const float4 *>(cst_s + ch * 16 + 4 * i);
                tc_wait_ld();
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    my[ch * 16 + 4 * i + 0] = (__uint_as_float(v0[4 * i + 0]) + __uint_as_float(v1[4 * i + 0])) + c0[i].x;
                    my[ch * 16 + 4 * i + 1] = (__uint_as_float(v0[4 * i + 1]) + __uint_as_float(v1[4 * i + 1])) + c0[i].y;
                    my[ch * 16 + 4 * i + 2] = (__uint_as_float(v0[4 * i + 2]) + __uint_as_float(v1[4 * i + 2])) + c0[i].z;
                    my[ch * 16 + 4 * i + 3] = (__uint_as_float(v0[4 * i + 3]) + __uint_as_float(v1[4 * i + 3])) + c0[i].w;
                }
            }
            // flag ring is 4 deep: transform(it+4) can only start after MMA(it+2), which waits for this d_empty arrive
            bool bad = false;
#pragma unroll
            for (int g = 0; g < TC_XF_GROUPS; ++g) bad |= bad_s[((it & 3) * TC_XF_GROUPS + g) * TC_TILE + row] != 0;
            tc_fence_before();
            mbar_arrive(d_empty + a);
            if (q == 2) TC_TRACE(4, 1);
            if (WITH_COMP) {
                const int64_t fr = tile * TC_TILE + row;
                if (fr < p.n_frames && !bad) {
                    float *co = p.comp + fr * KC;
                    if ((KC & 3) == 0) {
                        for (int kc = 0; kc < KC; kc += 4)
                            *reinterpret_cast<float4 *>(co + kc) = make_float4(my[kc], my[kc + 1], my[kc + 2], my[kc + 3]);
                    } else {
                        for (int kc = 0; kc < KC; ++kc) co[kc] = my[kc];
                    }
                }
            }
            // phase 2: the reference's private logsumexp (mixture_gaussian.py:141-155) per state, in place
            if (fast && !bad && C > 1 && !(p.dbg & 4)) {
                for (int ch = 0; ch < NP / 16; ++ch) {
                    float l[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) l[i] = my[ch * 16 + i];
                    if (C == 2) {
#pragma unroll
                        for (int g = 0; g < 8; ++g) my[ch * 8 + g] = lse_fast<2>(&l[2 * g]);
                    } else if (C == 4) {
#pragma unroll
                        for (int g = 0; g < 4; ++g) my[ch * 4 + g] = lse_fast<4>(&l[4 * g]);
                    } else if (C == 8) {
#pragma unroll
                        for (int g = 0; g < 2; ++g) my[ch * 2 + g] = lse_fast<8>(&l[8 * g]);
                    } else {
                        my[ch] = lse_fast<16>(&l[0]);
                    }
                }
            }
            const int64_t frame = tile * TC_TILE + row;
            if (frame < p.n_frames) {
                float *o = p.logb + frame * K;
                if (bad) {
                    // out-of-range frame: fp32 CUDA-core recomputation from the standardised parameters
                    const float *xg = p.x + frame * D;
                    const float *cst32 = p.packed32 + (size_t)D * p.NP2 * 4;
                    for (int kc = 0; kc < KC; ++kc) {
                        const int pr = kc >> 1, hi = kc & 1;
                        float acc = 0.f;
                        for (int d = 0; d < D; ++d) {
                            const float u = fmaf(xg[d], __ldg(p.packed32 + ((size_t)d * p.NP2 + pr) * 4 + hi),
                                                 __ldg(p.packed32 + ((size_t)d * p.NP2 + pr) * 4 + 2 + hi));
                            acc = fmaf(u, u, acc);
                        }
                        my[kc] = fmaf(-0.5f, acc, cst32[kc]);
                        if (WITH_COMP) p.comp[frame * KC + kc] = my[kc];
                    }
                    for (int k = 0; k < K; ++k) o[k] = lse_row(my + k * C, C);
                } else if (!fast) {
                    for (int k = 0; k < K; ++k) o[k] = lse_row(my + k * C, C);
                } else if (vec_out) {
                    for (int k = 0; k < K; k += 4)
                        *reinterpret_cast<float4 *>(o + k) = make_float4(my[k], my[k + 1], my[k + 2], my[k + 3]);
                } else {
                    for (int k = 0; k < K; ++k) o[k] = my[k];
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// ---- packing of the tensor-core section ----------------------------------------------------------------------
__global__ void gmm_pack_tc_center_kernel(const float *means, int KC, int D, int DP, float *tc) {
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d == 0) tc[0] = 1.f;
    if (d >= DP) return;
    double s = 0.0;
    if (d < D) for (int kc = 0; kc < KC; ++kc) s += (double)means[(size_t)kc * D + d];
    tc[tc_off_center() + d] = (d < D) ? (float)(s / KC) : 0.f;
}

__global__ void gmm_pack_tc_kernel(const float *means, const float *log_vars, float scale, const float *logw,
                                   int KC, int D, int DP, int NP, float *tc) {
    const int kc = blockIdx.x * blockDim.x + threadIdx.x;
    if (kc >= NP) return;
    __half *w = reinterpret_cast<__half *>(tc + tc_off_w(DP, NP));
    const size_t wsz = (size_t)NP * DP;
    double cst = 0.0;
    bool unsafe = false;
    for (int d = 0; d < DP; ++d) {
        float w1 = 0.f, w2 = 0.f;
        if (kc < KC && d < D) {
            const double lv = (double)scale * (double)log_vars[(size_t)kc * D + d];
            const double iv = exp(-lv);
            const double dm = (double)means[(size_t)kc * D + d] - (double)tc[tc_off_center() + d];
            w1 = (float)(dm * iv);
            w2 = (float)(-0.5 * iv);
            cst += -0.5 * (lv + 1.8378770664093454835606594728112) - 0.5 * dm * dm * iv;
            if (!(fabsf(w1) < 60000.f) || !(fabsf(w2) < 60000.f)) unsafe = true;
        }
        const __half h1 = __float2half_rn(w1), h2 = __float2half_rn(w2);
        const __half l1 = __float2half_rn(w1 - __half2float(h1)), l2 = __float2half_rn(w2 - __half2float(h2));
        const size_t idx = (((size_t)(kc >> 3) * (DP / 8) + (d >> 3)) * 8 + (kc & 7)) * 8 + (d & 7);
        w[0 * wsz + idx] = h1; w[1 * wsz + idx] = l1; w[2 * wsz + idx] = h2; w[3 * wsz + idx] = l2;
    }
    if (kc < KC) cst += logw ? (double)logw[kc] : 0.0;
    tc[tc_off_const(DP) + kc] = (kc < KC) ? (float)cst : 0.f;
    if (unsafe || !(fabs(cst) < 3.0e38)) tc[0] = 0.f;
}

bool tc_shape_ok(int K, int C, int D) {
    const int KC = K * C;
    if (!(D % 4 == 0 && D >= 4 && D <= 80 && KC >= 1 && KC <= 96)) return false;
    const int DP = (D + 15) & ~15, NP = (KC + 15) & ~15;
    return 4 * (DP + NP) <= 512;                                    // TMEM: 2 A buffers (2*DP cols) + 2 x 2 accumulators (NP cols)
}
size_t tc_floats(int K, int C, int D) {
    if (!tc_shape_ok(K, C, D)) return 0;
    const int DP = (D + 15) & ~15, NP = (K * C + 15) & ~15;
    return tc_section_floats(DP, NP);
}

int launch_pack_tc(const float *means, const float *log_vars, float scale, const float *logw, int K, int C, int D,
                   float *tc, cudaStream_t s) {
    const int KC = K * C, DP = (D + 15) & ~15, NP = (KC + 15) & ~15;
    gmm_pack_tc_center_kernel<<<(DP + 63) / 64, 64, 0, s>>>(means, KC, D, DP, tc);
    if (int rc = check_launch("gmm_pack_tc_center_kernel")) return rc;
    gmm_pack_tc_kernel<<<(NP + 31) / 32, 32, 0, s>>>(means, log_vars, scale, logw, KC, D, DP, NP, tc);
    return check_launch("gmm_pack_tc_kernel");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda).
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)ptr;
        else
            cudaGetLastError();
    }
    return fn;
}

int launch_emission_tc(const float *x, const float *tc, const float *packed32, int64_t n_frames, int K, int C, int D,
                       float *logb, cudaStream_t s, float *comp) {
    TcParams p;
    p.x = x; p.n_frames = n_frames; p.n_tiles = (n_frames + TC_TILE - 1) / TC_TILE;
    p.D = D; p.K = K; p.C = C; p.KC = K * C;
    p.DP = (D + 15) & ~15; p.NP = (p.KC + 15) & ~15;
    p.tc = tc; p.packed32 = packed32; p.NP2 = (p.KC + 1) / 2; p.logb = logb; p.comp = comp;
    { const char *e = getenv("HMMB200_TC_DBG"); p.dbg = e ? atoi(e) : 0; }
    const int nbox = (D + TC_BOXW - 1) / TC_BOXW;
    size_t smem = 128 + (size_t)4 * p.NP * p.DP * 2 + (size_t)(p.NP + p.DP) * 4 + 4 * TC_XF_GROUPS * TC_TILE;
    smem = (smem + 15) & ~(size_t)15;
    smem += (size_t)TC_TILE * (p.NP + 1) * 4;
    smem = (smem + 1023) & ~(size_t)1023;
    smem += (size_t)TC_STAGES * nbox * TC_BOX_FLOATS * 4;
    smem += 1024;                                                   // slack: the dynamic window itself is only 16-byte aligned
    if (smem > 227 * 1024) return 1;
    EncodeTiledFn encode = encode_tiled_fn();
    if (encode == nullptr) return 1;
    // x as a 2-D tensor [n_frames rows][D floats]; box = 32 floats x 128 rows, 128-byte swizzle, zero fill out of bounds
    CUtensorMap tmap;
    const cuuint64_t gdim[2] = {(cuuint64_t)D, (cuuint64_t)n_frames};
    const cuuint64_t gstride[1] = {(cuuint64_t)D * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)TC_BOXW, (cuuint32_t)TC_TILE};
    const cuuint32_t estr[2] = {1, 1};
    CUresult cr = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(x), gdim, gstride, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return 1;
    cudaError_t e = cudaFuncSetAttribute(gmm_emission_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gmm_emission_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_error(HMMB200_ELAUNCH, "emission_tc smem opt-in: %s", cudaGetErrorString(e));
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = (int)min((int64_t)sms, p.n_tiles);
    if (comp != nullptr) gmm_emission_tc_kernel<true><<<grid, TC_THREADS, smem, s>>>(tmap, p);
    else gmm_emission_tc_kernel<false><<<grid, TC_THREADS, smem, s>>>(tmap, p);
    return check_launch("gmm_emission_tc_kernel");
}

}  // namespace hmmb200

// debug aid (not part of the public ABI): copies the timing trace of the last traced launch to the host
HMMB200_EXPORT int hmmb200_debug_tc_trace(long long *out) {
    return cudaMemcpyFromSymbol(out, hmmb200::g_tc_trace, sizeof(long long) * 6 * 48 * 2) == cudaSuccess ? 0 : -1;
}
