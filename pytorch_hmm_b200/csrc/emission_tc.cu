// emission_tc.cu -- GMM emission log-likelihoods on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
//   l_kc(x) = const_kc + sum_d z_d * W1_kcd + z_d^2 * W2_kcd,   z = x - center,  W1 = (mu - center)/var,  W2 = -1/(2 var)
// is a dense [frames, 2D] x [2D, K*C] contraction.  fp32-grade accuracy on fp16 tensor-core inputs comes from a 3-term
// split: z = z_hi + z_lo, W = W_hi + W_lo (each fp16, 11 significant bits); all four partial products are accumulated.
// fp16 products are exact in the fp32 accumulator.  W and const are stored pre-multiplied by log2(e), so the accumulator is
// in log2 units and the mixture log-sum-exp runs on the hardware exp2 / log2 without per-term scaling.
//
// One persistent CTA per SM, tile = 128 frames (= the 128 TMEM lanes):
//   warp 0        producer : TMA tensor copies (cp.async.bulk.tensor.2d, 128B swizzle) of the tile's frames, three
//                            [128 rows x 32 floats] boxes per tile, completion by mbarrier transaction bytes
//   warp 1        MMA      : one thread issues 4 * D/16 tcgen05.mma (A from TMEM, B = [W_hi; W_lo] from smem, N = 2*K*C,
//                            D accumulates in TMEM)
//   warps 2-9     transform: two groups of 4 warps; thread = frame row; centre, square, split into fp16 hi/lo pairs,
//                            tcgen05.st into the A buffer (each group takes every other 8-dim half-chunk)
//   warps 10-13   epilogue : tcgen05.ld the 128 x 2*K*C accumulator row 16 columns at a time, hi + lo halves + const, mixture
//                            log-sum-exp and the store of log b, all in registers
// A and D are double-buffered in TMEM (2 x 2*DP + 2 x 2*NP columns <= 512) so transform(i+1), MMA(i) and epilogue(i-1)
// overlap; x stages are a 3-deep ring.  All hand-offs are mbarriers (tcgen05.commit for MMA completion).
//
// The kernel is bound by the SM's instruction issue, so the per-element instruction count is what was engineered:
//   transform, per PAIR of dims (10 instructions): packed fp32x2 subtract of the centre, cvt.rn.f16x2.f32 (hi halves),
//     two mixed-precision subtracts float(hi) - z (sub.rn.f32.f16: no unpacking), cvt.rn.f16x2.f32 (the NEGATED lo halves: the
//     MMAs of the lo segments run with the descriptor's negate-A bit), packed fp32x2 square, and the same three steps for z^2;
//   epilogue, per 16 accumulator columns: two tcgen05.ld, 16 packed adds, C-way max / exp2 / sum / log2 per state.
// Range guard: fp16 holds |z| < 65520 and z^2 < 65520.  Beyond that the converted operand is +-inf, the accumulator row becomes
// inf / NaN, the epilogue thread sees a non-finite output and recomputes ITS frame on the CUDA cores from the fp32 packed
// parameters.  Parameter sets whose W exceed the fp16 range are flagged at pack time and routed to the fp32 kernel (emission.cu).
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
constexpr float TC_LN2 = 0.69314718055994530942f;

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
__device__ __forceinline__ void tc_st4(uint32_t taddr, const uint32_t (&v)[4]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]) : "memory");
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
// packed fp32 pairs: one issue slot for two IEEE round-to-nearest operations
__device__ __forceinline__ float2 tc_add2(float2 a, float2 b) {
    unsigned long long ra = *reinterpret_cast<unsigned long long *>(&a), rb = *reinterpret_cast<unsigned long long *>(&b), rd;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    return *reinterpret_cast<float2 *>(&rd);
}
__device__ __forceinline__ float2 tc_mul2(float2 a, float2 b) {
    unsigned long long ra = *reinterpret_cast<unsigned long long *>(&a), rb = *reinterpret_cast<unsigned long long *>(&b), rd;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    return *reinterpret_cast<float2 *>(&rd);
}
// (lo, hi) floats -> packed fp16 pair, round to nearest (values beyond the fp16 range become +-inf: the range guard)
__device__ __forceinline__ uint32_t tc_pack_h2(float lo, float hi) {
    uint32_t d;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}
// float(h.lo) - a and float(h.hi) - b in one mixed-precision instruction each (no unpacking of the halves)
__device__ __forceinline__ float2 tc_h2_minus(uint32_t h, float a, float b) {
    float2 r;
    asm("{\n\t"
        ".reg .b16 l, u;\n\t"
        "mov.b32 {l, u}, %2;\n\t"
        "sub.rn.f32.f16 %0, l, %3;\n\t"
        "sub.rn.f32.f16 %1, u, %4;\n\t"
        "}" : "=f"(r.x), "=f"(r.y) : "r"(h), "f"(a), "f"(b));
    return r;
}
// hardware exp2 / log2 (relative error ~2^-22; the sums below are in [1, C])
__device__ __forceinline__ float tc_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float tc_lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

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
#ifdef HMMB200_DEBUG_HOOKS
    int dbg;                  // timing experiments only (HMMB200_TC_DBG): 1 skip MMA, 2 skip transform math, 4 skip epilogue math, 8 trace
#endif
};

// tensor-core section layout (floats): [0] usable flag, [4 .. 4+DP) centre, [.. +NP) const (log2 units), then 4 fp16 matrices
// (W1_hi, W1_lo, W2_hi, W2_lo; log2 units), each NP x DP halves in UMMA K-major no-swizzle order [n/8][k/8][n%8][k%8].
__host__ __device__ inline size_t tc_off_center() { return 4; }
__host__ __device__ inline size_t tc_off_const(int DP) { return 4 + (size_t)DP; }
__host__ __device__ inline size_t tc_off_w(int DP, int NP) { return (4 + (size_t)DP + NP + 3) & ~(size_t)3; }
__host__ __device__ inline size_t tc_section_floats(int DP, int NP) { return tc_off_w(DP, NP) + 2 * (size_t)NP * DP; }

// reference-private logsumexp over a state's C components read from memory, natural-log units (mixture_gaussian.py:141-155)
__device__ __forceinline__ float lse_row(const float *l, int C) {
    if (C == 1) return l[0];
    float m = l[0];
    for (int c = 1; c < C; ++c) m = fmaxf(m, l[c]);
    if (isinf(m)) m = 0.f;
    float s = 0.f;
    for (int c = 0; c < C; ++c) s += expf(l[c] - m);
    return logf(fmaxf(s, 1e-8f)) + m;
}

#ifdef HMMB200_DEBUG_HOOKS
// timing trace (HMMB200_TC_DBG & 8, debug builds only): CTA 0 records clock64() at role hand-offs for its first 48 tiles
__device__ long long g_tc_trace[6][48][2];
#define TC_TRACE(role, ev) do { if ((p.dbg & 8) && blockIdx.x == 0 && it < 48 && lane == 0) g_tc_trace[role][it][ev] = clock64(); } while (0)
#define TC_DBG(bit) (p.dbg & (bit))
#else
#define TC_TRACE(role, ev) do { } while (0)
#define TC_DBG(bit) 0
#endif

constexpr int TC_BOXW = 32;                                       // floats per TMA box row (= the 128-byte swizzle span)
constexpr int TC_BOX_FLOATS = TC_TILE * TC_BOXW;                  // 16 KB per box

// mixture log-sum-exp of CT register values in log2 units -> natural-log result.  The largest term contributes exp2(0) = 1, so
// the reference's clamp(sum, 1e-8) (mixture_gaussian.py:150) can never bind on finite inputs; non-finite rows are recomputed.
template <int CT>
__device__ __forceinline__ float lse2_regs(const float *l) {
    if (CT == 1) return l[0] * TC_LN2;
    float m;
    if (CT == 2) m = fmaxf(l[0], l[1]);
    else if (CT == 4) m = fmaxf(fmax3(l[0], l[1], l[2]), l[3]);
    else m = fmax3(fmax3(l[0], l[1], l[2]), fmax3(l[3], l[4], l[5]), fmaxf(l[6], l[7]));
    const float2 nm = make_float2(-m, -m);
    float e[CT];
#pragma unroll
    for (int c = 0; c < CT; c += 2) {
        const float2 d = tc_add2(make_float2(l[c], l[c + 1]), nm);
        e[c] = tc_ex2(d.x); e[c + 1] = tc_ex2(d.y);
    }
    float s;
    if (CT == 2) s = e[0] + e[1];
    else if (CT == 4) s = (e[0] + e[1]) + (e[2] + e[3]);
    else s = ((e[0] + e[1]) + (e[2] + e[3])) + ((e[4] + e[5]) + (e[6] + e[7]));
    return (tc_lg2(s) + m) * TC_LN2;
}

// WITH_COMP: also write the per-component values log w_kc + log N(x | mu_kc, var_kc) the epilogue holds before the mixture
// log-sum-exp (they are what the Baum-Welch E-step needs for the component responsibilities).
// CT: mixture components per state when every 16-column accumulator chunk holds whole states (1, 2, 4, 8), else 0 = any C
// through a shared-memory staging row.
template <bool WITH_COMP, int CT>
__global__ void __launch_bounds__(TC_THREADS, 1) gmm_emission_tc_kernel(const __grid_constant__ CUtensorMap tmap, TcParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int D = p.D, DP = p.DP, NP = p.NP, K = p.K, C = p.C, KC = p.KC;
    // A dependent kernel launched with programmatic stream serialisation (the fused recursion kernel) may be scheduled as soon as
    // SMs free up: its set-up then overlaps this kernel's last tiles; it waits for this grid's completion before reading log b.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
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
    float *nctr_s = reinterpret_cast<float *>(smem + off);        off += (size_t)DP * sizeof(float);       // NEGATED centre
    off = (off + 15) & ~(size_t)15;
    float *est = reinterpret_cast<float *>(smem + off);           // generic-C staging rows (CT == 0 only)
    if (CT == 0) off += (size_t)TC_TILE * (NP + 1) * sizeof(float);
    off += (1024u - ((smem_u32(smem) + (uint32_t)off) & 1023u)) & 1023u;   // 128B-swizzled TMA destinations: 1024-byte aligned ADDRESS
    float *xs = reinterpret_cast<float *>(smem + off);            // [S][NBOX][128][32], chunk c of row r at (c ^ (r & 7))

    // ---- one-time setup ----
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(p.tc + tc_off_w(DP, NP));
        uint4 *dst = reinterpret_cast<uint4 *>(wsm);
        const int n16 = 4 * NP * DP * 2 / 16;
        for (int i = threadIdx.x; i < n16; i += TC_THREADS) dst[i] = __ldg(src + i);
        for (int i = threadIdx.x; i < NP; i += TC_THREADS) cst_s[i] = __ldg(p.tc + tc_off_const(DP) + i);
        for (int i = threadIdx.x; i < DP; i += TC_THREADS) nctr_s[i] = -__ldg(p.tc + tc_off_center() + i);
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
            const uint32_t idesc_nega = idesc | (1u << 13);       // same with A negated: the lo segments hold -z_lo, -q_lo
            const uint32_t lbo = 128, sbo = (uint32_t)(DP / 8) * 128;
            const uint32_t wbytes = (uint32_t)NP * DP * sizeof(__half);
            const uint32_t w_addr = smem_u32(wsm);
            const uint32_t desc_hi = (uint32_t)(make_smem_desc(0, lbo, sbo) >> 32);
            uint32_t wlo[4];
#pragma unroll
            for (int w = 0; w < 4; ++w) wlo[w] = (uint32_t)make_smem_desc(w_addr + w * wbytes, lbo, sbo);
            const int KS = TC_DBG(1) ? 0 : DP / 16;
            const uint32_t half = DP / 2;
            auto mma = [&](uint32_t d_addr, uint32_t a_addr, uint32_t b_lo, int kk, uint32_t acc, uint32_t id) {
                const uint64_t bdesc = ((uint64_t)desc_hi << 32) | (uint64_t)(b_lo + kk * 16);
                tc_mma_ts(d_addr, a_addr + kk * 8, bdesc, id, acc);
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
                // A segments (columns): z_hi [0,DP/2)  -z_lo [DP/2,DP)  q_hi [DP,3DP/2)  -q_lo [3DP/2,2DP);  W: 0 W1_hi 1 W1_lo 2 W2_hi 3 W2_lo
                for (int kk = 0; kk < KS; ++kk) {
                    mma(d_addr, a_base + 0 * half, wlo[0], kk, kk > 0 ? 1u : 0u, idesc);        // z_hi * [W1_hi ; W1_lo]
                    mma(d_addr, a_base + 2 * half, wlo[2], kk, 1u, idesc);                      // q_hi * [W2_hi ; W2_lo]
                    mma(d_addr, a_base + 1 * half, wlo[0], kk, 1u, idesc_nega);                 // z_lo * [W1_hi ; W1_lo]
                    mma(d_addr, a_base + 3 * half, wlo[2], kk, 1u, idesc_nega);                 // q_lo * [W2_hi ; W2_lo]
                }
                tc_commit(a_empty + a);                           // A buffer free once these MMAs retire
                tc_commit(d_full + a);                            // accumulator ready for the epilogue
                TC_TRACE(1, 1);
            }
        }
    } else if (warp < 2 + 4 * TC_XF_GROUPS) {
        // ================= transform: thread = frame row -> fp16 hi/lo pairs of z and z^2 into TMEM =================
        // TC_XF_GROUPS warp groups share a tile: group g converts the 8-dim half-chunks hc = g, g + TC_XF_GROUPS, ...
        const int q = warp & 3;                                   // TMEM lane quarter this warp may access
        const int grp = (warp - 2) >> 2;
        const int row = q * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        const int sw = row & 7;                                   // 128B swizzle: 16-byte chunk c of row r sits at c ^ (r & 7)
        const int NHC = TC_DBG(2) ? 0 : DP / 8;
        const uint32_t half = DP / 2;
        // Half-chunk hc = grp + 2 i covers the 16-byte chunks cg = 2 hc, 2 hc + 1 of the row: box cg >> 3 = i >> 1, chunk-in-box
        // (2 grp) | 4 (i & 1) (+ 1) -- so the four swizzled offsets a thread ever uses are fixed and everything else is an
        // immediate of the unrolled loop (DP <= 80: at most 5 half-chunks per group).
        static_assert(TC_XF_GROUPS == 2, "the half-chunk -> (box, chunk) map below assumes two transform groups");
        const int g2 = 2 * grp;
        const int offA[2] = {((g2 ^ sw) & 7) << 2, (((g2 | 4) ^ sw) & 7) << 2};
        const int offB[2] = {(((g2 | 1) ^ sw) & 7) << 2, (((g2 | 5) ^ sw) & 7) << 2};
        const float *nc = nctr_s + grp * 8;
        for (int it = 0; it < n_my; ++it) {
            const int s = it % TC_STAGES, a = it & 1;
            mbar_wait(x_full + s, (it / TC_STAGES) & 1);
            if (q == 2) TC_TRACE(2 + grp, 0);
            mbar_wait(a_empty + a, ((it >> 1) & 1) ^ 1);
            tc_fence_after();
            const float *srow = xs + (size_t)s * NBOX * TC_BOX_FLOATS + row * TC_BOXW;
            const uint32_t a_base = tmem_base + lane_addr + a_col0 + a * ACOLS + grp * 4;
#pragma unroll
            for (int i = 0; i < 5; ++i) {                           // 8 dims -> 4 packed columns per segment
                if (grp + 2 * i >= NHC) break;
                // rows past the end of x and columns past D were zero-filled by the TMA unit (and the centre is 0 there)
                const float *bsrc = srow + (size_t)(i >> 1) * TC_BOX_FLOATS;
                const float4 t0 = *reinterpret_cast<const float4 *>(bsrc + offA[i & 1]);
                const float4 t1 = *reinterpret_cast<const float4 *>(bsrc + offB[i & 1]);
                const float4 c0 = *reinterpret_cast<const float4 *>(nc + i * 16);
                const float4 c1 = *reinterpret_cast<const float4 *>(nc + i * 16 + 4);
                float2 z[4];
                z[0] = tc_add2(make_float2(t0.x, t0.y), make_float2(c0.x, c0.y));
                z[1] = tc_add2(make_float2(t0.z, t0.w), make_float2(c0.z, c0.w));
                z[2] = tc_add2(make_float2(t1.x, t1.y), make_float2(c1.x, c1.y));
                z[3] = tc_add2(make_float2(t1.z, t1.w), make_float2(c1.z, c1.w));
                uint32_t zh[4], zl[4], qh[4], ql[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    zh[e] = tc_pack_h2(z[e].x, z[e].y);
                    const float2 nl = tc_h2_minus(zh[e], z[e].x, z[e].y);          // float(hi) - z = -lo, exact
                    zl[e] = tc_pack_h2(nl.x, nl.y);
                    const float2 qq = tc_mul2(z[e], z[e]);
                    qh[e] = tc_pack_h2(qq.x, qq.y);
                    const float2 nm = tc_h2_minus(qh[e], qq.x, qq.y);
                    ql[e] = tc_pack_h2(nm.x, nm.y);
                }
                tc_st4(a_base + 0 * half + i * 8, zh);
                tc_st4(a_base + 1 * half + i * 8, zl);
                tc_st4(a_base + 2 * half + i * 8, qh);
                tc_st4(a_base + 3 * half + i * 8, ql);
            }
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
        constexpr int OPC = CT > 0 ? 16 / CT : 1;                 // outputs (states) per 16-column accumulator chunk
        const bool vec4 = (K % 4) == 0 && (((uintptr_t)p.logb) & 15) == 0;
        const int NCH = NP / 16;
        for (int it = 0; it < n_my; ++it) {
            const int64_t tile = blockIdx.x + (int64_t)it * gridDim.x;
            const int a = it & 1;
            mbar_wait(d_full + a, (it >> 1) & 1);
            tc_fence_after();
            if (q == 2) TC_TRACE(4, 0);
            const uint32_t d_addr = tmem_base + lane_addr + d_col0 + (a * 2) * NP;      // A*W_hi part; the A*W_lo part follows at + NP
            const int64_t frame = tile * TC_TILE + row;
            const bool live = frame < p.n_frames;
            float *o = p.logb + (live ? frame : 0) * K;
            float *co = WITH_COMP ? p.comp + (live ? frame : 0) * KC : nullptr;
            float chk = 0.f;                                      // sum of the row's outputs: non-finite <=> some operand left the fp16 range
            float *my = est + (size_t)row * (NP + 1);             // CT == 0 only
#pragma unroll 1
            for (int ch = 0; ch < NCH; ++ch) {
                uint32_t v0[16], v1[16];
                tc_ld16(d_addr + ch * 16, v0);
                tc_ld16(d_addr + NP + ch * 16, v1);
                float4 c4[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) c4[i] = *reinterpret_cast<const float4 *>(cst_s + ch * 16 + 4 * i);
                tc_wait_ld();
                if (ch == NCH - 1) {                              // the row is in registers: hand the accumulator back to the MMA warp
                    tc_fence_before();
                    mbar_arrive(d_empty + a);
                    if (q == 2) TC_TRACE(4, 1);
                }
                float l[16];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float2 s0 = tc_add2(make_float2(__uint_as_float(v0[4 * i]), __uint_as_float(v0[4 * i + 1])),
                                              make_float2(__uint_as_float(v1[4 * i]), __uint_as_float(v1[4 * i + 1])));
                    const float2 s1 = tc_add2(make_float2(__uint_as_float(v0[4 * i + 2]), __uint_as_float(v0[4 * i + 3])),
                                              make_float2(__uint_as_float(v1[4 * i + 2]), __uint_as_float(v1[4 * i + 3])));
                    const float2 r0 = tc_add2(s0, make_float2(c4[i].x, c4[i].y));
                    const float2 r1 = tc_add2(s1, make_float2(c4[i].z, c4[i].w));
                    l[4 * i] = r0.x; l[4 * i + 1] = r0.y; l[4 * i + 2] = r1.x; l[4 * i + 3] = r1.y;
                }
                if (TC_DBG(4)) continue;
                if (WITH_COMP && live) {
                    if ((KC & 3) == 0) {
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            if (ch * 16 + 4 * i < KC)
                                *reinterpret_cast<float4 *>(co + ch * 16 + 4 * i) =
                                    make_float4(l[4 * i] * TC_LN2, l[4 * i + 1] * TC_LN2, l[4 * i + 2] * TC_LN2, l[4 * i + 3] * TC_LN2);
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; ++i) if (ch * 16 + i < KC) co[ch * 16 + i] = l[i] * TC_LN2;
                    }
                }
                if constexpr (CT > 0) {
                    float out[OPC];
#pragma unroll
                    for (int g = 0; g < OPC; ++g) out[g] = lse2_regs<CT>(&l[CT * g]);
                    const int k0 = ch * OPC;
#pragma unroll
                    for (int g = 0; g < OPC; ++g) if (k0 + g < K) chk += out[g];
                    if (live) {
                        if (OPC % 4 == 0 && vec4) {
#pragma unroll
                            for (int g = 0; g < OPC; g += 4)
                                if (k0 + g < K) *reinterpret_cast<float4 *>(o + k0 + g) = make_float4(out[g], out[g + 1], out[g + 2], out[g + 3]);
                        } else {
#pragma unroll
                            for (int g = 0; g < OPC; ++g) if (k0 + g < K) o[k0 + g] = out[g];
                        }
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) my[ch * 16 + i] = l[i] * TC_LN2;      // natural-log units for lse_row
                }
            }
            if (CT == 0 && !TC_DBG(4)) {
                for (int k = 0; k < K; ++k) {
                    const float v = lse_row(my + k * C, C);
                    chk += v;
                    if (live) o[k] = v;
                }
            }
            if (live && !(fabsf(chk) < INFINITY)) {
                // out-of-range (or NaN) frame: fp32 CUDA-core recomputation from the standardised parameters
                const float *xg = p.x + frame * D;
                const float *cst32 = p.packed32 + (size_t)D * p.NP2 * 4;
                for (int k = 0; k < K; ++k) {
                    float m = -INFINITY, first = 0.f, sum = 0.f;
                    for (int pass = 0; pass < 2; ++pass) {
                        for (int c = 0; c < C; ++c) {
                            const int kc = k * C + c, pr = kc >> 1, hi = kc & 1;
                            float acc = 0.f;
                            for (int d = 0; d < D; ++d) {
                                const float u = fmaf(xg[d], __ldg(p.packed32 + ((size_t)d * p.NP2 + pr) * 4 + hi),
                                                     __ldg(p.packed32 + ((size_t)d * p.NP2 + pr) * 4 + 2 + hi));
                                acc = fmaf(u, u, acc);
                            }
                            const float lv = fmaf(-0.5f, acc, cst32[kc]);
                            if (pass == 0) { m = fmaxf(m, lv); first = lv; if (WITH_COMP) co[kc] = lv; }
                            else sum += expf(lv - m);
                        }
                        if (pass == 0) { if (C == 1) break; if (isinf(m)) m = 0.f; }
                    }
                    o[k] = (C == 1) ? first : logf(fmaxf(sum, 1e-8f)) + m;
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

// one warp per component (lanes over the dimensions): W1, W2 (fp16 hi/lo, log2 units) and the component's constant
__global__ void __launch_bounds__(32) gmm_pack_tc_kernel(const float *means, const float *log_vars, float scale, const float *logw,
                                                         int KC, int D, int DP, int NP, float *tc) {
    const int kc = blockIdx.x, lane = threadIdx.x;
    if (kc >= NP) return;
    __half *w = reinterpret_cast<__half *>(tc + tc_off_w(DP, NP));
    const size_t wsz = (size_t)NP * DP;
    const double L2E = 1.4426950408889634073599246810019;
    double cst = 0.0;
    bool unsafe = false;
    for (int d = lane; d < DP; d += 32) {
        float w1 = 0.f, w2 = 0.f;
        if (kc < KC && d < D) {
            const double lv = (double)scale * (double)log_vars[(size_t)kc * D + d];
            const double iv = exp(-lv);
            const double dm = (double)means[(size_t)kc * D + d] - (double)tc[tc_off_center() + d];
            w1 = (float)(dm * iv * L2E);
            w2 = (float)(-0.5 * iv * L2E);
            cst += -0.5 * (lv + 1.8378770664093454835606594728112) - 0.5 * dm * dm * iv;
            if (!(fabsf(w1) < 60000.f) || !(fabsf(w2) < 60000.f)) unsafe = true;
        }
        const __half h1 = __float2half_rn(w1), h2 = __float2half_rn(w2);
        const __half l1 = __float2half_rn(w1 - __half2float(h1)), l2 = __float2half_rn(w2 - __half2float(h2));
        const size_t idx = (((size_t)(kc >> 3) * (DP / 8) + (d >> 3)) * 8 + (kc & 7)) * 8 + (d & 7);
        w[0 * wsz + idx] = h1; w[1 * wsz + idx] = l1; w[2 * wsz + idx] = h2; w[3 * wsz + idx] = l2;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cst += __shfl_xor_sync(FULL_MASK, cst, o);
    unsafe = __any_sync(FULL_MASK, unsafe);
    if (lane == 0) {
        if (kc < KC) cst += logw ? (double)logw[kc] : 0.0;
        tc[tc_off_const(DP) + kc] = (kc < KC) ? (float)(cst * L2E) : 0.f;
        if (unsafe || !(fabs(cst) < 2.0e38)) tc[0] = 0.f;
    }
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
    gmm_pack_tc_kernel<<<NP, 32, 0, s>>>(means, log_vars, scale, logw, KC, D, DP, NP, tc);
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

template <bool WITH_COMP, int CT>
static int launch_tc_variant(const CUtensorMap &tmap, const TcParams &p, int grid, size_t smem, cudaStream_t s) {
    // the opt-in to > 48 KB of dynamic shared memory is a per-device function attribute: set once per (device, variant), not per
    // launch (idempotent; concurrent first calls both set the same value)
    static bool done[64];
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !done[dev]) {
        cudaError_t e = cudaFuncSetAttribute(gmm_emission_tc_kernel<WITH_COMP, CT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return set_error(HMMB200_ELAUNCH, "emission_tc smem opt-in: %s", cudaGetErrorString(e));
        if (dev >= 0 && dev < 64) done[dev] = true;
    }
    gmm_emission_tc_kernel<WITH_COMP, CT><<<grid, TC_THREADS, smem, s>>>(tmap, p);
    return check_launch("gmm_emission_tc_kernel");
}

template <bool WITH_COMP>
static int launch_tc_comp(const CUtensorMap &tmap, const TcParams &p, int grid, size_t smem, cudaStream_t s) {
    switch (p.C) {
        case 1: return launch_tc_variant<WITH_COMP, 1>(tmap, p, grid, smem, s);
        case 2: return launch_tc_variant<WITH_COMP, 2>(tmap, p, grid, smem, s);
        case 4: return launch_tc_variant<WITH_COMP, 4>(tmap, p, grid, smem, s);
        case 8: return launch_tc_variant<WITH_COMP, 8>(tmap, p, grid, smem, s);
        default: return launch_tc_variant<WITH_COMP, 0>(tmap, p, grid, smem, s);
    }
}

int launch_emission_tc(const float *x, const float *tc, const float *packed32, int64_t n_frames, int K, int C, int D,
                       float *logb, cudaStream_t s, float *comp) {
    TcParams p = {};
    p.x = x; p.n_frames = n_frames; p.n_tiles = (n_frames + TC_TILE - 1) / TC_TILE;
    p.D = D; p.K = K; p.C = C; p.KC = K * C;
    p.DP = (D + 15) & ~15; p.NP = (p.KC + 15) & ~15;
    p.tc = tc; p.packed32 = packed32; p.NP2 = (p.KC + 1) / 2; p.logb = logb; p.comp = comp;
#ifdef HMMB200_DEBUG_HOOKS
    { const char *e = getenv("HMMB200_TC_DBG"); p.dbg = e ? atoi(e) : 0; }
#endif
    const bool staged = !(C == 1 || C == 2 || C == 4 || C == 8);   // generic C: the epilogue stages its row in shared memory
    const int nbox = (D + TC_BOXW - 1) / TC_BOXW;
    size_t smem = 128 + (size_t)4 * p.NP * p.DP * 2 + (size_t)(p.NP + p.DP) * 4;
    smem = (smem + 15) & ~(size_t)15;
    if (staged) smem += (size_t)TC_TILE * (p.NP + 1) * 4;
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
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = (int)min((int64_t)sms, p.n_tiles);
    if (comp != nullptr) return launch_tc_comp<true>(tmap, p, grid, smem, s);
    return launch_tc_comp<false>(tmap, p, grid, smem, s);
}

}  // namespace hmmb200

#ifdef HMMB200_DEBUG_HOOKS
// debug aid (debug builds only, not part of the public ABI): copies the timing trace of the last traced launch to the host
HMMB200_EXPORT int hmmb200_debug_tc_trace(long long *out) {
    return cudaMemcpyFromSymbol(out, hmmb200::g_tc_trace, sizeof(long long) * 6 * 48 * 2) == cudaSuccess ? 0 : -1;
}
#endif
