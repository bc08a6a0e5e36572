// bw_tc.cu -- Baum-Welch Gaussian-mixture statistics on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
//   occ[kc] = sum_f w[f][kc]      sx[kc][d] = sum_f w[f][kc] x[f][d]      sxx[kc][d] = sum_f w[f][kc] x[f][d]^2
// with w = gamma_t(k) * responsibility(c | k) is ONE GEMM with the frames as the reduction dimension:
//   D[128, N] += A[128, 64 frames] * B[N, 64 frames]^T       per 64-frame tile, N = 2 D + 16
//   A rows (TMEM lanes)  0 .. KC-1  : bf16(w)           rows 64 .. 64+KC-1 : bf16(w - bf16(w))         (hi / lo halves of the weights)
//   B rows (shared mem)  0 .. D-1   : x    D .. 2D-1 : x^2    row 2D : ones (the occupancies)           issued twice: B_hi, B_lo
// so the accumulator rows kc and 64 + kc hold (w_hi + w_lo)(x_hi + x_lo) between them: all four partial products, i.e. operands
// that carry 16 significant bits each (relative error ~2^-17 per product, far inside the 1e-4 contract of the statistics).  bf16
// rather than fp16 because x^2 has no bound.  The accumulator stays in TMEM over ALL tiles of the CTA and is read once.
//
// One persistent CTA per SM, 512 threads, tile = 64 frames:
//   * inputs (x, comp, gamma, log b rows of the tile: four contiguous byte ranges) arrive by cp.async.bulk in a 3-stage ring;
//   * every thread: w pass (gamma * exp(comp - log b) -> shared memory), then the operand pass: A pairs -> tcgen05.st (thread = TMEM lane
//     = weight row), B core-matrix rows -> st.shared.v4 (thread = (column, 8 frames): transposes x on the way);
//   * one thread issues 2 x 4 tcgen05.mma (kind::f16, bf16 inputs, A from TMEM, B K-major no-swizzle) and commits to an mbarrier that
//     the next tile's operand pass waits on (the w pass of tile i+1 overlaps the MMAs of tile i).
// mma.sync (the previous version of this kernel, bw.cu) runs the same product at ~240 MAC/clk/SM on this part; CUDA-core FMAs at 57 %
// of the fp32 pipe.
#include "common.cuh"

#include <cuda_bf16.h>

namespace hmmb200 {

constexpr int ST_F = 64;            // frames per tile (the K extent of a tile's MMAs: 4 instructions of K = 16)
constexpr int ST_STAGES = 3;
constexpr int ST_THREADS = 512;
constexpr int ST_ACOLS = ST_F / 2;  // TMEM columns of the A operand (two bf16 per column)

__device__ __forceinline__ uint32_t st_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void st_mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(st_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void st_mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(st_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void st_mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "STWAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra STDONE_%=;\n\t"
        "bra STWAIT_%=;\n\t"
        "STDONE_%=:\n\t"
        "}" ::"r"(st_smem_u32(bar)), "r"(parity), "r"(1000000u) : "memory");
}
__device__ __forceinline__ void st_bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(st_smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(st_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void st_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void st_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void st_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(st_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void st_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void st_st4(uint32_t taddr, const uint32_t (&v)[4]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]) : "memory");
}
__device__ __forceinline__ void st_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void st_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void st_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// K-major, no-swizzle shared-memory matrix descriptor (as emission_tc.cu): core matrix = 8 rows x 16 bytes,
//   LBO = byte distance between the two 16-byte K chunks of a K = 16 slice, SBO = byte distance between 8-row groups
__device__ __forceinline__ uint64_t st_smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr >> 4) & 0x3fffu) | ((uint64_t)((lbo >> 4) & 0x3fffu) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
// (v0, v1) -> packed bf16 pair of the rounded values (v0 in the low half) and the packed pair of the exact residuals' roundings
__device__ __forceinline__ void st_split2(float v0, float v1, uint32_t &hi, uint32_t &lo) {
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(v1), "f"(v0));
    const float r0 = v0 - __uint_as_float(hi << 16), r1 = v1 - __uint_as_float(hi & 0xffff0000u);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(r1), "f"(r0));
}

struct StatsTcParams {
    const float *x, *comp, *logb, *gamma;
    int64_t n;               // frames (a multiple of ST_F)
    int K, C, D;
    double *occ, *sx, *sxx;
};

__global__ void __launch_bounds__(ST_THREADS, 1) bw_gmm_stats_tc_kernel(StatsTcParams p) {
    extern __shared__ __align__(128) uint8_t smem_st[];
    const int K = p.K, C = p.C, D = p.D, KC = K * C;
    const int N = 2 * D + 16;                                   // B rows: x, x^2, then a 16-row block whose first row is ones
    const int WP = KC + 1;                                      // w row pitch (floats): the A pass reads a column per thread
    const int stage_floats = ST_F * (D + KC + 2 * K);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // ---- carve shared memory ----
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_st);     // [ST_STAGES]
    uint64_t *mma_done = full + ST_STAGES;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(mma_done + 1);
    size_t off = 128;
    const uint32_t b_bytes = (uint32_t)(N / 8) * (ST_F / 8) * 128;   // one B operand: [N/8][F/8] core matrices of 8 x 8 bf16
    uint8_t *b_hi = smem_st + off;                              off += b_bytes;
    uint8_t *b_lo = smem_st + off;                              off += b_bytes;
    float *w_s = reinterpret_cast<float *>(smem_st + off);      off += (size_t)ST_F * WP * sizeof(float);
    off = (off + 15) & ~(size_t)15;
    float *ring = reinterpret_cast<float *>(smem_st + off);

    const int64_t n_tiles = p.n / ST_F;
    const int64_t my_tiles = (n_tiles > blockIdx.x) ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    auto issue = [&](int64_t it) {
        const int64_t base = (blockIdx.x + it * gridDim.x) * ST_F;
        float *st = ring + (size_t)(it % ST_STAGES) * stage_floats;
        uint64_t *bar = full + (it % ST_STAGES);
        const uint32_t bx = ST_F * D * 4, bc = ST_F * KC * 4, bk = ST_F * K * 4;
        st_mbar_expect_tx(bar, bx + bc + 2 * bk);
        st_bulk_g2s(st, p.x + base * D, bx, bar);
        st_bulk_g2s(st + ST_F * D, p.comp + base * KC, bc, bar);
        st_bulk_g2s(st + ST_F * (D + KC), p.gamma + base * K, bk, bar);
        st_bulk_g2s(st + ST_F * (D + KC + K), p.logb + base * K, bk, bar);
    };
    if (tid == 0) {
        for (int i = 0; i < ST_STAGES; ++i) st_mbar_init(full + i, 1);
        st_mbar_init(mma_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int64_t it = 0; it < min((int64_t)ST_STAGES, my_tiles); ++it) issue(it);
    }
    if (warp == 0) {                                            // TMEM: 256 columns (A: 32, accumulator: N <= 208), allocated and freed by warp 0
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(st_smem_u32(tmem_slot)), "r"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // the constant block of B: rows 2D .. 2D+15, row 2D = 1.0 in B_hi (bf16 0x3f80), everything else 0; the lo copy is all zero
    for (int e = tid; e < 2 * (ST_F / 8) * 8; e += ST_THREADS) {       // 2 row groups x F/8 k-chunks x 8 rows, 16 bytes each
        const int g = e / ((ST_F / 8) * 8), kc8 = (e / 8) % (ST_F / 8), r = e % 8;
        const uint32_t o = (uint32_t)((2 * D / 8 + g) * (ST_F / 8) + kc8) * 128 + r * 16;
        const uint32_t v = (g == 0 && r == 0) ? 0x3f803f80u : 0u;
        *reinterpret_cast<uint4 *>(b_hi + o) = make_uint4(v, v, v, v);
        *reinterpret_cast<uint4 *>(b_lo + o) = make_uint4(0u, 0u, 0u, 0u);
    }
    st_fence_before();
    __syncthreads();
    st_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t a_col0 = 0, d_col0 = ST_ACOLS;
    // this thread's TMEM lane and the weight row it feeds: lanes 0 .. 63 take bf16(w), lanes 64 .. 127 the residual
    const int q = warp & 3, tl = q * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const int a_row = (tl < 64) ? tl : tl - 64;
    const bool a_lo = tl >= 64, a_live = a_row < KC;
    const int col_part = warp >> 2;                             // the four warps of a lane quarter take 8 of the 32 A columns each
    const int step_f = ST_THREADS / KC, step_kc = ST_THREADS % KC;
    const float inv_c = 1.f / (float)C;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // bf16 x bf16 -> f32, M = 128, K-major A and B
    const uint32_t lbo = 128, sbo = (uint32_t)(ST_F / 8) * 128;

    for (int64_t it = 0; it < my_tiles; ++it) {
        const float *st = ring + (size_t)(it % ST_STAGES) * stage_floats;
        st_mbar_wait(full + (it % ST_STAGES), (uint32_t)((it / ST_STAGES) & 1));
        const float *x_s = st, *c_s = st + ST_F * D, *g_s = st + ST_F * (D + KC), *l_s = st + ST_F * (D + KC + K);
        // ---- w pass ----
        {
            int f = tid / KC, kc = tid % KC;
            for (int e = tid; e < ST_F * KC; e += ST_THREADS) {
                const int k = __float2int_rz(((float)kc + 0.5f) * inv_c);
                w_s[f * WP + kc] = g_s[f * K + k] * __expf(c_s[e] - l_s[f * K + k]);   // gamma_t(k) * responsibility(c | k)
                f += step_f; kc += step_kc;
                if (kc >= KC) { kc -= KC; ++f; }
            }
        }
        __syncthreads();
        // ---- operand pass: the previous tile's MMAs must have retired before A and B are overwritten ----
        if (it > 0) st_mbar_wait(mma_done, (uint32_t)((it - 1) & 1));
        st_fence_after();
        {   // A: this lane's weight row, frames 16 * col_part .. + 15 as 8 packed columns
            const uint32_t a_addr = tmem_base + lane_addr + a_col0 + col_part * 8;
            const float *wcol = w_s + (col_part * 16) * WP + a_row;
#pragma unroll
            for (int c4 = 0; c4 < 2; ++c4) {
                uint32_t v[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int f = c4 * 8 + 2 * e;
                    const float w0 = a_live ? wcol[f * WP] : 0.f, w1 = a_live ? wcol[(f + 1) * WP] : 0.f;
                    uint32_t hi, lo;
                    st_split2(w0, w1, hi, lo);
                    v[e] = a_lo ? lo : hi;
                }
                st_st4(a_addr + c4 * 4, v);
            }
        }
        // B: task = (dim d, group of 8 frames) -> the 16-byte core-matrix rows of x (row d) and x^2 (row D + d) in B_hi and B_lo
        {
            int d = tid % D, kg = tid / D;                      // (walked without integer divisions in the loop)
            const int step_d = ST_THREADS % D, step_kg = ST_THREADS / D;
            while (kg < ST_F / 8) {
                const float *src = x_s + (kg * 8) * D + d;
                uint32_t h[4], l[4], h2[4], l2[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float v0 = src[(2 * e) * D], v1 = src[(2 * e + 1) * D];
                    st_split2(v0, v1, h[e], l[e]);
                    st_split2(v0 * v0, v1 * v1, h2[e], l2[e]);
                }
                const uint32_t o = (uint32_t)((d >> 3) * (ST_F / 8) + kg) * 128 + (d & 7) * 16;
                const uint32_t o2 = o + (uint32_t)(D >> 3) * (ST_F / 8) * 128;
                *reinterpret_cast<uint4 *>(b_hi + o) = make_uint4(h[0], h[1], h[2], h[3]);
                *reinterpret_cast<uint4 *>(b_lo + o) = make_uint4(l[0], l[1], l[2], l[3]);
                *reinterpret_cast<uint4 *>(b_hi + o2) = make_uint4(h2[0], h2[1], h2[2], h2[3]);
                *reinterpret_cast<uint4 *>(b_lo + o2) = make_uint4(l2[0], l2[1], l2[2], l2[3]);
                d += step_d; kg += step_kg;
                if (d >= D) { d -= D; ++kg; }
            }
        }
        st_wait_st();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // B written with generic stores, read by the MMA (async proxy)
        st_fence_before();
        __syncthreads();
        if (tid == 0) {
            st_fence_after();
            const uint32_t d_addr = tmem_base + d_col0, a_addr = tmem_base + a_col0;
            const uint64_t dh = st_smem_desc(st_smem_u32(b_hi), lbo, sbo), dl = st_smem_desc(st_smem_u32(b_lo), lbo, sbo);
#pragma unroll
            for (int kk = 0; kk < ST_F / 16; ++kk) {
                st_mma_ts(d_addr, a_addr + kk * 8, dh + (uint64_t)(kk * 16), idesc, (it > 0 || kk > 0) ? 1u : 0u);
                st_mma_ts(d_addr, a_addr + kk * 8, dl + (uint64_t)(kk * 16), idesc, 1u);
            }
            st_commit(mma_done);
            // the tile's inputs have been consumed (by generic loads, all before the barrier above): refill the stage
            if (it + ST_STAGES < my_tiles) issue(it + ST_STAGES);
        }
    }
    // ---- the accumulator: rows kc (hi weights) and 64 + kc (lo weights), columns [x | x^2 | ones] ----
    if (my_tiles > 0) {
        st_mbar_wait(mma_done, (uint32_t)((my_tiles - 1) & 1));
        st_fence_after();
        float *fold = ring;                                     // [128][N] floats
        if (warp < 4) {
            const uint32_t d_addr = tmem_base + lane_addr + d_col0;
            for (int ch = 0; ch < N / 16; ++ch) {
                uint32_t v[16];
                st_ld16(d_addr + ch * 16, v);
                st_wait_ld();
#pragma unroll
                for (int i = 0; i < 16; ++i) fold[(size_t)tl * N + ch * 16 + i] = __uint_as_float(v[i]);
            }
        }
        st_fence_before();
        __syncthreads();
        for (int e = tid; e < KC * (2 * D + 1); e += ST_THREADS) {
            const int kc = e / (2 * D + 1), c = e % (2 * D + 1);
            const double v = (double)fold[(size_t)kc * N + c] + (double)fold[(size_t)(64 + kc) * N + c];
            if (c < D) atomicAdd(p.sx + (size_t)kc * D + c, v);
            else if (c < 2 * D) atomicAdd(p.sxx + (size_t)kc * D + (c - D), v);
            else atomicAdd(p.occ + kc, v);
        }
    }
    __syncthreads();
    if (warp == 0) {
        st_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256) : "memory");
    }
}

// 0 launched, 1 shape outside this kernel (caller uses the other forms), <0 error.  n_full = frames, a multiple of 64.
int launch_gmm_stats_tc(const float *x, const float *comp, const float *logb, const float *gamma, int64_t n_full, int K, int C, int D,
                        double *occ, double *sx, double *sxx, cudaStream_t s) {
    const int KC = K * C, N = 2 * D + 16;
    if (KC > 64 || D % 8 != 0 || N > 208 || n_full < ST_F || n_full % ST_F != 0) return 1;
    const size_t b_bytes = (size_t)(N / 8) * (ST_F / 8) * 128;
    const size_t stage = (size_t)ST_F * (D + KC + 2 * K) * sizeof(float);
    size_t smem = 128 + 2 * b_bytes + (size_t)ST_F * (KC + 1) * sizeof(float) + 16 + ST_STAGES * stage;
    if (smem > 200 * 1024 || ST_STAGES * stage < (size_t)128 * N * sizeof(float)) return 1;
    static bool attr_done[64];
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (dev < 0 || dev >= 64 || !attr_done[dev]) {
        cudaError_t e = cudaFuncSetAttribute(bw_gmm_stats_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return set_error(HMMB200_ELAUNCH, "gmm_stats (tcgen05) smem opt-in: %s", cudaGetErrorString(e));
        if (dev >= 0 && dev < 64) attr_done[dev] = true;
    }
    StatsTcParams p;
    p.x = x; p.comp = comp; p.logb = logb; p.gamma = gamma; p.n = n_full; p.K = K; p.C = C; p.D = D; p.occ = occ; p.sx = sx; p.sxx = sxx;
    bw_gmm_stats_tc_kernel<<<(unsigned)min((int64_t)sms, n_full / ST_F), ST_THREADS, smem, s>>>(p);
    return check_launch("bw_gmm_stats_tc_kernel");
}

}  // namespace hmmb200
