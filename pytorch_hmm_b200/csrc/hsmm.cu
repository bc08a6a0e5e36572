// hsmm.cu -- explicit-duration (semi-Markov) recursions for sm_100a.
//
//   hsmm_viterbi_kernel   replaces HSMMLayer._viterbi_decode_single (pytorch_hmm/hsmm.py:245-354, a 5-deep Python loop that
//                         runs at 2.4 frames/s) and SemiMarkovHMM.viterbi_decode (semi_markov.py:455-570).
//   hsmm_forward_kernel   replaces SemiMarkovHMM._unsupervised_forward (semi_markov.py:308-383), in the two-vector form
//                         begin[t][s] / end[t][s] (O(T K (K + Dmax)) instead of O(T K^2 Dmax^2)).
//
// Viterbi keeps the reference's exact floating-point semantics: delta[te][s][d] for a segment of state s, duration d, ending at
// te = t + d - 1 is   ((delta[t-1][s'][d'] + logA[s'][s]) + seg(t,d,s)) + logdur[s][d]   maximised over (s' != s, d') in
// lexicographic order with a strict '>' (first maximum wins), and seg(t,d,s) is summed in ATen's strided-sum order (four
// interleaved partial sums).  One CTA per sequence; thread (s,d) scans its (K-1)*Dmax candidates in the reference's order, so
// scores AND backpointers are bit-identical by construction.  The DP table is a (Dmax+1)-slot ring in shared memory; the
// uint8 backpointers go to a workspace in HBM and are walked by one thread at the end.  The DP ring has Dmax+2 slots.
#include "common.cuh"

#include <stdlib.h>

namespace hmmb200 {

// hsmm_fb.cu: the fp32 two-warp forward-backward (0 launched, 1 shape not covered, < 0 error) and its workspace
size_t hsmm_fb2_workspace_bytes(int B, int T, int K);
int launch_hsmm_fb2(const float *f, const float *segc, const float *logdur, const float *logA, const float *logpi, int B, int T, int K, int Dm,
                    float *gamma, float *total, float *bbegin, float *bend, void *workspace, cudaStream_t s);

struct HsmmVitParams {
    const float *f;        // [B,T,K] per-frame log-emission term
    const float *segc;     // [K] per-segment constant or null (SemiMarkovHMM counts the Gaussian constant once per segment)
    const float *logdur;   // [K,Dm]
    const float *logA;     // [K,K]
    const float *logpi;    // [K] or null (HSMMLayer has no prior on the first segment, hsmm.py:262-268)
    int B, T, K, Dm;
    int sum_order;         // 0: ATen strided row_sum (4 interleaved partials), 1: sequential
    int64_t *states;       // [B,T]
    float *score;          // [B]
    uint8_t *psi_s, *psi_d;  // [B,T,K*Dm] workspace: predecessor (state, duration) of the segment (s,d) STARTING at frame t
    int dbg;               // debug builds only: 1 = stop after the recursion (timing experiments)
};

// sum of d values col[0], col[stride], ... in the order torch.sum uses on a strided fp32 slice
__device__ __forceinline__ float seg_sum(const float *col, int stride, int d, int order) {
    if (order == 1) {
        float a = 0.f;
        for (int i = 0; i < d; ++i) a = __fadd_rn(a, col[(size_t)i * stride]);
        return a;
    }
    float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f;
    const int q = d >> 2;
    for (int i = 0; i < q; ++i) {
        p0 = __fadd_rn(p0, col[(size_t)(4 * i + 0) * stride]);
        p1 = __fadd_rn(p1, col[(size_t)(4 * i + 1) * stride]);
        p2 = __fadd_rn(p2, col[(size_t)(4 * i + 2) * stride]);
        p3 = __fadd_rn(p3, col[(size_t)(4 * i + 3) * stride]);
    }
    for (int i = 4 * q; i < d; ++i) p0 = __fadd_rn(p0, col[(size_t)i * stride]);
    p0 = __fadd_rn(p0, p1);
    p0 = __fadd_rn(p0, p2);
    p0 = __fadd_rn(p0, p3);
    return p0;
}

// Candidate evaluation without the reference's (K-1)*Dmax scan per cell, still bit-identical:
//   tot(s',d') = fl(fl(fl(prev[s'][d'] + a) + oseg) + dsc) is a composition of fp32 roundings, each monotone non-decreasing
//   in prev[s'][d'].  Hence max_{d'} tot(s',d') = tot evaluated at Mx[s'] = max_{d'} prev[s'][d'] (computed once per step and
//   shared by all cells), the best value is max_{s'} of those K-1 numbers, and the reference's winner -- the FIRST (s',d') in
//   lexicographic order with tot == best (strict '>') -- is the first s' attaining best and, within it, the first d' whose own
//   tot rounds to best (checked up to the first arg-max of prev[s'][.], which certainly does).  Work per cell drops from
//   3(K-1)Dmax adds to about 3(K-1) + 3 Dmax/2.
__device__ __forceinline__ float seg_sum_win(const float *win, int K, int Dm, int head, int s, int d, int order) {
    // win: ring of the frames t .. t+Dm-1 (row `head` = frame t); same summation orders as seg_sum
    auto at = [&](int i) { int r = head + i; if (r >= Dm) r -= Dm; return win[r * K + s]; };
    if (order == 1) {
        float a = 0.f;
        for (int i = 0; i < d; ++i) a = __fadd_rn(a, at(i));
        return a;
    }
    float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f;
    const int q = d >> 2;
    for (int i = 0; i < q; ++i) {
        p0 = __fadd_rn(p0, at(4 * i + 0));
        p1 = __fadd_rn(p1, at(4 * i + 1));
        p2 = __fadd_rn(p2, at(4 * i + 2));
        p3 = __fadd_rn(p3, at(4 * i + 3));
    }
    for (int i = 4 * q; i < d; ++i) p0 = __fadd_rn(p0, at(i));
    p0 = __fadd_rn(p0, p1);
    p0 = __fadd_rn(p0, p2);
    p0 = __fadd_rn(p0, p3);
    return p0;
}

__global__ void __launch_bounds__(1024) hsmm_viterbi_kernel(HsmmVitParams p) {
    extern __shared__ __align__(16) float smem_h[];
    const int K = p.K, Dm = p.Dm, T = p.T;
    const int KD = K * Dm, R = Dm + 2;          // slots t-2 (being cleared), t-1 (read) and t .. t+Dm-1 (written) are distinct
    const int NQ = Dm / 4 + 1;                  // prefix-table rows per state (order 0: 4 partial sums per row)
    const int TABW = (p.sum_order == 0) ? NQ * 4 : Dm + 1;
    float *ring = smem_h;                       // [R][K][Dm]   delta for segments ending at te, slot te % R
    float *A_s = ring + (size_t)R * KD;         // [K][K]
    float *dur_s = A_s + K * K;                 // [K][Dm]
    const int WR = Dm + 1;                      // window rows: frames t .. t+Dm-1 are read at step t, frame t+Dm lands in the spare row
    float *win = dur_s + KD;                    // [Dm+1][K]    frames of f (ring, row `head` = frame t)
    float *mx_s = win + (size_t)WR * K;         // [K]          max_d' prev[s'][d']
    int *arg_s = reinterpret_cast<int *>(mx_s + K);        // [K]   first d' (0-based) attaining it
    float *U_s = reinterpret_cast<float *>(arg_s + K);     // [K][K] fl(Mx[s'] + logA[s'][s])
    float *V_s = U_s + K * K;                   // [K]          max_{s' != s} U[s'][s]
    int *varg_s = reinterpret_cast<int *>(V_s + K);        // [K]   first s' attaining it
    float *tab = reinterpret_cast<float *>(varg_s + K);    // [K][TABW] prefix sums of the frame window in the reference's order
    const int b = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const float *f = p.f + (size_t)b * T * K;
    uint8_t *ps = p.psi_s + (size_t)b * T * KD;
    uint8_t *pd = p.psi_d + (size_t)b * T * KD;

    for (int i = tid; i < R * KD; i += blockDim.x) ring[i] = -INFINITY;
    for (int i = tid; i < K * K; i += blockDim.x) A_s[i] = p.logA[i];
    for (int i = tid; i < KD; i += blockDim.x) dur_s[i] = p.logdur[i];
    for (int i = tid; i < WR * K; i += blockDim.x) {                    // frames 0 .. Dm-1 (the spare row is filled at step 0)
        const int fr = i / K, s = i % K;
        win[i] = (fr < Dm && fr < T) ? f[(size_t)fr * K + s] : 0.f;
    }
    __syncthreads();

    // The segment sum seg(t,d,s) in ATen's order is four interleaved running sums plus a tail: the running sums depend on d only through
    // q = d/4, so they are tabulated once per (t, s) and shared by the Dmax cells of the state.  The table of step t+1 only needs the
    // frame window, so it is built DURING step t (two buffers) by threads that have no cell, off the step's critical path.
    auto build_tab = [&](float *tbuf, int first_row) {                  // window rows first_row .. first_row + Dm - 1
        for (int s = blockDim.x - 1 - tid; s < K; s += blockDim.x) {   // the LAST threads of the block
            int r = first_row;
            float *tb = tbuf + s * TABW;
            if (p.sum_order == 0) {
                float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f;
                tb[0] = 0.f; tb[1] = 0.f; tb[2] = 0.f; tb[3] = 0.f;
                for (int q = 1; q < NQ; ++q) {
                    float x[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) { x[j] = win[r * K + s]; if (++r == WR) r = 0; }
                    p0 = __fadd_rn(p0, x[0]); p1 = __fadd_rn(p1, x[1]); p2 = __fadd_rn(p2, x[2]); p3 = __fadd_rn(p3, x[3]);
                    tb[q * 4 + 0] = p0; tb[q * 4 + 1] = p1; tb[q * 4 + 2] = p2; tb[q * 4 + 3] = p3;
                }
            } else {
                float a = 0.f;
                tb[0] = 0.f;
                for (int d = 1; d <= Dm; ++d) { a = __fadd_rn(a, win[r * K + s]); if (++r == WR) r = 0; tb[d] = a; }
            }
        }
    };
    float *tab2 = tab + (size_t)K * TABW;                                // second table buffer
    build_tab(tab, 0);
    __syncthreads();

    int head = 0, slot_prev = R - 1, slot_t = 0;                         // ring rows without integer division on the loop
    for (int t = 0; t < T; ++t) {
        const float *prev = ring + (size_t)slot_prev * KD;              // segments ending at t-1
        const float *tabc = (t & 1) ? tab2 : tab;                        // this step's table; the other buffer receives step t+1's
        // slide the frame window: frame t + Dm goes into the spare row (the one that held frame t - 1), which no cell reads during this
        // step; the two barriers below order it before the table of step t+1 is built from it.
        {
            const int spare = (head == 0) ? WR - 1 : head - 1;
            for (int s = tid; s < K; s += blockDim.x) win[spare * K + s] = (t + Dm < T) ? f[(size_t)(t + Dm) * K + s] : 0.f;
        }
        if (t > 0) {
            // ---- phase A: Mx[s'] = max_d' prev[s'][d'] and its first arg-max (one warp per s') ----
            for (int sp = warp; sp < K; sp += nwarps) {
                float m = -INFINITY;
                int mi = Dm;
                for (int dp = lane; dp < Dm; dp += 32) {
                    const float v = prev[sp * Dm + dp];
                    if (v > m) { m = v; mi = dp; }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const float om = __shfl_xor_sync(FULL_MASK, m, o);
                    const int oi = __shfl_xor_sync(FULL_MASK, mi, o);
                    if (om > m || (om == m && oi < mi)) { m = om; mi = oi; }
                }
                if (lane == 0) { mx_s[sp] = m; arg_s[sp] = (mi < Dm) ? mi : 0; }
            }
            __syncthreads();
            // ---- phase V: V[s] = max_{s' != s} fl(Mx[s'] + logA[s'][s]) with its first arg-max: 32 / VL states per warp, VL lanes each ----
            const int VL = (K <= 4) ? 4 : ((K <= 8) ? 8 : ((K <= 16) ? 16 : 32));
            for (int s0 = warp * (32 / VL); s0 < K; s0 += nwarps * (32 / VL)) {
                const int s = s0 + lane / VL, l = lane % VL;
                float v = -INFINITY;
                int vi = K;
                if (s < K) {
                    for (int sp = l; sp < K; sp += VL) {
                        if (sp == s) continue;
                        const float u = __fadd_rn(mx_s[sp], A_s[sp * K + s]);
                        if (u > v) { v = u; vi = sp; }
                    }
                }
                for (int o = VL / 2; o > 0; o >>= 1) {
                    const float ov = __shfl_xor_sync(FULL_MASK, v, o);
                    const int oi = __shfl_xor_sync(FULL_MASK, vi, o);
                    if (ov > v || (ov == v && oi < vi)) { v = ov; vi = oi; }
                }
                if (s < K && l == 0) { V_s[s] = v; varg_s[s] = (vi < K) ? vi : -1; }
            }
            __syncthreads();
        } else {
            __syncthreads();                                             // (step 0 has no phases A / V: order the window slide before the table build)
        }
        // ---- the table of step t+1 (threads without a cell), beside phase B ----
        if (t + 1 < T) build_tab((t & 1) ? tab : tab2, (head + 1 == WR) ? 0 : head + 1);
        // ---- phase B: one thread per cell (s, d) ----
        for (int pr = tid; pr < KD; pr += blockDim.x) {
            const int s = pr / Dm, d = pr % Dm + 1;
            const int te = t + d - 1;
            if (te < T) {
                float osum;
                {
                    const float *tb = tabc + s * TABW;
                    if (p.sum_order == 0) {
                        const int q = d >> 2;
                        float p0 = tb[q * 4];
                        int r = head + 4 * q;
                        if (r >= WR) r -= WR;
                        for (int i = 4 * q; i < d; ++i) { p0 = __fadd_rn(p0, win[r * K + s]); if (++r == WR) r = 0; }
                        p0 = __fadd_rn(p0, tb[q * 4 + 1]);
                        p0 = __fadd_rn(p0, tb[q * 4 + 2]);
                        osum = __fadd_rn(p0, tb[q * 4 + 3]);
                    } else {
                        osum = tb[d];
                    }
                }
                const float oseg = p.segc ? __fadd_rn(p.segc[s], osum) : osum;
                const float dsc = dur_s[s * Dm + d - 1];
                float best;
                int bs = 0, bd = 1;
                if (t == 0) {
                    best = p.logpi ? __fadd_rn(__fadd_rn(p.logpi[s], oseg), dsc) : __fadd_rn(oseg, dsc);
                } else {
                    // fp32 rounding is monotone: the maximum over all (s', d') is the cell's map applied to V[s]
                    best = __fadd_rn(__fadd_rn(V_s[s], oseg), dsc);
                    if (best > -INFINITY) {
                        // the reference's winner is the FIRST (s', d') in lexicographic order whose own total equals best
                        const int sp_last = varg_s[s];
                        int bsp = sp_last;
                        for (int sp = 0; sp < sp_last; ++sp) {
                            if (sp == s) continue;
                            if (__fadd_rn(__fadd_rn(__fadd_rn(mx_s[sp], A_s[sp * K + s]), oseg), dsc) == best) { bsp = sp; break; }
                        }
                        const float a = A_s[bsp * K + s];
                        const float *pv = prev + bsp * Dm;
                        const int last = arg_s[bsp];
                        int dp = 0;
                        for (; dp < last; ++dp)
                            if (__fadd_rn(__fadd_rn(__fadd_rn(pv[dp], a), oseg), dsc) == best) break;
                        bs = bsp; bd = dp + 1;
                    }
                }
                int slot = slot_t + d - 1;
                if (slot >= R) slot -= R;
                ring[(size_t)slot * KD + s * Dm + d - 1] = best;
                ps[(size_t)t * KD + pr] = (uint8_t)bs;              // indexed by the segment's START frame: a step's
                pd[(size_t)t * KD + pr] = (uint8_t)bd;              // backpointers are KD contiguous bytes (coalesced)
            }
            // slot t-2 has been fully consumed by the previous step: clear it for reuse
            if (t >= 2) {
                int sc = slot_t - 2;
                if (sc < 0) sc += R;
                ring[(size_t)sc * KD + pr] = -INFINITY;
            }
        }
        __syncthreads();
        if (++head == WR) head = 0;
        slot_prev = slot_t;
        if (++slot_t == R) slot_t = 0;
    }
    __syncthreads();

    if (threadIdx.x == 0) {
        const float *last = ring + (size_t)((T - 1) % R) * KD;
        float best = -INFINITY;
        int cs = 0, cd = 1;
        for (int s = 0; s < K; ++s)
            for (int d = 1; d <= Dm; ++d) {
                const float v = last[s * Dm + d - 1];
                if (v > best) { best = v; cs = s; cd = d; }
            }
        if (p.score) p.score[b] = best;
        int64_t *st = p.states + (size_t)b * T;
        int t = T - 1;
        while (t >= 0) {
            int st0 = t - cd + 1;
            if (st0 < 0) st0 = 0;
            for (int u = st0; u <= t; ++u) st[u] = cs;
            if (st0 > 0) {
                const int ns = ps[(size_t)st0 * KD + cs * Dm + cd - 1], nd = pd[(size_t)st0 * KD + cs * Dm + cd - 1];
                t = st0 - 1; cs = ns; cd = nd;
            } else break;
        }
    }
}

// Second form of the same recursion: ONE block barrier per frame instead of four.
//   * Mx[s'] = max_d' delta[t-1][s'][d'] (and its first arg-max) is not reduced from the ring at step t: the cell (s, d) of step t'
//     writes its value into slot te = t' + d - 1 AND folds it into a running maximum of (te, s).  The candidates of a (te, s) arrive in
//     the order d = Dmax .. 1, one per step, from one thread each, so "replace when >=" leaves the maximum with the smallest d.
//   * V[s] = max_{s' != s} fl(Mx[s'] + logA[s'][s]) is recomputed by every cell of the state (K adds) instead of a phase of its own.
//   * the prefix table of step t+1 and the incoming frame of the window are prepared during step t by threads that own no cell.
// Same additions in the same order as hsmm_viterbi_kernel: scores and backpointers are bit-identical.
// KT / DT: compile-time K and max_duration for the shape worth specialising (0 = run-time).  The cell's searches -- V[s], then the FIRST s'
// and the FIRST d' whose own total rounds to the cell's value -- are then straight-line code over registers (all candidates evaluated
// with independent loads and adds, the first match picked by a descending select chain) instead of data-dependent loops of
// load -> add -> add -> add -> compare round trips, which were nine tenths of a step.
__device__ __forceinline__ float2 hs_fadd2(float2 a, float2 b) {
    unsigned long long ra = *reinterpret_cast<unsigned long long *>(&a), rb = *reinterpret_cast<unsigned long long *>(&b), rd;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    return *reinterpret_cast<float2 *>(&rd);
}
constexpr int HS_TB_W = 128;                    // frames of backpointer rows staged per traceback round
template <int KT, int DT>
__global__ void __launch_bounds__(1024) hsmm_viterbi2_kernel(HsmmVitParams p) {
    extern __shared__ __align__(16) float smem_h[];
    const int K = KT > 0 ? KT : p.K, Dm = DT > 0 ? DT : p.Dm, T = p.T;
    const int KD = K * Dm, R = Dm + 2;          // slots t-2 (being cleared), t-1 (read) and t .. t+Dm-1 (written) are distinct
    const int NQ = Dm / 4 + 1;
    const int TABW = (p.sum_order == 0) ? NQ * 4 : Dm + 1;
    const int WR = Dm + 2;                      // window rows: frames t .. t+Dm are live at step t, frame t+Dm+1 lands in the free row
    float *ring = smem_h;                       // [R][K][Dm]   delta for segments ending at te, slot te % R
    float *mr = ring + (size_t)R * KD;          // [R][K]       running max_d' of the slot
    int *ar = reinterpret_cast<int *>(mr + R * K);          // [R][K]  its first arg-max d' (0-based)
    float *A_s = reinterpret_cast<float *>(ar + R * K);     // [K][K]
    float *dur_s = A_s + K * K;                 // [K][Dm]
    float *win = dur_s + KD;                    // [WR][K]      frames of f (ring, row `head` = frame t)
    float *tab = win + (size_t)WR * K;          // [2][K][TABW] prefix sums of the frame window in the reference's order
    const int b = blockIdx.x;
    const int tid = threadIdx.x;
    const float *f = p.f + (size_t)b * T * K;
    // backpointers of this kernel: ONE 16-bit entry per cell (predecessor state | predecessor duration << 8) in the space of the two
    // byte tables -- one store per cell and frame, one load per hop of the traceback
    uint16_t *psd = reinterpret_cast<uint16_t *>(p.psi_s) + (size_t)b * T * KD;

    for (int i = tid; i < R * KD; i += blockDim.x) ring[i] = -INFINITY;
    for (int i = tid; i < R * K; i += blockDim.x) { mr[i] = -INFINITY; ar[i] = 0; }
    for (int i = tid; i < K * K; i += blockDim.x) A_s[i] = p.logA[i];
    for (int i = tid; i < KD; i += blockDim.x) dur_s[i] = p.logdur[i];
    for (int i = tid; i < WR * K; i += blockDim.x) {                    // frames 0 .. Dm (the free row is filled during step 0)
        const int fr = i / K, s = i % K;
        win[i] = (fr <= Dm && fr < T) ? f[(size_t)fr * K + s] : 0.f;
    }
    __syncthreads();
    auto build_tab = [&](float *tbuf, int first_row) {                  // window rows first_row .. first_row + Dm - 1
        for (int s = blockDim.x - 1 - tid; s < K; s += blockDim.x) {   // the LAST threads of the block
            int r = first_row;
            float *tb = tbuf + s * TABW;
            if (p.sum_order == 0) {
                float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f;
                tb[0] = 0.f; tb[1] = 0.f; tb[2] = 0.f; tb[3] = 0.f;
                for (int q = 1; q < NQ; ++q) {
                    float x[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) { x[j] = win[r * K + s]; if (++r == WR) r = 0; }
                    p0 = __fadd_rn(p0, x[0]); p1 = __fadd_rn(p1, x[1]); p2 = __fadd_rn(p2, x[2]); p3 = __fadd_rn(p3, x[3]);
                    tb[q * 4 + 0] = p0; tb[q * 4 + 1] = p1; tb[q * 4 + 2] = p2; tb[q * 4 + 3] = p3;
                }
            } else {
                float a = 0.f;
                tb[0] = 0.f;
                for (int d = 1; d <= Dm; ++d) { a = __fadd_rn(a, win[r * K + s]); if (++r == WR) r = 0; tb[d] = a; }
            }
        }
    };
    float *tab2 = tab + (size_t)K * TABW;
    build_tab(tab, 0);
    // the frame that enters the window at step t (frame t + Dm + 1) is loaded one step ahead into a register, so that the load is in
    // flight across the barrier instead of being waited for before it
    const int fs = blockDim.x - 1 - tid - K;                            // the K threads below the table builders
    float incoming = (fs >= 0 && fs < K && Dm + 1 < T) ? f[(size_t)(Dm + 1) * K + fs] : 0.f;
    __syncthreads();

    int head = 0, slot_prev = R - 1, slot_t = 0;
    // Specialised shape: one cell per thread for the whole sweep, so everything that depends on the cell only -- its state and
    // duration, its column of logA, its duration score, its table offsets, its backpointer addresses -- is set up once, and the
    // searches are straight-line packed adds (add.rn.f32x2 rounds each half like add.rn.f32: same bits).  The frame loop below this
    // block is the general form of the same step; it spent half of its ~430 instructions per cell and frame on that bookkeeping.
    if constexpr (KT > 0 && DT > 0 && DT % 4 == 0) {
        const bool is_cell = tid < KT * DT;
        const int s = is_cell ? tid / DT : 0, dd = is_cell ? tid % DT : 0, d = dd + 1;
        const int q4 = (d >> 2) * 4, rem = d & 3;
        float acol[KT];                                                 // logA[s'][s]; -inf at s' = s (no self transition)
#pragma unroll
        for (int sp = 0; sp < KT; ++sp) acol[sp] = (sp == s) ? -INFINITY : A_s[sp * KT + s];
        const float a_first = A_s[s];                                   // logA[0][s]
        const float dsc = dur_s[s * DT + dd];
        const bool has_seg = p.segc != nullptr;
        const float segv = has_seg ? p.segc[s] : 0.f;
        const float pi0 = p.logpi ? p.logpi[s] : 0.f;
        uint16_t *psp = psd + tid;
        const int tab_off = s * TABW + q4;
        for (int t = 0; t < T; ++t) {
            const float *prev = ring + (size_t)slot_prev * (KT * DT);
            const float *mxp = mr + slot_prev * KT;
            const int *agp = ar + slot_prev * KT;
            const float *tabc = (t & 1) ? tab2 : tab;
            if (fs >= 0 && fs < K) {                                    // window: frame t + Dm + 1 into the row that held frame t - 1
                int row = head + Dm + 1;
                if (row >= WR) row -= WR;
                win[row * K + fs] = incoming;
                incoming = (t + Dm + 2 < T) ? f[(size_t)(t + Dm + 2) * K + fs] : 0.f;
            }
            if (t + 1 < T) build_tab((t & 1) ? tab : tab2, (head + 1 == WR) ? 0 : head + 1);
            if (is_cell) {
                if (t + dd < T) {
                    float osum;
                    if (p.sum_order == 0) {
                        // (all loads first, none under a branch: the adds are the only chain)
                        const float4 tb = *reinterpret_cast<const float4 *>(tabc + tab_off);
                        int r0 = head + q4;
                        if (r0 >= WR) r0 -= WR;
                        const int r1 = (r0 + 1 == WR) ? 0 : r0 + 1, r2 = (r1 + 1 == WR) ? 0 : r1 + 1;
                        const float w0 = win[r0 * KT + s], w1 = win[r1 * KT + s], w2 = win[r2 * KT + s];
                        float p0 = tb.x;
                        p0 = (rem > 0) ? __fadd_rn(p0, w0) : p0;
                        p0 = (rem > 1) ? __fadd_rn(p0, w1) : p0;
                        p0 = (rem > 2) ? __fadd_rn(p0, w2) : p0;
                        p0 = __fadd_rn(p0, tb.y);
                        p0 = __fadd_rn(p0, tb.z);
                        osum = __fadd_rn(p0, tb.w);
                    } else {
                        osum = tabc[s * TABW + d];
                    }
                    const float oseg = has_seg ? __fadd_rn(segv, osum) : osum;
                    const float2 oseg2 = make_float2(oseg, oseg), dsc2 = make_float2(dsc, dsc);
                    float best;
                    int bs = 0, bd = 1;
                    if (t == 0) {
                        best = p.logpi ? __fadd_rn(__fadd_rn(pi0, oseg), dsc) : __fadd_rn(oseg, dsc);
                    } else {
                        static_assert(KT % 2 == 0, "pairs of predecessor states");
                        float2 u2[KT / 2], w2[KT / 2];                  // u = Mx[s'] + logA[s'][s];  w = its total with this cell's terms
#pragma unroll
                        for (int i = 0; i < KT / 2; ++i) {
                            const float2 m2 = *reinterpret_cast<const float2 *>(mxp + 2 * i);
                            u2[i] = hs_fadd2(m2, make_float2(acol[2 * i], acol[2 * i + 1]));
                            w2[i] = hs_fadd2(hs_fadd2(u2[i], oseg2), dsc2);
                        }
                        float v = fmaxf(u2[0].x, u2[0].y);
#pragma unroll
                        for (int i = 1; i < KT / 2; ++i) v = fmaxf(v, fmaxf(u2[i].x, u2[i].y));
                        best = __fadd_rn(__fadd_rn(v, oseg), dsc);
                        if (best > -INFINITY) {
                            int bsp = 0;
                            float a = a_first;
#pragma unroll
                            for (int sp = KT - 1; sp >= 0; --sp) {
                                const float w = (sp & 1) ? w2[sp / 2].y : w2[sp / 2].x;
                                if (w == best) { bsp = sp; a = acol[sp]; }
                            }
                            // The first d' whose total equals best is at or before the first arg-max of the row (that one's total IS
                            // best, and only an earlier, smaller entry can round to the same total): candidates are d' < agp[bsp]
                            // only.  The bound is made uniform over the warp so that the loop does not diverge; when the winners
                            // are short segments it is zero or one group of four instead of all twenty candidates.
                            const float4 *pv4 = reinterpret_cast<const float4 *>(prev + bsp * DT);
                            const float2 a2 = make_float2(a, a);
                            const int last = agp[bsp];
                            const int wmax = (int)__reduce_max_sync(__activemask(), (unsigned)last);
                            int bdp = last;
#pragma unroll 1
                            for (int i4 = 0; 4 * i4 < wmax; ++i4) {
                                const float4 q = pv4[i4];
                                const float2 hi = hs_fadd2(hs_fadd2(hs_fadd2(make_float2(q.z, q.w), a2), oseg2), dsc2);
                                const float2 lo = hs_fadd2(hs_fadd2(hs_fadd2(make_float2(q.x, q.y), a2), oseg2), dsc2);
                                int m = DT;
                                if (hi.y == best) m = 4 * i4 + 3;
                                if (hi.x == best) m = 4 * i4 + 2;
                                if (lo.y == best) m = 4 * i4 + 1;
                                if (lo.x == best) m = 4 * i4;
                                bdp = min(bdp, m);
                            }
                            bs = bsp; bd = bdp + 1;
                        }
                    }
                    int slot = slot_t + dd;
                    if (slot >= R) slot -= R;
                    ring[(size_t)slot * (KT * DT) + tid] = best;
                    // running maximum of (te, s): this step's only candidate for it; ">=" so that the smallest d wins ties
                    if (best >= mr[slot * KT + s]) { mr[slot * KT + s] = best; ar[slot * KT + s] = dd; }
                    psp[(size_t)t * (KT * DT)] = (uint16_t)(bs | (bd << 8));
                }
                if (t >= 2) {                                            // slot t-2 has been fully consumed by the previous step
                    int sc = slot_t - 2;
                    if (sc < 0) sc += R;
                    ring[(size_t)sc * (KT * DT) + tid] = -INFINITY;
                    if (tid < KT) { mr[sc * KT + tid] = -INFINITY; ar[sc * KT + tid] = 0; }
                }
            }
            __syncthreads();
            if (++head == WR) head = 0;
            slot_prev = slot_t;
            if (++slot_t == R) slot_t = 0;
        }
    } else
    for (int t = 0; t < T; ++t) {
        const float *prev = ring + (size_t)slot_prev * KD;              // segments ending at t-1
        const float *mxp = mr + slot_prev * K;
        const int *agp = ar + slot_prev * K;
        const float *tabc = (t & 1) ? tab2 : tab;
        if (fs >= 0 && fs < K) {                                        // window: frame t + Dm + 1 into the row that held frame t - 1
            int row = head + Dm + 1;
            if (row >= WR) row -= WR;
            win[row * K + fs] = incoming;
            incoming = (t + Dm + 2 < T) ? f[(size_t)(t + Dm + 2) * K + fs] : 0.f;
        }
        if (t + 1 < T) build_tab((t & 1) ? tab : tab2, (head + 1 == WR) ? 0 : head + 1);
        for (int pr = tid; pr < KD; pr += blockDim.x) {
            const int s = pr / Dm, d = pr % Dm + 1;
            const int te = t + d - 1;
            if (te < T) {
                float osum;
                {
                    const float *tb = tabc + s * TABW;
                    if (p.sum_order == 0) {
                        const int q = d >> 2;
                        float p0 = tb[q * 4];
                        int r = head + 4 * q;
                        if (r >= WR) r -= WR;
                        for (int i = 4 * q; i < d; ++i) { p0 = __fadd_rn(p0, win[r * K + s]); if (++r == WR) r = 0; }
                        p0 = __fadd_rn(p0, tb[q * 4 + 1]);
                        p0 = __fadd_rn(p0, tb[q * 4 + 2]);
                        osum = __fadd_rn(p0, tb[q * 4 + 3]);
                    } else {
                        osum = tb[d];
                    }
                }
                const float oseg = p.segc ? __fadd_rn(p.segc[s], osum) : osum;
                const float dsc = dur_s[s * Dm + d - 1];
                float best;
                int bs = 0, bd = 1;
                if (t == 0) {
                    best = p.logpi ? __fadd_rn(__fadd_rn(p.logpi[s], oseg), dsc) : __fadd_rn(oseg, dsc);
                } else {
                    // V[s] = max_{s' != s} fl(Mx[s'] + logA[s'][s]); fp32 rounding is monotone, so the maximum over all (s', d') is the
                    // cell's map applied to V[s].  The reference's winner is the FIRST (s', d') in lexicographic order whose own total
                    // equals that value: the first s' whose Mx does, and within it the first d' (Mx[s'] itself certainly does).
                    if constexpr (KT > 0 && DT > 0 && DT % 4 == 0) {
                        float u[KT];
#pragma unroll
                        for (int sp = 0; sp < KT; ++sp) u[sp] = (sp == s) ? -INFINITY : __fadd_rn(mxp[sp], A_s[sp * KT + s]);
                        float v = u[0];
#pragma unroll
                        for (int sp = 1; sp < KT; ++sp) v = fmaxf(v, u[sp]);
                        best = __fadd_rn(__fadd_rn(v, oseg), dsc);
                        if (best > -INFINITY) {
                            int bsp = 0;
#pragma unroll
                            for (int sp = KT - 1; sp >= 0; --sp)
                                if (__fadd_rn(__fadd_rn(u[sp], oseg), dsc) == best) bsp = sp;
                            const float a = A_s[bsp * KT + s];
                            const float4 *pv4 = reinterpret_cast<const float4 *>(prev + bsp * DT);
                            float pvv[DT];
#pragma unroll
                            for (int i4 = 0; i4 < DT / 4; ++i4) {
                                const float4 q4 = pv4[i4];
                                pvv[4 * i4] = q4.x; pvv[4 * i4 + 1] = q4.y; pvv[4 * i4 + 2] = q4.z; pvv[4 * i4 + 3] = q4.w;
                            }
                            int bdp = 0;
#pragma unroll
                            for (int dp = DT - 1; dp >= 0; --dp)
                                if (__fadd_rn(__fadd_rn(__fadd_rn(pvv[dp], a), oseg), dsc) == best) bdp = dp;
                            bs = bsp; bd = bdp + 1;
                        }
                    } else {
                        float v = -INFINITY;
                        int vi = -1;
                        for (int sp = 0; sp < K; ++sp) {
                            if (sp == s) continue;
                            const float u = __fadd_rn(mxp[sp], A_s[sp * K + s]);
                            if (u > v) { v = u; vi = sp; }
                        }
                        best = __fadd_rn(__fadd_rn(v, oseg), dsc);
                        if (best > -INFINITY) {
                            int bsp = vi;
                            for (int sp = 0; sp < vi; ++sp) {
                                if (sp == s) continue;
                                if (__fadd_rn(__fadd_rn(__fadd_rn(mxp[sp], A_s[sp * K + s]), oseg), dsc) == best) { bsp = sp; break; }
                            }
                            const float a = A_s[bsp * K + s];
                            const float *pv = prev + bsp * Dm;
                            const int last = agp[bsp];
                            int dp = 0;
                            for (; dp < last; ++dp)
                                if (__fadd_rn(__fadd_rn(__fadd_rn(pv[dp], a), oseg), dsc) == best) break;
                            bs = bsp; bd = dp + 1;
                        }
                    }
                }
                int slot = slot_t + d - 1;
                if (slot >= R) slot -= R;
                ring[(size_t)slot * KD + s * Dm + d - 1] = best;
                // running maximum of (te, s): this step's only candidate for it; ">=" so that the smallest d wins ties
                if (best >= mr[slot * K + s]) { mr[slot * K + s] = best; ar[slot * K + s] = d - 1; }
                psd[(size_t)t * KD + pr] = (uint16_t)(bs | (bd << 8));
            }
            if (t >= 2) {                                                // slot t-2 has been fully consumed by the previous step
                int sc = slot_t - 2;
                if (sc < 0) sc += R;
                ring[(size_t)sc * KD + pr] = -INFINITY;
                if (pr < K) { mr[sc * K + pr] = -INFINITY; ar[sc * K + pr] = 0; }
            }
        }
        __syncthreads();
        if (++head == WR) head = 0;
        slot_prev = slot_t;
        if (++slot_t == R) slot_t = 0;
    }

    // ---- traceback.  The chain of backpointers is serial (one hop per segment) and every hop used to be a dependent global load
    // (~0.8 us: 1.5 ms for 2000 one-frame segments).  The hops move monotonically back in time, so the backpointer rows are staged
    // HS_TB_W frames at a time in shared memory by the whole block (one contiguous copy) and one thread walks inside the staged
    // rows; the path is collected in shared memory and written out by all threads.
#ifdef HMMB200_DEBUG_HOOKS
    if (p.dbg & 1) return;
#endif
    __shared__ int tb_state[4];                                         // t, state, duration, final-score bits
    const int tb_al = 8 / ((KD % 8 == 0) ? 8 : ((KD % 4 == 0) ? 4 : ((KD % 2 == 0) ? 2 : 1)));   // rows per 16-byte boundary
    const size_t stage_elems = (size_t)((HS_TB_W + 8) * KD + 7) & ~(size_t)7;
    uint16_t *stage = reinterpret_cast<uint16_t *>(smem_h);             // [HS_TB_W + 8][KD]  (the DP tables are dead)
    uint8_t *path = reinterpret_cast<uint8_t *>(stage + stage_elems);   // [T]
    if (tid < 32) {
        const float *last = ring + (size_t)((T - 1) % R) * KD;
        float best = -INFINITY;
        int bi = KD;                                                    // first cell (s-major, then d) attaining the maximum
        for (int i = tid; i < KD; i += 32) { const float v = last[i]; if (v > best) { best = v; bi = i; } }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(FULL_MASK, best, o);
            const int oi = __shfl_xor_sync(FULL_MASK, bi, o);
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        if (tid == 0) {
            if (bi >= KD) bi = 0;                                       // everything -inf: the reference's (state 0, duration 1)
            tb_state[0] = T - 1; tb_state[1] = bi / Dm; tb_state[2] = bi % Dm + 1;
            if (p.score) p.score[b] = best;
        }
    }
    __syncthreads();                                                    // (also: every thread is done with the ring before it is overwritten)
    while (true) {
        const int t = tb_state[0];
        if (t < 0) break;
        // the next backpointer row to be read is the start frame of the current segment; stage rows [lo, hi] = the HS_TB_W rows up to it
        const int cd0 = tb_state[2];
        const int hi = max(t - cd0 + 1, 0);
        int lo = max(hi - HS_TB_W + 1, 0);
        lo -= lo % tb_al;                                               // row lo starts on a 16-byte boundary (vector copies)
        {
            const size_t nbytes = (size_t)(hi - lo + 1) * KD * 2;
            const uint8_t *gs = reinterpret_cast<const uint8_t *>(psd + (size_t)lo * KD);
            uint8_t *ss = reinterpret_cast<uint8_t *>(stage);
            if (((uintptr_t)gs & 15) == 0) {
                // all of a thread's loads before its stores (as load -> store pairs the compiler keeps them in order)
                const size_t n16 = nbytes / 16;
                constexpr int NPT = 4;
                for (size_t i0 = 0; i0 < n16; i0 += (size_t)NPT * blockDim.x) {
                    uint4 vs[NPT];
#pragma unroll
                    for (int j = 0; j < NPT; ++j) {
                        const size_t i = i0 + (size_t)j * blockDim.x + tid;
                        if (i < n16) vs[j] = __ldcs(reinterpret_cast<const uint4 *>(gs) + i);
                    }
#pragma unroll
                    for (int j = 0; j < NPT; ++j) {
                        const size_t i = i0 + (size_t)j * blockDim.x + tid;
                        if (i < n16) reinterpret_cast<uint4 *>(ss)[i] = vs[j];
                    }
                }
                for (size_t i = n16 * 16 + tid; i < nbytes; i += blockDim.x) ss[i] = gs[i];
            } else {
                for (size_t i = tid; i < nbytes; i += blockDim.x) ss[i] = gs[i];
            }
        }
        __syncthreads();
        if (tid == 0) {
            // The walk is a chain of dependent hops (with one-frame segments: T of them), so a hop is kept to a handful of
            // instructions: one 16-bit load, and a path fill that is a plain rolled loop (unrolled by 16 with its remainder ladders it
            // was 60 instructions per hop, a fifth of the kernel's time).
            int tt = t, cs = tb_state[1], cd = cd0;
#ifdef HMMB200_DEBUG_HOOKS
            if (p.dbg & 2) tt = -1;
#endif
            while (tt >= 0) {
                const int st0 = max(tt - cd + 1, 0);
                if (st0 > 0 && st0 < lo) break;                         // its backpointer row is not staged: next round
#pragma unroll 1
                for (int u = tt; u >= st0; --u) path[u] = (uint8_t)cs;
                if (st0 == 0) { tt = -1; break; }
                const unsigned e = stage[(size_t)(st0 - lo) * KD + cs * Dm + cd - 1];
                tt = st0 - 1; cs = (int)(e & 255u); cd = (int)(e >> 8);
            }
            tb_state[0] = tt; tb_state[1] = cs; tb_state[2] = cd;
        }
        __syncthreads();
    }
    {
        int64_t *st = p.states + (size_t)b * T;
#ifdef HMMB200_DEBUG_HOOKS
        if (p.dbg & 4) return;
#endif
        for (int u = tid; u < T; u += blockDim.x) st[u] = (int64_t)path[u];
    }
}

// ----------------------------------------------------------------------------------------------------------
// forward (log-sum-exp semiring), one warp per sequence, lane = state
//   alpha[te][s][d] = begin(te-d+1, s) + segc[s] + sum_{tau=te-d+1..te} f[tau][s] + logdur[s][d]
//   begin(0, s) = logpi[s];  begin(t, s) = LSE_{s' != s}( end[t-1][s'] + logA[s'][s] ),  end[t][s] = LSE_d alpha[t][s][d]
//   total = LSE_s end[T-1][s]                                                     (semi_markov.py:339-378)
// ----------------------------------------------------------------------------------------------------------
struct HsmmFwdParams {
    const float *f, *segc, *logdur, *logA, *logpi;
    int B, T, K, Dm;
    float *alpha;          // [B,T,K,Dm] or null  ("forward_variables")
    float *end_out;        // [B,T,K] or null
    float *total;          // [B]
};

__device__ __forceinline__ float lse2f(float a, float b) {
    const float m = fmaxf(a, b);
    if (!(m > -INFINITY)) return -INFINITY;
    return m + log1pf(expf(fminf(a, b) - m));
}

__global__ void __launch_bounds__(32) hsmm_forward_kernel(HsmmFwdParams p) {
    extern __shared__ __align__(16) float smem_h[];
    const int K = p.K, Dm = p.Dm, T = p.T;
    float *begin_r = smem_h;                    // [Dm][K] ring: begin(t, s), slot t % Dm
    float *f_r = begin_r + Dm * K;              // [Dm][K] ring of the last Dm frames of f
    float *end_s = f_r + Dm * K;                // [K]
    const int b = blockIdx.x, s = threadIdx.x;
    const bool ok = s < K;
    const float *f = p.f + (size_t)b * T * K;
    const float segc = (ok && p.segc) ? p.segc[s] : 0.f;
    for (int t = 0; t < T; ++t) {
        if (ok) {
            f_r[(t % Dm) * K + s] = f[(size_t)t * K + s];
            float bg;
            if (t == 0) bg = p.logpi ? p.logpi[s] : 0.f;
            else {
                bg = -INFINITY;
                for (int sp = 0; sp < K; ++sp) if (sp != s) bg = lse2f(bg, end_s[sp] + p.logA[sp * K + s]);
            }
            begin_r[(t % Dm) * K + s] = bg;
        }
        __syncwarp();
        if (ok) {
            // segments of state s ending at t: duration d starts at t-d+1
            float acc = 0.f, m = -INFINITY, sum = 0.f;
            for (int d = 1; d <= Dm && d <= t + 1; ++d) {
                const int st = t - d + 1;
                acc += f_r[(st % Dm) * K + s];
                const float a = begin_r[(st % Dm) * K + s] + (segc + acc) + p.logdur[s * Dm + d - 1];
                if (p.alpha) p.alpha[(((size_t)b * T + t) * K + s) * Dm + d - 1] = a;
                if (a > m) { sum = sum * expf(m - a) + 1.f; m = a; }
                else if (a > -INFINITY) sum += expf(a - m);
            }
            if (p.alpha) for (int d = t + 2; d <= Dm; ++d) p.alpha[(((size_t)b * T + t) * K + s) * Dm + d - 1] = -INFINITY;
            const float e = (m > -INFINITY) ? m + logf(sum) : -INFINITY;
            if (p.end_out) p.end_out[((size_t)b * T + t) * K + s] = e;
            end_s[s] = e;
        }
        __syncwarp();
    }
    if (s == 0) {
        float tot = -INFINITY;
        for (int k = 0; k < K; ++k) tot = lse2f(tot, end_s[k]);
        p.total[b] = tot;
    }
}

// ----------------------------------------------------------------------------------------------------------
// forward-backward + posteriors (new functionality: the reference has no HSMM backward pass; BASELINE config 4).
// One warp per sequence, lane = state, SCALED PROBABILITY space in double precision:
//   log-space fp32 cannot carry this recursion -- log p(o) is O(-100 T), one fp32 ulp there is ~1e-2, and the
//   posteriors are exponentials of differences of such numbers.  Instead every quantity is a double with the per-frame
//   emission maxima m_t divided out (b~_t(s) = exp(f_t(s) - m_t)) and one shared power-of-two exponent that is pushed
//   out of the (Dmax x K) ring whenever the values drift; no transcendental on the recursion.
//     Bg(t,s)  = pi(s) (t = 0) | sum_{s' != s} E(t-1,s') A(s',s)            a segment of s begins at t
//     E(t,s)   = sum_d Bg(t-d+1,s) dur(s,d) c(s) prod_{tau=t-d+1..t} b~_tau(s)   a segment of s ends at t
//     bend(T-1,s) = 1;  bend(t,s) = sum_{s' != s} A(s,s') bbeg(t+1,s')
//     bbeg(t,s)   = sum_d dur(s,d) c(s) prod_{tau=t..t+d-1} b~_tau(s) bend(t+d-1,s)
//     P(begin at t) = Bg bbeg / p(o),  P(end at t) = E bend / p(o),  gamma_t(s) = sum_{tau<=t} P(begin) - sum_{tau<t} P(end)
// ----------------------------------------------------------------------------------------------------------
struct HsmmFbParams {
    const float *f, *segc, *logdur, *logA, *logpi;
    int B, T, K, Dm;
    float *gamma, *total;                     // [B,T,K], [B]
    float *bbegin_out, *bend_out;             // [B,T,K] log values or null
    double *ws_E, *ws_Bg;                     // [B,T,K] scaled forward values
    double *ws_M;                             // [B,T]   cumulative emission maxima
    int *ws_k;                                // [B,T]   forward exponent at frame t
    float *ws_pend;                           // [B,T,K] end posteriors
};

constexpr int HSF_PF = 8;
__device__ __forceinline__ void hs_cp4(void *dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void hs_cp8(void *dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void hs_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void hs_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(HSF_PF - 1) : "memory"); }

__device__ __forceinline__ double warp_max_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(FULL_MASK, v, o));
    return v;
}

// KT / DT: compile-time K and max_duration for the shapes worth specialising (0 = run-time): the transition and duration
// loops then unroll, and the independent ring updates of a step overlap instead of queueing behind loop control.
template <int KT, int DT>
__global__ void __launch_bounds__(32) hsmm_fb_kernel(HsmmFbParams p) {
    extern __shared__ __align__(16) double smem_d[];
    const int K = KT > 0 ? KT : p.K, Dm = DT > 0 ? DT : p.Dm, T = p.T;
    double *ring = smem_d;                      // [Dm][K]  Bg (forward) / bend (backward), slot t % Dm
    double *bt_r = ring + Dm * K;               // [Dm][K]  b~ of the last Dm frames
    double *durc = bt_r + Dm * K;               // [Dm][K]  dur(s,d) * c(s)
    double *A_s = durc + Dm * K;                // [K][K]   A(s',s) as probabilities
    double *vec = A_s + K * K;                  // [K]      E(t-1,.) / bbeg(t+1,.)
    // prefetch rings (cp.async, HSF_PF steps ahead): the per-step global reads must not sit on the serial chain
    double *pf_E = vec + K;                     // [PF][32]
    double *pf_Bg = pf_E + HSF_PF * 32;         // [PF][32]
    float *pf_f = reinterpret_cast<float *>(pf_Bg + HSF_PF * 32);   // [PF][32]
    int *pf_k = reinterpret_cast<int *>(pf_f + HSF_PF * 32);        // [PF]
    // LPS lanes per state (K = 10 -> 3, K = 12 or 16 -> 2, K <= 8 -> 4, K > 16 -> 1): lane (s, r) takes the durations
    // d = 1 + r, 1 + r + LPS, ... of state s, so the Dmax-long inner loops shrink by LPS and the partial sums meet in shuffles
    const int LPS = (K <= 8) ? 4 : ((K <= 10) ? 3 : ((K <= 16) ? 2 : 1));
    const int b = blockIdx.x, lane_id = threadIdx.x;
    const int s = lane_id / LPS, r = lane_id % LPS;
    const bool ok = s < K;
    const bool lead = ok && r == 0;                          // one lane per state does the stores
    const float *f = p.f + (size_t)b * T * K;
    const size_t base = (size_t)b * T;
    for (int i = lane_id; i < K * K; i += 32) A_s[i] = exp((double)p.logA[i]);
    if (lead) {
        const double c = p.segc ? exp((double)p.segc[s]) : 1.0;
        for (int d = 0; d < Dm; ++d) durc[d * K + s] = exp((double)p.logdur[s * Dm + d]) * c;
        for (int d = 0; d < Dm; ++d) ring[d * K + s] = 0.0;
    }
    __syncwarp();
    // sum of the LPS lanes of a state, identical in all of them (fixed order)
    auto group_sum = [&](double v) {
        const int g0 = ok ? s * LPS : lane_id;
        double t = __shfl_sync(FULL_MASK, v, g0);
        for (int i = 1; i < LPS; ++i) t += __shfl_sync(FULL_MASK, v, ok ? g0 + i : lane_id);
        return ok ? t : 0.0;
    };
    // a shared power-of-two exponent keeps the ring near 1: rescale when the newest values drift by more than 2^24
    auto rescale = [&](double &v, int &kexp) -> double {
        const double mx = warp_max_d(ok ? v : 0.0);
        if (mx > 0.0) {
            const int ex = ilogb(mx);
            if (ex > 24 || ex < -24) {
                const double sc = scalbn(1.0, -ex);
                v *= sc;
                if (ok) for (int d = r; d < Dm; d += LPS) ring[d * K + s] *= sc;
                kexp += ex;
                return sc;
            }
        }
        return 1.0;
    };

    // ---------------- forward ----------------
    int kf = 0;
    double Mc = 0.0;
    int cur = 0;                                            // t % Dm, kept without integer division
    auto pf_fwd = [&](int t) {
        if (lead && t < T) hs_cp4(pf_f + (t % HSF_PF) * 32 + s, f + (size_t)t * K + s);
        hs_commit();
    };
    for (int t = 0; t < HSF_PF - 1; ++t) pf_fwd(t);
    for (int t = 0; t < T; ++t) {
        pf_fwd(t + HSF_PF - 1);
        hs_wait();
        __syncwarp();                                       // the lead lane's copy is read by the state's other lanes
        const float ft = ok ? pf_f[(t % HSF_PF) * 32 + s] : -INFINITY;
        float m = ft;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(FULL_MASK, m, o));
        if (!(m > -INFINITY)) m = 0.f;
        Mc += (double)m;
        double e = 0.0, bg = 0.0;
        if (ok) {
            const double bq = (double)expf(ft - m);
            if (t == 0) bg = p.logpi ? exp((double)p.logpi[s]) : 1.0;
            else {
                double g0 = 0.0, g1 = 0.0;
#pragma unroll
                for (int sp = 0; sp + 1 < K; sp += 2) {
                    if (sp != s) g0 = fma(vec[sp], A_s[sp * K + s], g0);
                    if (sp + 1 != s) g1 = fma(vec[sp + 1], A_s[(sp + 1) * K + s], g1);
                }
                if ((K & 1) && K - 1 != s) g0 = fma(vec[K - 1], A_s[(K - 1) * K + s], g0);
                bg = g0 + g1;
            }
            // the ring holds RUNNING segment products R[st] = Bg(st,s) prod_{tau=st..t} b~_tau(s): one independent multiply
            // per open segment and step instead of a dependent prefix-product chain over the durations
            if (r == 0) ring[cur * K + s] = bg;             // duration 1 belongs to lane r = 0, which reads it back below
            double e0 = 0.0;
            int st = cur - r;                               // slot of duration d = 1 + r: (t - d + 1) % Dm
            if (st < 0) st += Dm;
            // (slots of segments that would begin before frame 0 still hold their initial zeros: no bound on t needed)
#pragma unroll
            for (int d = 1 + r; d <= Dm; d += LPS) {
                const double r0 = ring[st * K + s] * bq;
                ring[st * K + s] = r0;
                e0 = fma(r0, durc[(d - 1) * K + s], e0);
                st -= LPS;
                if (st < 0) st += Dm;
            }
            e = e0;
        }
        e = group_sum(e);
        __syncwarp();                                       // every lane has read E(t-1,.)
        if ((t & 7) == 7) bg *= rescale(e, kf);              // drift is checked every 8th frame: doubles have the headroom
        if (lead) {
            vec[s] = e;
            p.ws_E[(base + t) * K + s] = e;
            p.ws_Bg[(base + t) * K + s] = bg;
        }
        if (lane_id == 0) { p.ws_k[base + t] = kf; p.ws_M[base + t] = Mc; }
        __syncwarp();
        if (++cur == Dm) cur = 0;
    }
    double sumE = lead ? vec[s] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sumE += __shfl_xor_sync(FULL_MASK, sumE, o);
    const int kfT = kf;
    const double Mtot = Mc;
    if (lane_id == 0) p.total[b] = (float)(log(sumE) + 0.69314718055994530942 * (double)kfT + Mtot);
    const double inv = 1.0 / sumE;

    // ---------------- backward + begin/end posteriors ----------------
    if (lead) for (int d = 0; d < Dm; ++d) ring[d * K + s] = 0.0;
    __syncwarp();
    int kb = 0;
    cur = (T - 1) % Dm;
    auto pf_bwd = [&](int t) {                              // frame t, slot t % PF
        if (t >= 0) {
            const int sl = t % HSF_PF;
            if (lead) {
                hs_cp4(pf_f + sl * 32 + s, f + (size_t)t * K + s);
                hs_cp8(pf_E + sl * 32 + s, p.ws_E + (base + t) * K + s);
                hs_cp8(pf_Bg + sl * 32 + s, p.ws_Bg + (base + t) * K + s);
            }
            if (lane_id == 0) hs_cp4(pf_k + sl, p.ws_k + base + t);
        }
        hs_commit();
    };
    __syncwarp();
    for (int i = 0; i < HSF_PF - 1; ++i) pf_bwd(T - 1 - i);
    for (int t = T - 1; t >= 0; --t) {
        pf_bwd(t - (HSF_PF - 1));
        hs_wait();
        __syncwarp();                                       // lane 0's ws_k copy is read by every lane
        const float ft = ok ? pf_f[(t % HSF_PF) * 32 + s] : -INFINITY;
        float m = ft;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(FULL_MASK, m, o));
        if (!(m > -INFINITY)) m = 0.f;
        double bb = 0.0, be = 0.0;
        if (ok) {
            const double bq = (double)expf(ft - m);
            if (t < T - 1) {
                double g0 = 0.0, g1 = 0.0;
#pragma unroll
                for (int sn = 0; sn + 1 < K; sn += 2) {
                    if (sn != s) g0 = fma(A_s[s * K + sn], vec[sn], g0);
                    if (sn + 1 != s) g1 = fma(A_s[s * K + sn + 1], vec[sn + 1], g1);
                }
                if ((K & 1) && K - 1 != s) g0 = fma(A_s[s * K + K - 1], vec[K - 1], g0);
                be = g0 + g1;
            } else {
                be = scalbn(1.0, -kb);
            }
            // ring: Q[en] = bend(en,s) prod_{tau=t..en} b~_tau(s), updated by one multiply per open segment
            if (r == 0) ring[cur * K + s] = be;
            double b0 = 0.0;
            int en = cur + r;                               // slot of duration d = 1 + r: (t + d - 1) % Dm
            if (en >= Dm) en -= Dm;
#pragma unroll
            for (int d = 1 + r; d <= Dm; d += LPS) {          // (segments ending after frame T-1: zeros, as above)
                const double q0 = ring[en * K + s] * bq;
                ring[en * K + s] = q0;
                b0 = fma(q0, durc[(d - 1) * K + s], b0);
                en += LPS;
                if (en >= Dm) en -= Dm;
            }
            bb = b0;
        }
        bb = group_sum(bb);
        __syncwarp();                                       // every lane has read bbeg(t+1,.)
        if ((t & 7) == 0) be *= rescale(bb, kb);
        if (lead) {
            vec[s] = bb;
            const size_t o = (base + t) * K + s;
            const int sl = t % HSF_PF;
            const int ke = pf_k[sl] + kb - kfT;
            p.gamma[o] = (float)scalbn(pf_Bg[sl * 32 + s] * bb * inv, ke);      // P(begins at t); turned into gamma below
            p.ws_pend[o] = (float)scalbn(pf_E[sl * 32 + s] * be * inv, ke);     // P(ends at t)
            if (p.bend_out) p.bend_out[o] = (float)(log(be) + 0.69314718055994530942 * (double)kb + (Mtot - p.ws_M[base + t]));
            if (p.bbegin_out)
                p.bbegin_out[o] = (float)(log(bb) + 0.69314718055994530942 * (double)kb + (Mtot - (t > 0 ? p.ws_M[base + t - 1] : 0.0)));
        }
        __syncwarp();
        cur = (cur == 0) ? Dm - 1 : cur - 1;
    }
    // ---------------- state-occupancy posterior (same thread wrote both arrays: program order suffices) ----------------
    // (loads are batched 8 frames at a time: a load-use-store loop over T would expose the global latency 2T times)
    if (lead) {
        double cum = 0.0;
        for (int t0 = 0; t0 < T; t0 += 8) {
            float gb[8], pe[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int t = min(t0 + i, T - 1);
                const size_t o = (base + t) * K + s;
                gb[i] = p.gamma[o];
                pe[i] = p.ws_pend[o];
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (t0 + i < T) {
                    cum += (double)gb[i];
                    const float g = (float)cum;
                    cum -= (double)pe[i];
                    p.gamma[(base + t0 + i) * K + s] = fminf(fmaxf(g, 0.f), 1.f);
                }
            }
        }
    }
}

}  // namespace hmmb200

using namespace hmmb200;

HMMB200_EXPORT size_t hmmb200_hsmm_viterbi_workspace_bytes(int B, int T, int K, int Dm) {
    if (B <= 0 || T <= 0 || K <= 0 || Dm <= 0) return 0;
    return 2 * (size_t)B * T * K * Dm;
}

HMMB200_EXPORT int hmmb200_hsmm_viterbi_f32(const float *frame_logp, const float *seg_const, const float *log_dur,
                                            const float *log_trans, const float *log_init, int B, int T, int K, int Dm,
                                            int sum_order, int64_t *states, float *score,
                                            void *workspace, size_t workspace_bytes, void *stream) {
    if (B < 0 || T < 0 || K <= 0 || Dm <= 0) return set_error(HMMB200_EINVAL, "hsmm_viterbi: bad shape");
    if (B == 0 || T == 0) return HMMB200_OK;
    if (!frame_logp || !log_dur || !log_trans || !states) return set_error(HMMB200_EINVAL, "hsmm_viterbi: null argument");
    if (K > 255 || Dm > 255) return set_error(HMMB200_EUNSUPPORTED, "hsmm_viterbi: K and max_duration must be <= 255");
    const size_t need = hmmb200_hsmm_viterbi_workspace_bytes(B, T, K, Dm);
    if (!workspace || workspace_bytes < need) return set_error(HMMB200_EWORKSPACE, "hsmm_viterbi: workspace %zu < %zu", workspace_bytes, need);
    if (int rc = require_sm100()) return rc;
    const size_t smem = ((size_t)(Dm + 2) * K * Dm + 2 * (size_t)K * K + (size_t)2 * K * Dm + 5 * (size_t)K + 2 * (size_t)K * (Dm + 8)) * sizeof(float);
    if (smem > 200 * 1024) return set_error(HMMB200_EUNSUPPORTED, "hsmm_viterbi: K=%d, max_duration=%d need %zu bytes of shared memory", K, Dm, smem);
    cudaError_t e = cudaFuncSetAttribute(hsmm_viterbi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_error(HMMB200_ELAUNCH, "hsmm_viterbi smem opt-in: %s", cudaGetErrorString(e));
    HsmmVitParams p;
    p.f = frame_logp; p.segc = seg_const; p.logdur = log_dur; p.logA = log_trans; p.logpi = log_init;
    p.B = B; p.T = T; p.K = K; p.Dm = Dm; p.sum_order = sum_order; p.states = states; p.score = score;
    p.psi_s = (uint8_t *)workspace; p.psi_d = (uint8_t *)workspace + (size_t)B * T * K * Dm;
    p.dbg = 0;
#ifdef HMMB200_DEBUG_HOOKS
    if (const char *e = getenv("HMMB200_HSMM_VIT_DBG")) p.dbg = atoi(e);
#endif
    int threads = ((K * Dm + 31) / 32) * 32;
    if (threads > 1024) threads = 1024;
    // one-barrier form: needs 2 K threads beside the cells (table builders, window feeders) and its own shared-memory layout
    const int R = Dm + 2;
    size_t smem2 = ((size_t)R * K * Dm + 2 * (size_t)R * K + (size_t)K * K + (size_t)K * Dm + (size_t)(Dm + 2) * K + 2 * (size_t)K * (Dm + 8)) * sizeof(float);
    const size_t smem_tb = 2 * ((size_t)(HS_TB_W + 8) * K * Dm + 8) + (size_t)T + 64;         // traceback staging (reuses the DP tables' space)
    if (smem_tb > smem2) smem2 = smem_tb;
    // (the helpers get warps of their own: sharing a warp with cells, that warp ran the cell code AND the table builder one after the
    // other and the frame barrier waited for it)
    int threads2 = ((K * Dm + 31) / 32) * 32 + ((2 * K + 31) / 32) * 32;
    bool v1 = threads2 > 1024 || smem2 > 200 * 1024;
#ifdef HMMB200_DEBUG_HOOKS
    if (getenv("HMMB200_HSMM_VIT_V1")) v1 = true;
#endif
    if (!v1) {
        const bool spec = (K == 10 && Dm == 20);                // BASELINE config 4 (the reference factory's HSMM defaults)
        e = spec ? cudaFuncSetAttribute(hsmm_viterbi2_kernel<10, 20>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2)
                 : cudaFuncSetAttribute(hsmm_viterbi2_kernel<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
        if (e != cudaSuccess) return set_error(HMMB200_ELAUNCH, "hsmm_viterbi smem opt-in: %s", cudaGetErrorString(e));
        if (spec) hsmm_viterbi2_kernel<10, 20><<<B, threads2, smem2, (cudaStream_t)stream>>>(p);
        else hsmm_viterbi2_kernel<0, 0><<<B, threads2, smem2, (cudaStream_t)stream>>>(p);
        return check_launch("hsmm_viterbi2_kernel");
    }
    hsmm_viterbi_kernel<<<B, threads, smem, (cudaStream_t)stream>>>(p);
    return check_launch("hsmm_viterbi_kernel");
}

HMMB200_EXPORT int hmmb200_hsmm_forward_f32(const float *frame_logp, const float *seg_const, const float *log_dur,
                                            const float *log_trans, const float *log_init, int B, int T, int K, int Dm,
                                            float *alpha, float *end_scores, float *total, void *stream) {
    if (B < 0 || T < 0 || K <= 0 || Dm <= 0) return set_error(HMMB200_EINVAL, "hsmm_forward: bad shape");
    if (B == 0 || T == 0) return HMMB200_OK;
    if (!frame_logp || !log_dur || !log_trans || !total) return set_error(HMMB200_EINVAL, "hsmm_forward: null argument");
    if (K > 32) return set_error(HMMB200_EUNSUPPORTED, "hsmm_forward: K <= 32 (got %d)", K);
    if (int rc = require_sm100()) return rc;
    const size_t smem = ((size_t)2 * Dm * K + K) * sizeof(float);
    if (smem > 200 * 1024) return set_error(HMMB200_EUNSUPPORTED, "hsmm_forward: max_duration too large");
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(hsmm_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return set_error(HMMB200_ELAUNCH, "hsmm_forward smem opt-in: %s", cudaGetErrorString(e));
    }
    HsmmFwdParams p;
    p.f = frame_logp; p.segc = seg_const; p.logdur = log_dur; p.logA = log_trans; p.logpi = log_init;
    p.B = B; p.T = T; p.K = K; p.Dm = Dm; p.alpha = alpha; p.end_out = end_scores; p.total = total;
    hsmm_forward_kernel<<<B, 32, smem, (cudaStream_t)stream>>>(p);
    return check_launch("hsmm_forward_kernel");
}

HMMB200_EXPORT size_t hmmb200_hsmm_fb_workspace_bytes(int B, int T, int K) {
    if (B <= 0 || T <= 0 || K <= 0) return 0;
    const size_t n = (size_t)B * T;
    const size_t v1 = n * K * (2 * sizeof(double) + sizeof(float)) + n * (sizeof(double) + sizeof(int)) + 64;
    const size_t v2 = hsmm_fb2_workspace_bytes(B, T, K);
    return v1 > v2 ? v1 : v2;
}

HMMB200_EXPORT int hmmb200_hsmm_forward_backward_f32(const float *frame_logp, const float *seg_const, const float *log_dur,
                                                     const float *log_trans, const float *log_init, int B, int T, int K, int Dm,
                                                     float *gamma, float *total, float *beta_begin, float *beta_end,
                                                     void *workspace, size_t workspace_bytes, void *stream) {
    if (B < 0 || T < 0 || K <= 0 || Dm <= 0) return set_error(HMMB200_EINVAL, "hsmm_forward_backward: bad shape");
    if (B == 0 || T == 0) return HMMB200_OK;
    if (!frame_logp || !log_dur || !log_trans || !gamma || !total) return set_error(HMMB200_EINVAL, "hsmm_forward_backward: null argument");
    if (K > 32) return set_error(HMMB200_EUNSUPPORTED, "hsmm_forward_backward: K <= 32 (got %d)", K);
    const size_t need = hmmb200_hsmm_fb_workspace_bytes(B, T, K);
    if (!workspace || workspace_bytes < need) return set_error(HMMB200_EWORKSPACE, "hsmm_forward_backward: workspace %zu < %zu", workspace_bytes, need);
    if (int rc = require_sm100()) return rc;
    {
        bool v1 = false;
#ifdef HMMB200_DEBUG_HOOKS
        v1 = getenv("HMMB200_HSMM_FB_V1") != nullptr;            // A/B timing against the double-precision one-warp kernel
#endif
        if (!v1) {
            const int rc = launch_hsmm_fb2(frame_logp, seg_const, log_dur, log_trans, log_init, B, T, K, Dm, gamma, total, beta_begin, beta_end,
                                           workspace, (cudaStream_t)stream);
            if (rc <= 0) return rc;                              // launched (0) or failed (< 0); 1 = shape for the general kernel below
        }
    }
    const size_t smem = ((size_t)3 * Dm * K + (size_t)K * K + K) * sizeof(double) + (size_t)HSF_PF * 32 * (8 + 8 + 4) + HSF_PF * 4 + 64;
    if (smem > 200 * 1024) return set_error(HMMB200_EUNSUPPORTED, "hsmm_forward_backward: max_duration too large");
    const bool spec = (K == 10 && Dm == 20);                // BASELINE config 4 (HSMMLayer defaults of the reference's factory)
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(hsmm_fb_kernel<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return set_error(HMMB200_ELAUNCH, "hsmm_forward_backward smem opt-in: %s", cudaGetErrorString(e));
    }
    const size_t n = (size_t)B * T;
    uint8_t *w = (uint8_t *)workspace;
    HsmmFbParams p;
    p.f = frame_logp; p.segc = seg_const; p.logdur = log_dur; p.logA = log_trans; p.logpi = log_init;
    p.B = B; p.T = T; p.K = K; p.Dm = Dm; p.gamma = gamma; p.total = total; p.bbegin_out = beta_begin; p.bend_out = beta_end;
    p.ws_E = (double *)w;  w += n * K * sizeof(double);
    p.ws_Bg = (double *)w; w += n * K * sizeof(double);
    p.ws_M = (double *)w;  w += n * sizeof(double);
    p.ws_pend = (float *)w; w += n * K * sizeof(float);
    p.ws_k = (int *)w;
    if (spec) hsmm_fb_kernel<10, 20><<<B, 32, smem, (cudaStream_t)stream>>>(p);
    else hsmm_fb_kernel<0, 0><<<B, 32, smem, (cudaStream_t)stream>>>(p);
    return check_launch("hsmm_fb_kernel");
}
