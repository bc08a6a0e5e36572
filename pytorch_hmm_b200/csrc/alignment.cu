// alignment.cu -- CTC forward / backward trellises and dynamic time warping for sm_100a.
//
//   ctc_trellis_kernel   replaces ctc_forward_algorithm / ctc_backward_algorithm (pytorch_hmm/alignment/ctc.py:32-121, :124-199):
//                        the same log-semiring recursion as the HMM forward pass over a banded (3-diagonal) transition structure
//                        on the blank-expanded target.  One CTA per utterance, one thread per expanded position, the previous row
//                        in shared memory (ping-pong), the gathered emission column of the NEXT frame loaded before the barrier so
//                        that the global gather is off the time chain.  The reference runs three nested Python loops with a
//                        torch.logsumexp call per cell.
//   dtw_kernel           replaces compute_dtw_path (pytorch_hmm/alignment/dtw.py:47-153): min-plus wavefront over the anti-diagonals
//                        of the [N,M] distance matrix, the last two diagonals in shared memory, one CTA per pair; the backpointer of
//                        every cell (the reference's tie rule: diagonal, then up, then left) is recorded on the way so that the
//                        traceback is a walk over bytes instead of a re-evaluation of three costs per step.  Costs are sums and
//                        minima of fp32 numbers in the reference's order: bit-identical.
#include "common.cuh"

#include <math.h>

namespace hmmb200 {

struct CtcParams {
    const float *lp;            // [T,B,C]
    const int64_t *targets;     // [B,L]
    const int64_t *in_len, *tg_len;   // [B]
    int blank, T, B, C, L;
    float *table;               // [B,T,2L+1] log alpha / log beta, or null
    float *loglik;              // [B] (forward only) or null
};

// torch.logsumexp of up to three candidates, all > -inf, in the order given: max, sum of exp(x - max), log, + max
__device__ __forceinline__ float lse_cands(const float *c, int n) {
    float m = c[0];
    for (int i = 1; i < n; ++i) m = fmaxf(m, c[i]);
    float s = 0.f;
    for (int i = 0; i < n; ++i) s += expf(c[i] - m);
    return logf(s) + m;
}
// The same log-sum-exp over three candidates of which some may be -inf (= dropped by the reference): the terms are added in the same
// order and a dropped one contributes exp(-inf) = +0, which changes no partial sum -- the same fp32 result without the candidate list
// (a dynamically indexed local array lives in local memory, and its round trips were most of a frame's time).  -inf if all three are.
__device__ __forceinline__ float lse3(float a, float b, float c) {
    const float m = fmaxf(fmaxf(a, b), c);
    if (!(m > -INFINITY)) return -INFINITY;
    const float s = (__expf(a - m) + __expf(b - m)) + __expf(c - m);      // (arguments <= 0, sum in [1, 3]: the fast forms are ~1e-7 here)
    return __logf(s) + m;
}

template <int DIR>
__global__ void __launch_bounds__(1024) ctc_trellis_kernel(CtcParams p) {
    extern __shared__ __align__(16) uint8_t smem_c[];
    const int S = 2 * p.L + 1, T = p.T, b = blockIdx.x;
    int *ext = reinterpret_cast<int *>(smem_c);                 // [S] blank-expanded target (ctc.py:8-29)
    float *row0 = reinterpret_cast<float *>(ext + S);           // [S] x 2: previous / current row
    float *row1 = row0 + S;
    const int Tb = (int)p.in_len[b];
    const int E = min(2 * (int)p.tg_len[b] + 1, S);
    for (int s = threadIdx.x; s < S; s += blockDim.x) ext[s] = (s & 1) ? (int)p.targets[(size_t)b * p.L + (s >> 1)] : p.blank;
    __syncthreads();
    float *tab = p.table ? p.table + (size_t)b * T * S : nullptr;
    auto emit = [&](int t, int s) { return p.lp[((size_t)t * p.B + b) * p.C + ext[s]]; };
    const float NEG = -INFINITY;

    // The gathered emission lp[t, b, ext[s]] of the thread's own position is loaded ONE FRAME AHEAD (issued before the barrier, used
    // after it), so the L2 latency of the gather is not on the time chain.  (Threads that own more than one position -- targets
    // longer than 511 labels -- gather the others in place.)
    const int s0 = threadIdx.x;
    if (DIR == 0) {
        // t = 0: positions 0 (blank) and 1 (first label) only (ctc.py:62-70)
        for (int s = threadIdx.x; s < S; s += blockDim.x) {
            float v = NEG;
            if (s == 0) v = emit(0, 0);
            else if (s == 1 && p.tg_len[b] > 0) v = emit(0, 1);
            row0[s] = v;
            if (tab) tab[s] = v;
        }
        float e_next = (s0 < E && 1 < Tb && 1 < T) ? emit(1, s0) : 0.f;
        __syncthreads();
        float *prev = row0, *cur = row1;
        for (int t = 1; t < T; ++t) {
            const bool live = t < Tb;
            const float e_own = e_next;
            if (s0 < E && t + 1 < Tb && t + 1 < T) e_next = emit(t + 1, s0);
            for (int s = threadIdx.x; s < S; s += blockDim.x) {
                float v = NEG;
                if (live && s < E) {
                    const float a0 = prev[s];
                    const float a1 = (s > 0) ? prev[s - 1] : NEG;
                    const float a2 = (s > 1 && ext[s] != ext[s - 2]) ? prev[s - 2] : NEG;
                    const float l = lse3(a0, a1, a2);
                    if (l > NEG) v = ((s == s0) ? e_own : emit(t, s)) + l;
                }
                // (rows at or after the utterance's length stay -inf in the table but do not advance the recursion: ctc.py:75-76)
                if (live) cur[s] = v;
                if (tab) tab[(size_t)t * S + s] = v;
            }
            __syncthreads();
            if (live) { float *tmp = prev; prev = cur; cur = tmp; }
        }
        if (threadIdx.x == 0 && p.loglik) {
            // the utterance may end in the last label or the last blank (ctc.py:104-119); prev = row in_len - 1
            float c[2];
            int n = 0;
            if (E >= 1) c[n++] = prev[E - 1];
            if (E >= 2) c[n++] = prev[E - 2];
            float m = c[0];
            for (int i = 1; i < n; ++i) m = fmaxf(m, c[i]);
            float out = NEG;
            if (m > NEG) {
                float ssum = 0.f;
                for (int i = 0; i < n; ++i) ssum += expf(c[i] - m);
                out = logf(ssum) + m;
            }
            p.loglik[b] = out;
        }
    } else {
        // rows t >= in_len stay -inf; row in_len - 1 holds the two terminal zeros (ctc.py:157-166).  The shared-memory row holds
        // beta[t+1][s] + lp[t+1][ext[s]] -- the term every one of its (up to three) consumers adds, formed once by the owner: the same
        // fp32 number -- or -inf where beta[t+1][s] is -inf (the reference drops those candidates).
        float *next = row0, *cur = row1;
        const int t_top = min(Tb, T) - 1;
        float e_next = (s0 < E && t_top >= 0) ? emit(t_top, s0) : 0.f;
        for (int t = T - 1; t >= 0; --t) {
            const float e_own = e_next;                                   // lp[t][ext[s0]] when t <= t_top
            if (s0 < E && t - 1 >= 0 && t - 1 <= t_top) e_next = emit(t - 1, s0);
            for (int s = threadIdx.x; s < S; s += blockDim.x) {
                float v = NEG;
                if (t == Tb - 1) {
                    if (s == E - 1 || s == E - 2) v = 0.f;
                } else if (t < Tb - 1 && s < E) {
                    const float b0 = next[s];
                    const float b1 = (s + 1 < E) ? next[s + 1] : NEG;
                    const float b2 = (s + 2 < E && ext[s] != ext[s + 2]) ? next[s + 2] : NEG;
                    v = lse3(b0, b1, b2);
                }
                float hat = NEG;
                if (v > NEG) hat = v + ((s == s0) ? e_own : emit(t, s));
                cur[s] = hat;
                if (tab) tab[(size_t)t * S + s] = v;
            }
            __syncthreads();
            float *tmp = next; next = cur; cur = tmp;
        }
    }
}

static int launch_ctc(int dir, const CtcParams &p, cudaStream_t s) {
    const int S = 2 * p.L + 1;
    const int threads = min(1024, ((S + 31) / 32) * 32);
    const size_t smem = (size_t)S * (sizeof(int) + 2 * sizeof(float));
    if (smem > 200 * 1024) return set_error(HMMB200_EUNSUPPORTED, "ctc: target length %d too long for one CTA's shared memory", p.L);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(dir == 0 ? ctc_trellis_kernel<0> : ctc_trellis_kernel<1>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return set_error(HMMB200_ELAUNCH, "ctc smem opt-in: %s", cudaGetErrorString(e));
    }
    if (dir == 0) ctc_trellis_kernel<0><<<p.B, threads, smem, s>>>(p);
    else ctc_trellis_kernel<1><<<p.B, threads, smem, s>>>(p);
    return check_launch("ctc_trellis_kernel");
}

// ----------------------------------------------------------------------------------------------------------
// dynamic time warping
// ----------------------------------------------------------------------------------------------------------
struct DtwParams {
    const float *dist;      // [P,N,M]
    int P, N, M, pattern;   // 0 symmetric, 1 asymmetric, 2 rabiner_juang (dtw.py:68-122)
    float *cost;            // [P,N,M]
    uint8_t *dir;           // [P,N,M] backpointer of every cell: 0 diagonal, 1 up (i-1), 2 left (j-1)
    int64_t *path_i, *path_j;   // [P,N+M-1], written front to back in path order
    int *path_len;          // [P]
};

__global__ void __launch_bounds__(1024) dtw_kernel(DtwParams p) {
    extern __shared__ __align__(16) float smem_w[];
    const int N = p.N, M = p.M, pr = blockIdx.x;
    // three anti-diagonals indexed by the row i: k-2, k-1 and the one being written
    float *d2 = smem_w, *d1 = smem_w + N, *d0 = smem_w + 2 * N;
    const float *dist = p.dist + (size_t)pr * N * M;
    float *cost = p.cost + (size_t)pr * N * M;
    uint8_t *dir = p.dir + (size_t)pr * N * M;
    // A thread's own row i0 = threadIdx.x walks j = k - i0 = 0, 1, 2, ... as k grows: its local distances are consecutive floats, loaded
    // one diagonal ahead (issued before the barrier, used after it) so that the load's latency is not part of a diagonal's time.
    const int i0 = threadIdx.x;
    float d_own = (i0 < N) ? dist[(size_t)i0 * M] : 0.f;               // (i0, 0): first used at k = i0
    const bool vec4 = (M % 4 == 0) && ((((uintptr_t)cost) & 15) == 0) && ((((uintptr_t)dir) & 3) == 0);
    float cb0 = 0.f, cb1 = 0.f, cb2 = 0.f;
    unsigned db = 0u;
    for (int k = 0; k <= N + M - 2; ++k) {
        const int i_lo = max(0, k - (M - 1)), i_hi = min(N - 1, k);
        const float d_cur = d_own;
        if (i0 < N && k + 1 >= i0 && k + 1 - i0 < M && k >= i0) d_own = dist[(size_t)i0 * M + (k + 1 - i0)];
        for (int i = threadIdx.x; i <= i_hi; i += blockDim.x) {            // (fixed rows per thread: the first one is i0)
            if (i < i_lo) continue;
            const int j = k - i;
            const float d = (i == i0) ? d_cur : dist[(size_t)i * M + j];
            float c;
            uint8_t bp = 0;
            if (k == 0) {
                c = d;
            } else {
                // predecessor with the lowest cost; exact ties go to the diagonal, then to (i-1, j), then to (i, j-1): the order in
                // which the reference's traceback compares (cost, i, j) tuples (dtw.py:129-143)
                float best = INFINITY;
                bool have = false;
                if (i > 0 && j > 0) { best = d2[i - 1]; bp = 0; have = true; }
                if (i > 0) { const float u = d1[i - 1]; if (!have || u < best) { best = u; bp = 1; have = true; } }
                if (j > 0) { const float l = d1[i]; if (!have || l < best) { best = l; bp = 2; have = true; } }
                if (p.pattern == 2) {
                    // rabiner_juang: the diagonal step costs twice the local distance, so the CHEAPEST total decides (dtw.py:104-122)
                    float tot = INFINITY;
                    if (i > 0 && j > 0) tot = __fadd_rn(d2[i - 1], __fmul_rn(2.f, d));
                    if (i > 0) tot = fminf(tot, __fadd_rn(d1[i - 1], d));
                    if (j > 0) tot = fminf(tot, __fadd_rn(d1[i], d));
                    c = tot;
                } else {
                    c = __fadd_rn(d, best);                   // symmetric: d + min(...); asymmetric: min(c + d) -- the same number
                }
            }
            d0[i] = c;
            if (i == i0 && vec4) {
                // the thread's own row is written left to right, one cell per diagonal: four cells at a time as one 16-byte and one
                // 4-byte store (cell-by-cell stores along an anti-diagonal touch a different sector per thread -- 64 store
                // transactions per warp and diagonal, which was the time of a diagonal)
                const int q = j & 3;                                         // (selects, not an indexed array: that would live in local memory)
                cb0 = (q == 0) ? c : cb0; cb1 = (q == 1) ? c : cb1; cb2 = (q == 2) ? c : cb2;
                db |= (unsigned)bp << (8 * q);
                if (q == 3) {
                    *reinterpret_cast<float4 *>(cost + (size_t)i * M + j - 3) = make_float4(cb0, cb1, cb2, c);
                    *reinterpret_cast<unsigned *>(dir + (size_t)i * M + j - 3) = db;
                    db = 0u;
                }
            } else {
                cost[(size_t)i * M + j] = c;
                dir[(size_t)i * M + j] = bp;
            }
        }
        __syncthreads();
        float *tmp = d2; d2 = d1; d1 = d0; d0 = tmp;
    }
    // traceback (dtw.py:124-151): the step is chosen on the COST of the three predecessors, which is what `dir` recorded
    // One walk (each hop is a dependent load of a direction byte), recorded backwards in shared memory; all threads then write the path
    // out in forward order.  [It was walked twice -- once for the length, once to fill -- by one thread with 64-bit global stores.]
    __shared__ int path_n;
    int *pbuf = reinterpret_cast<int *>(smem_w + 3 * N);               // [N + M - 1] packed (i << 16) | j   (N, M <= 65535)
    if (threadIdx.x == 0) {
        int i = N - 1, j = M - 1, len = 0;
        while (i > 0 || j > 0) {
            pbuf[len++] = (i << 16) | j;
            const uint8_t bp = dir[(size_t)i * M + j];
            if (bp == 0) { --i; --j; } else if (bp == 1) --i; else --j;
        }
        pbuf[len++] = 0;
        path_n = len;
        p.path_len[pr] = len;
    }
    __syncthreads();
    {
        int64_t *pi = p.path_i + (size_t)pr * (N + M - 1), *pj = p.path_j + (size_t)pr * (N + M - 1);
        const int len = path_n;
        for (int q = threadIdx.x; q < len; q += blockDim.x) {
            const int e = pbuf[len - 1 - q];
            pi[q] = e >> 16; pj[q] = e & 0xffff;
        }
    }
}

}  // namespace hmmb200

using namespace hmmb200;

HMMB200_EXPORT int hmmb200_ctc_trellis_f32(int direction, const float *log_probs, const int64_t *targets, const int64_t *input_lengths,
                                           const int64_t *target_lengths, int blank, int T, int B, int C, int L,
                                           float *table, float *loglik, void *stream) {
    if (T < 0 || B < 0 || C <= 0 || L < 0) return set_error(HMMB200_EINVAL, "ctc: bad shape T=%d B=%d C=%d L=%d", T, B, C, L);
    if (direction != 0 && direction != 1) return set_error(HMMB200_EINVAL, "ctc: direction must be 0 (forward) or 1 (backward)");
    if (blank < 0 || blank >= C) return set_error(HMMB200_EINVAL, "ctc: blank id %d outside [0, %d)", blank, C);
    if (T == 0 || B == 0) return HMMB200_OK;
    if (!log_probs || !input_lengths || !target_lengths || (L > 0 && !targets)) return set_error(HMMB200_EINVAL, "ctc: null argument");
    if (int rc = require_sm100()) return rc;
    CtcParams p;
    p.lp = log_probs; p.targets = targets; p.in_len = input_lengths; p.tg_len = target_lengths;
    p.blank = blank; p.T = T; p.B = B; p.C = C; p.L = L; p.table = table; p.loglik = (direction == 0) ? loglik : nullptr;
    return launch_ctc(direction, p, (cudaStream_t)stream);
}

HMMB200_EXPORT int hmmb200_dtw_f32(const float *dist, int n_pairs, int N, int M, int step_pattern, float *cost, void *dir_ws,
                                   int64_t *path_i, int64_t *path_j, int *path_len, void *stream) {
    if (n_pairs < 0 || N <= 0 || M <= 0) return set_error(HMMB200_EINVAL, "dtw: bad shape pairs=%d N=%d M=%d", n_pairs, N, M);
    if (step_pattern < 0 || step_pattern > 2) return set_error(HMMB200_EINVAL, "dtw: unknown step pattern %d", step_pattern);
    if (n_pairs == 0) return HMMB200_OK;
    if (!dist || !cost || !dir_ws || !path_i || !path_j || !path_len) return set_error(HMMB200_EINVAL, "dtw: null argument");
    if (int rc = require_sm100()) return rc;
    if (N > 65535 || M > 65535) return set_error(HMMB200_EUNSUPPORTED, "dtw: sequences of at most 65535 frames (got %d x %d)", N, M);
    const size_t smem = 3 * (size_t)N * sizeof(float) + ((size_t)N + M) * sizeof(int);     // three anti-diagonals + the path
    if (smem > 200 * 1024) return set_error(HMMB200_EUNSUPPORTED, "dtw: N=%d, M=%d exceed one CTA's shared memory (pass the shorter sequence first)", N, M);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(dtw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return set_error(HMMB200_ELAUNCH, "dtw smem opt-in: %s", cudaGetErrorString(e));
    }
    DtwParams p;
    p.dist = dist; p.P = n_pairs; p.N = N; p.M = M; p.pattern = step_pattern; p.cost = cost; p.dir = (uint8_t *)dir_ws;
    p.path_i = path_i; p.path_j = path_j; p.path_len = path_len;
    const int threads = min(1024, ((min(N, M) + 31) / 32) * 32);
    dtw_kernel<<<n_pairs, threads, smem, (cudaStream_t)stream>>>(p);
    return check_launch("dtw_kernel");
}
