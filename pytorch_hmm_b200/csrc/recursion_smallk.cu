// recursion_smallk.cu -- small-K (K <= 32) HMM recursions for sm_100a.
//
//   fb_sweep_kernel     forward and backward sweeps (two independent roles of one launch), scaled-probability
//                       space, G lanes per sequence, transition column/row in registers, warp shuffles.
//                       Replaces the per-time-step ATen launches of pytorch_hmm/hmm.py:95-117.
//   fb_combine_kernel   posterior / exp(log alpha) / exp(log beta) from the two scaled sweeps (hmm.py:120-128).
//   viterbi_kernel      max-plus recursion with packed uint8 backpointers in shared memory and a
//                       chunk-parallel on-device traceback (hmm.py:159-178; mixture_gaussian.py:312-336).
//
// Lane layout: a warp carries NS = 32/G sequences; lane (sub, j) owns state j of sequence `sub`.
// Every step broadcasts the K previous values with K shuffles and reduces in registers; nothing on the
// per-step critical path touches memory (emissions are prefetched a block of U frames ahead).
#include "common.cuh"

namespace hmmb200 {

// ----------------------------------------------------------------------------------------------------------
// emission -> per-frame scaled probability b~ and the log-scale m that was divided out
// ----------------------------------------------------------------------------------------------------------
template <int G>
__device__ __forceinline__ float group_max(float v) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL_MASK, v, o, G));
    return v;
}
template <int G>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o, G);
    return v;
}

template <int G>
__device__ __forceinline__ void emis_to_scaled(int mode, float eps, float e, bool lane_ok, float &bt, float &m) {
    if (mode == HMMB200_EMIS_PROB_FLOOR) {
        m = 0.f;
        bt = lane_ok ? e + eps : 0.f;
    } else if (mode == HMMB200_EMIS_LOG_EXP_FLOOR) {
        m = 0.f;
        bt = lane_ok ? __expf(e) + eps : 0.f;
    } else {
        float ev = lane_ok ? e : -INFINITY;
        m = group_max<G>(ev);
        if (!(m > -INFINITY)) m = 0.f;                 // all states impossible: keep the frame finite
        float b = lane_ok ? __expf(ev - m) : 0.f;
        if (mode == HMMB200_EMIS_LOG_NORM_FLOOR && lane_ok) b += eps;
        bt = b;
    }
}

// The log-emission the Viterbi recursion adds (fp32, same formula as the reference for each input kind).
template <int G>
__device__ __forceinline__ float emis_to_log(int mode, float eps, float e, bool lane_ok) {
    if (mode == HMMB200_EMIS_LOG) return e;
    if (mode == HMMB200_EMIS_PROB_FLOOR) return logf(e + eps);
    if (mode == HMMB200_EMIS_LOG_EXP_FLOOR) return logf(expf(e) + eps);
    float ev = lane_ok ? e : -INFINITY;
    float m = group_max<G>(ev);
    return logf(expf(e - m) + eps);
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
};

// One sweep over one warp's sequences.  DIR 0: alpha_t(j) = (sum_i alpha_{t-1}(i) P(i,j)) b_t(j)   (hmm.py:98-101)
//                                       DIR 1: beta_t(i)  = sum_j P(i,j) b_{t+1}(j) beta_{t+1}(j)   (hmm.py:113-117)
// Both are the same register recursion  w <- (sum_i shfl(w,i) * M[i]) * b~ * r  on w = alpha (DIR 0) or
// w = beta .* b~ (DIR 1), with M the column (DIR 0) or row (DIR 1) of P owned by the lane.  r is a lagged
// normaliser 1/sum(w) from the previous step, so it never sits on the dependent chain; its log is
// accumulated (in double) into the per-frame log-scale.
template <int G, int KP, int DIR>
__device__ __forceinline__ void fb_sweep(const FbParams &p) {
    constexpr int NS = 32 / G;
    constexpr int U = 16;
    const int lane = threadIdx.x & 31;
    const int sub = lane / G, j = lane % G;
    const int seq = blockIdx.x * NS + sub;
    const int K = p.K, T = p.T;
    const bool seq_ok = seq < p.B;
    const bool lane_ok = seq_ok && j < K;
    const int seq_c = seq_ok ? seq : p.B - 1;
    const int j_c = j < K ? j : K - 1;
    const int mode = p.mode;
    const float eps = p.eps;
    const bool add_m = (mode == HMMB200_EMIS_LOG) || (mode == HMMB200_EMIS_LOG_NORM_FLOOR && p.add_rowmax);

    float M[KP];
#pragma unroll
    for (int i = 0; i < KP; ++i) {
        float v = 0.f;
        if (lane_ok && i < K) v = (DIR == 0) ? __ldg(p.trans + i * K + j) : __ldg(p.trans + j * K + i);
        M[i] = v;
    }
    const float *ep = p.emis + (size_t)seq_c * T * K + j_c;
    float *out = ((DIR == 0) ? p.ws_a : p.ws_b) + (size_t)seq_c * T * K + j_c;
    float *outL = ((DIR == 0) ? p.ws_la : p.ws_lb) + (size_t)seq_c * T;

    // ---- step 0 -------------------------------------------------------------------------------------
    float w, bt0, m0;
    double L;
    float m_carry = 0.f;   // DIR 1: log-scale of the frame consumed by the previous step
    {
        const int f0 = (DIR == 0) ? 0 : T - 1;
        float e0 = __ldg(ep + (size_t)f0 * K);
        emis_to_scaled<G>(mode, eps, e0, lane_ok, bt0, m0);
        if (DIR == 0) {
            float pi = lane_ok ? __ldg(p.init + j) : 0.f;
            w = pi * bt0;
            L = add_m ? (double)m0 : 0.0;
            if (lane_ok) out[(size_t)f0 * K] = w;
        } else {
            w = bt0;
            L = 0.0;
            m_carry = add_m ? m0 : 0.f;
            if (lane_ok) out[(size_t)f0 * K] = 1.0f;
        }
        if (seq_ok && j == 0) outL[f0] = (float)L;
    }

    // ---- steps 1 .. T-1, emissions prefetched one block (U frames) ahead ---------------------------------
    float r_cur = 1.f;
    float eb[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        int n = 1 + u;
        int f = (DIR == 0) ? n : T - 1 - n;
        eb[u] = (n < T) ? __ldg(ep + (size_t)f * K) : 0.f;
    }
    for (int n0 = 1; n0 < T; n0 += U) {
        float en[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            int n = n0 + U + u;
            int f = (DIR == 0) ? n : T - 1 - n;
            en[u] = (n < T) ? __ldg(ep + (size_t)f * K) : 0.f;
        }
        float bt[U], mt[U];
#pragma unroll
        for (int u = 0; u < U; ++u) emis_to_scaled<G>(mode, eps, eb[u], lane_ok, bt[u], mt[u]);

#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int n = n0 + u;
            if (n < T) {
                const int f = (DIR == 0) ? n : T - 1 - n;
                float v[KP];
#pragma unroll
                for (int i = 0; i < KP; ++i) v[i] = __shfl_sync(FULL_MASK, w, i, G);
                float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
                float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
                for (int i = 0; i < KP; i += 4) {
                    a0 = fmaf(v[i + 0], M[i + 0], a0); s0 += v[i + 0];
                    a1 = fmaf(v[i + 1], M[i + 1], a1); s1 += v[i + 1];
                    a2 = fmaf(v[i + 2], M[i + 2], a2); s2 += v[i + 2];
                    a3 = fmaf(v[i + 3], M[i + 3], a3); s3 += v[i + 3];
                }
                const float acc = (a0 + a1) + (a2 + a3);
                const float S = (s0 + s1) + (s2 + s3);
                const float mb = bt[u] * r_cur;              // ready long before acc
                w = acc * mb;
                const float lr = __logf(r_cur);
                if (DIR == 0) {
                    L += (double)(add_m ? mt[u] : 0.f) - (double)lr;
                    if (lane_ok) out[(size_t)f * K] = w;
                } else {
                    L += (double)m_carry - (double)lr;
                    m_carry = add_m ? mt[u] : 0.f;
                    if (lane_ok) out[(size_t)f * K] = acc * r_cur;
                }
                if (seq_ok && j == 0) outL[f] = (float)L;
                r_cur = (S > 1e-30f && S < 1e30f) ? __fdividef(1.f, S) : 1.f;   // used by the NEXT step
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) eb[u] = en[u];
    }

    if (DIR == 0 && p.loglik != nullptr) {
        float tot = group_sum<G>(lane_ok ? w : 0.f);
        if (seq_ok && j == 0) p.loglik[seq] = (float)(L + (double)logf(tot));
    }
}

template <int G, int KP>
__global__ void __launch_bounds__(32) fb_sweep_kernel(FbParams p) {
    if (blockIdx.y == 0) fb_sweep<G, KP, 0>(p);
    else fb_sweep<G, KP, 1>(p);
}

// ----------------------------------------------------------------------------------------------------------
// combine: gamma = a.*b / sum, fwd = a*exp(la), bwd = b*exp(lb)           (hmm.py:120-128)
// ----------------------------------------------------------------------------------------------------------
struct CombineParams {
    const float *ws_a, *ws_b, *ws_la, *ws_lb;
    int64_t n_frames;
    int K;
    float *gamma, *fwd, *bwd, *log_alpha, *log_beta;
};

template <int VEC>
__global__ void __launch_bounds__(256) fb_combine_kernel(CombineParams p) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= p.n_frames) return;
    const int K = p.K;
    const float *a = p.ws_a + idx * K, *b = p.ws_b + idx * K;
    const float la = p.ws_la[idx], lb = p.ws_lb[idx];
    const float ea = expf(la), eb = expf(lb);
    float Z = 0.f;
    if (VEC == 4) {
        for (int k = 0; k < K; k += 4) {
            float4 x = *reinterpret_cast<const float4 *>(a + k), y = *reinterpret_cast<const float4 *>(b + k);
            Z += x.x * y.x + x.y * y.y + x.z * y.z + x.w * y.w;
        }
    } else {
        for (int k = 0; k < K; ++k) Z += a[k] * b[k];
    }
    const float inv = 1.f / Z;
    if (VEC == 4) {
        for (int k = 0; k < K; k += 4) {
            float4 x = *reinterpret_cast<const float4 *>(a + k), y = *reinterpret_cast<const float4 *>(b + k);
            if (p.gamma) *reinterpret_cast<float4 *>(p.gamma + idx * K + k) =
                make_float4(x.x * y.x * inv, x.y * y.y * inv, x.z * y.z * inv, x.w * y.w * inv);
            if (p.fwd) *reinterpret_cast<float4 *>(p.fwd + idx * K + k) = make_float4(x.x * ea, x.y * ea, x.z * ea, x.w * ea);
            if (p.bwd) *reinterpret_cast<float4 *>(p.bwd + idx * K + k) = make_float4(y.x * eb, y.y * eb, y.z * eb, y.w * eb);
            if (p.log_alpha) *reinterpret_cast<float4 *>(p.log_alpha + idx * K + k) =
                make_float4(logf(x.x) + la, logf(x.y) + la, logf(x.z) + la, logf(x.w) + la);
            if (p.log_beta) *reinterpret_cast<float4 *>(p.log_beta + idx * K + k) =
                make_float4(logf(y.x) + lb, logf(y.y) + lb, logf(y.z) + lb, logf(y.w) + lb);
        }
    } else {
        for (int k = 0; k < K; ++k) {
            float x = a[k], y = b[k];
            if (p.gamma) p.gamma[idx * K + k] = x * y * inv;
            if (p.fwd) p.fwd[idx * K + k] = x * ea;
            if (p.bwd) p.bwd[idx * K + k] = y * eb;
            if (p.log_alpha) p.log_alpha[idx * K + k] = logf(x) + la;
            if (p.log_beta) p.log_beta[idx * K + k] = logf(y) + lb;
        }
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
};

constexpr int VIT_THREADS = 128;

// Shared-memory layout (bytes), NS sequences per CTA:
//   [psi: NS*T*G if psi_in_smem][st: NS*T][exit: NS*n_chunks*G][entry: NS*n_chunks][final: NS ints + NS floats]
template <int G, int KP>
__global__ void __launch_bounds__(VIT_THREADS) viterbi_kernel(VitParams p) {
    constexpr int NS = 32 / G;
    constexpr int U = 16;
    extern __shared__ __align__(16) uint8_t smem[];
    const int K = p.K, T = p.T;
    const int tid = threadIdx.x;
    const int seq_base = blockIdx.x * NS;

    uint8_t *psi_s = smem;
    size_t off = p.psi_in_smem ? (size_t)NS * T * G : 0;
    uint8_t *st_s = smem + off;            off += (size_t)NS * T;
    uint8_t *exit_s = smem + off;          off += (size_t)NS * p.n_chunks * G;
    uint8_t *entry_s = smem + off;         off += (size_t)NS * p.n_chunks;
    off = (off + 15) & ~(size_t)15;
    int *final_s = reinterpret_cast<int *>(smem + off);

    // ---------------- recursion: warp 0 ----------------------------------------------------------------
    if (tid < 32) {
        const int lane = tid;
        const int sub = lane / G, j = lane % G;
        const int seq = seq_base + sub;
        const bool seq_ok = seq < p.B;
        const bool lane_ok = seq_ok && j < K;
        const int seq_c = seq_ok ? seq : p.B - 1;
        const int j_c = j < K ? j : K - 1;
        const int mode = p.mode;
        const float eps = p.eps;

        float M[KP];
#pragma unroll
        for (int i = 0; i < KP; ++i) M[i] = (lane_ok && i < K) ? __ldg(p.log_trans + i * K + j) : -INFINITY;

        const float *ep = p.emis + (size_t)seq_c * T * K + j_c;
        float *dout = p.delta ? p.delta + (size_t)seq_c * T * K + j_c : nullptr;
        uint8_t *pout = p.psi_out ? p.psi_out + (size_t)seq_c * T * K + j_c : nullptr;
        uint8_t *pst = p.psi_in_smem ? psi_s + (size_t)sub * T * G + j : p.psi_ws + (size_t)seq_c * T * G + j;

        float d;
        {
            float e0 = emis_to_log<G>(mode, eps, __ldg(ep), lane_ok);
            d = lane_ok ? __fadd_rn(__ldg(p.log_init + j), e0) : -INFINITY;
            if (lane_ok && dout) dout[0] = d;
            if (lane_ok && pout) pout[0] = 0;
            if (seq_ok || p.psi_in_smem) pst[0] = 0;
        }
        float eb[U];
#pragma unroll
        for (int u = 0; u < U; ++u) eb[u] = (1 + u < T) ? __ldg(ep + (size_t)(1 + u) * K) : 0.f;
        for (int t0 = 1; t0 < T; t0 += U) {
            float en[U];
#pragma unroll
            for (int u = 0; u < U; ++u) en[u] = (t0 + U + u < T) ? __ldg(ep + (size_t)(t0 + U + u) * K) : 0.f;
            float le[U];
#pragma unroll
            for (int u = 0; u < U; ++u) le[u] = emis_to_log<G>(mode, eps, eb[u], lane_ok);
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int t = t0 + u;
                if (t < T) {
                    float c[KP];
#pragma unroll
                    for (int i = 0; i < KP; ++i) c[i] = __fadd_rn(__shfl_sync(FULL_MASK, d, i, G), M[i]);
                    // exact max (order-independent), then the lowest index attaining it (torch.max tie rule)
                    float m4[KP / 4];
#pragma unroll
                    for (int i = 0; i < KP; i += 4) m4[i / 4] = fmaxf(fmaxf(c[i], c[i + 1]), fmaxf(c[i + 2], c[i + 3]));
                    float best = m4[0];
#pragma unroll
                    for (int q = 1; q < KP / 4; ++q) best = fmaxf(best, m4[q]);
                    d = __fadd_rn(best, le[u]);
                    int arg = 0;
#pragma unroll
                    for (int i = KP - 1; i >= 0; --i) arg = (c[i] == best) ? i : arg;
                    if (!lane_ok) d = -INFINITY;
                    if (lane_ok && dout) dout[(size_t)t * K] = d;
                    if (lane_ok && pout) pout[(size_t)t * K] = (uint8_t)arg;
                    if (seq_ok || p.psi_in_smem) pst[(size_t)t * G] = (uint8_t)arg;
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) eb[u] = en[u];
        }
        // final state: first index of the maximum (hmm.py:174)
        float bv = lane_ok ? d : -INFINITY;
        int bi = j;
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) {
            float ov = __shfl_xor_sync(FULL_MASK, bv, o, G);
            int oi = __shfl_xor_sync(FULL_MASK, bi, o, G);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (j == 0) {
            final_s[sub] = (bi < K) ? bi : 0;
            if (seq_ok && p.score) p.score[seq] = bv;
        }
        if (!p.psi_in_smem) __threadfence_block();
    }
    __syncthreads();

    // ---------------- chunk-parallel traceback: all threads --------------------------------------------
    // chunk c covers t in [1 + c*L, min(T-1, (c+1)*L)]; following psi from its top frame to its bottom frame
    // maps the state at t_hi to the state at t_lo - 1.
    const int L = p.chunk, nC = p.n_chunks;
    auto psi_at = [&](int sub, int seq_c, int t, int s) -> int {
        return p.psi_in_smem ? psi_s[((size_t)sub * T + t) * G + s] : p.psi_ws[((size_t)seq_c * T + t) * G + s];
    };
    // phase A: exit state for every (sequence, chunk, entry state)
    for (int task = tid; task < NS * nC * K; task += VIT_THREADS) {
        const int e = task % K, c = (task / K) % nC, sub = task / (K * nC);
        const int seq_c = min(seq_base + sub, p.B - 1);
        const int t_lo = 1 + c * L, t_hi = min(T - 1, t_lo + L - 1);
        int s = e;
        for (int t = t_hi; t >= t_lo; --t) s = psi_at(sub, seq_c, t, s);
        exit_s[((size_t)sub * nC + c) * G + e] = (uint8_t)s;
    }
    __syncthreads();
    // phase B: the true entry state of every chunk (serial over chunks, one thread per sequence)
    if (tid < NS) {
        int s = final_s[tid];
        st_s[(size_t)tid * T + (T - 1)] = (uint8_t)s;
        for (int c = nC - 1; c >= 0; --c) {
            entry_s[(size_t)tid * nC + c] = (uint8_t)s;
            s = exit_s[((size_t)tid * nC + c) * G + s];
        }
    }
    __syncthreads();
    // phase C: re-walk every chunk from its true entry state, recording the path
    for (int task = tid; task < NS * nC; task += VIT_THREADS) {
        const int c = task % nC, sub = task / nC;
        const int seq_c = min(seq_base + sub, p.B - 1);
        const int t_lo = 1 + c * L, t_hi = min(T - 1, t_lo + L - 1);
        int s = entry_s[(size_t)sub * nC + c];
        for (int t = t_hi; t >= t_lo; --t) {
            s = psi_at(sub, seq_c, t, s);
            st_s[(size_t)sub * T + (t - 1)] = (uint8_t)s;
        }
    }
    __syncthreads();
    // phase D: coalesced int64 store
    for (int i = tid; i < NS * T; i += VIT_THREADS) {
        const int sub = i / T, t = i % T;
        const int seq = seq_base + sub;
        if (seq < p.B) p.states[(size_t)seq * T + t] = (int64_t)st_s[i];
    }
}

// ----------------------------------------------------------------------------------------------------------
// host-side dispatch
// ----------------------------------------------------------------------------------------------------------
template <int G, int KP>
static int launch_fb(const FbParams &p, cudaStream_t s) {
    constexpr int NS = 32 / G;
    dim3 grid((p.B + NS - 1) / NS, 2);
    fb_sweep_kernel<G, KP><<<grid, 32, 0, s>>>(p);
    return check_launch("fb_sweep_kernel");
}

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

static size_t vit_smem_bytes(int NS, int T, int G, int nC, bool psi_in_smem) {
    size_t off = psi_in_smem ? (size_t)NS * T * G : 0;
    off += (size_t)NS * T + (size_t)NS * nC * G + (size_t)NS * nC;
    off = (off + 15) & ~(size_t)15;
    return off + NS * sizeof(int) * 2;
}

static void vit_plan(int T, int G, int &L, int &nC, bool &psi_in_smem, size_t &smem) {
    const int NS = 32 / G;
    L = 64;
    if (T - 1 > 64 * 1024) L = (T - 1 + 1023) / 1024;
    nC = (T <= 1) ? 0 : (T - 1 + L - 1) / L;
    psi_in_smem = true;
    smem = vit_smem_bytes(NS, T, G, nC, true);
    if (smem > 200 * 1024) {
        psi_in_smem = false;
        smem = vit_smem_bytes(NS, T, G, nC, false);
    }
}

template <int G, int KP>
static int launch_vit(VitParams p, cudaStream_t s) {
    constexpr int NS = 32 / G;
    bool in_smem; size_t smem;
    vit_plan(p.T, G, p.chunk, p.n_chunks, in_smem, smem);
    p.psi_in_smem = in_smem ? 1 : 0;
    if (smem > 200 * 1024) return set_error(HMMB200_EUNSUPPORTED, "viterbi: T=%d too long for the traceback tables", p.T);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(viterbi_kernel<G, KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return set_error(HMMB200_ELAUNCH, "viterbi smem opt-in: %s", cudaGetErrorString(e));
    }
    dim3 grid((p.B + NS - 1) / NS);
    viterbi_kernel<G, KP><<<grid, VIT_THREADS, smem, s>>>(p);
    return check_launch("viterbi_kernel");
}

#define DISPATCH_GK(FN, K, ...)                                                           \
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

static int dispatch_fb(const FbParams &p, cudaStream_t s) { DISPATCH_GK(launch_fb, p.K, p, s); }
static int dispatch_vit(const VitParams &p, cudaStream_t s) { DISPATCH_GK(launch_vit, p.K, p, s); }

}  // namespace hmmb200

using namespace hmmb200;

HMMB200_EXPORT size_t hmmb200_fb_workspace_bytes(int B, int T, int K) {
    if (B <= 0 || T <= 0 || K <= 0) return 0;
    size_t n = (size_t)B * T;
    return 2 * align256(n * K * sizeof(float)) + 2 * align256(n * sizeof(float));
}

HMMB200_EXPORT int hmmb200_forward_backward_f32(const float *emis, int emis_mode, float floor_eps, int add_rowmax,
                                                const float *trans_prob, const float *init_prob, int B, int T, int K,
                                                float *gamma, float *fwd_prob, float *bwd_prob,
                                                float *log_alpha, float *log_beta, float *loglik,
                                                void *workspace, size_t workspace_bytes, void *stream) {
    if (B < 0 || T < 0 || K <= 0) return set_error(HMMB200_EINVAL, "forward_backward: bad shape B=%d T=%d K=%d", B, T, K);
    if (B == 0 || T == 0) return HMMB200_OK;
    if (K > 32) return set_error(HMMB200_EUNSUPPORTED, "forward_backward: small-K path covers K <= 32 (got %d)", K);
    if (!emis || !trans_prob || !init_prob) return set_error(HMMB200_EINVAL, "forward_backward: null input");
    if (emis_mode < 0 || emis_mode > 3) return set_error(HMMB200_EINVAL, "forward_backward: bad emis_mode %d", emis_mode);
    if (!workspace || workspace_bytes < hmmb200_fb_workspace_bytes(B, T, K))
        return set_error(HMMB200_EWORKSPACE, "forward_backward: workspace %zu < %zu bytes", workspace_bytes,
                         hmmb200_fb_workspace_bytes(B, T, K));
    if (int rc = require_sm100()) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    size_t n = (size_t)B * T;
    uint8_t *w = (uint8_t *)workspace;
    FbParams p;
    p.emis = emis; p.mode = emis_mode; p.eps = floor_eps; p.add_rowmax = add_rowmax;
    p.trans = trans_prob; p.init = init_prob; p.B = B; p.T = T; p.K = K;
    p.ws_a = (float *)w;  w += align256(n * K * sizeof(float));
    p.ws_b = (float *)w;  w += align256(n * K * sizeof(float));
    p.ws_la = (float *)w; w += align256(n * sizeof(float));
    p.ws_lb = (float *)w;
    p.loglik = loglik;
    if (int rc = dispatch_fb(p, s)) return rc;
    if (gamma || fwd_prob || bwd_prob || log_alpha || log_beta) {
        CombineParams c;
        c.ws_a = p.ws_a; c.ws_b = p.ws_b; c.ws_la = p.ws_la; c.ws_lb = p.ws_lb;
        c.n_frames = (int64_t)n; c.K = K;
        c.gamma = gamma; c.fwd = fwd_prob; c.bwd = bwd_prob; c.log_alpha = log_alpha; c.log_beta = log_beta;
        const int threads = 256;
        const unsigned blocks = (unsigned)((n + threads - 1) / threads);
        auto al16 = [](const void *q) { return q == nullptr || ((uintptr_t)q & 15) == 0; };
        if (K % 4 == 0 && al16(gamma) && al16(fwd_prob) && al16(bwd_prob) && al16(log_alpha) && al16(log_beta))
            fb_combine_kernel<4><<<blocks, threads, 0, s>>>(c);
        else
            fb_combine_kernel<1><<<blocks, threads, 0, s>>>(c);
        if (int rc = check_launch("fb_combine_kernel")) return rc;
    }
    return HMMB200_OK;
}

HMMB200_EXPORT size_t hmmb200_viterbi_workspace_bytes(int B, int T, int K) {
    if (B <= 0 || T <= 0 || K <= 0 || K > 32) return 0;
    int G = group_lanes(K), L, nC; bool in_smem; size_t smem;
    vit_plan(T, G, L, nC, in_smem, smem);
    return in_smem ? 0 : (size_t)B * T * G;
}

HMMB200_EXPORT int hmmb200_viterbi_f32(const float *emis, int emis_mode, float floor_eps,
                                       const float *log_trans, const float *log_init, int B, int T, int K,
                                       float *delta, uint8_t *psi, int64_t *states, float *score,
                                       void *workspace, size_t workspace_bytes, void *stream) {
    if (B < 0 || T < 0 || K <= 0) return set_error(HMMB200_EINVAL, "viterbi: bad shape B=%d T=%d K=%d", B, T, K);
    if (B == 0 || T == 0) return HMMB200_OK;
    if (K > 32) return set_error(HMMB200_EUNSUPPORTED, "viterbi: small-K path covers K <= 32 (got %d)", K);
    if (!emis || !log_trans || !log_init || !states) return set_error(HMMB200_EINVAL, "viterbi: null argument");
    if (emis_mode < 0 || emis_mode > 3) return set_error(HMMB200_EINVAL, "viterbi: bad emis_mode %d", emis_mode);
    size_t need = hmmb200_viterbi_workspace_bytes(B, T, K);
    if (need && (!workspace || workspace_bytes < need))
        return set_error(HMMB200_EWORKSPACE, "viterbi: workspace %zu < %zu bytes", workspace_bytes, need);
    if (int rc = require_sm100()) return rc;
    VitParams p;
    p.emis = emis; p.mode = emis_mode; p.eps = floor_eps; p.log_trans = log_trans; p.log_init = log_init;
    p.B = B; p.T = T; p.K = K; p.delta = delta; p.psi_out = psi; p.states = states; p.score = score;
    p.psi_ws = (uint8_t *)workspace; p.psi_in_smem = 1; p.chunk = 64; p.n_chunks = 0;
    return dispatch_vit(p, (cudaStream_t)stream);
}
