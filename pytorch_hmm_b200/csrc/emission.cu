// emission.cu -- diagonal-Gaussian / GMM emission log-likelihoods for sm_100a.
//
//   gmm_pack_kernel            layer parameters -> packed standardised form (once per parameter update)
//   gmm_emission_fp32_kernel   CUDA-core path: packed fp32x2 FMAs (FFMA2), parameters broadcast from shared memory,
//                              frames staged through shared memory, mixture log-sum-exp in registers
//   gmm_emission_generic_kernel any (K, C, D); correctness fallback
//
// Replaces the [B,T,K,C,D] broadcast temporaries of pytorch_hmm/mixture_gaussian.py:175-214 (7.9 GB each at the
// headline shape), hmm_layer.py:283-321 and hsmm.py:194-204 with one pass over x.
//
// Arithmetic: with s = 1/sigma and nms = -mu/sigma, u = fma(x, s, nms) = (x - mu)/sigma and
//   l_kc = const_kc - 0.5 * sum_d u^2,   const_kc = log w_kc - 0.5 * (sum_d log var + D log 2pi)   (const in double)
// which has the accuracy of the reference's difference form at two FMAs per (frame, component, dim).
#include "common.cuh"

namespace hmmb200 {

struct EmisParams {
    const float *x;
    const float *packed;
    int64_t n_frames;
    int K, C, D;
    int NP;            // ceil(K*C / 2): component pairs in the packed layout
    float *logb;
    const float *skip_if_one;   // when non-null and *skip_if_one == 1 the tensor-core kernel has produced logb already
};

// packed layout: float4 prm[D][NP] = (s_{2p}, s_{2p+1}, nms_{2p}, nms_{2p+1}); then float cst[2*NP].
// One warp per component slot (lanes over the dimensions; the double-precision exp per element is the cost).
__global__ void __launch_bounds__(32) gmm_pack_kernel(const float *means, const float *log_vars, float scale, const float *logw,
                                                      int KC, int D, int NP, float *packed) {
    const int kc = blockIdx.x, lane = threadIdx.x;
    if (kc >= 2 * NP) return;
    float *cst = packed + (size_t)D * NP * 4;
    const int pr = kc >> 1, hi = kc & 1;
    if (kc >= KC) {
        for (int d = lane; d < D; d += 32) {
            packed[((size_t)d * NP + pr) * 4 + hi] = 0.f;
            packed[((size_t)d * NP + pr) * 4 + 2 + hi] = 0.f;
        }
        if (lane == 0) cst[kc] = -INFINITY;
        return;
    }
    double sum_lv = 0.0;
    for (int d = lane; d < D; d += 32) {
        const double lv = (double)scale * (double)log_vars[(size_t)kc * D + d];
        const double s = exp(-0.5 * lv);
        packed[((size_t)d * NP + pr) * 4 + hi] = (float)s;
        packed[((size_t)d * NP + pr) * 4 + 2 + hi] = (float)(-(double)means[(size_t)kc * D + d] * s);
        sum_lv += lv;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum_lv += __shfl_xor_sync(FULL_MASK, sum_lv, o);
    if (lane == 0) {
        const double lw = logw ? (double)logw[kc] : 0.0;
        cst[kc] = (float)(lw - 0.5 * (sum_lv + (double)D * 1.8378770664093454835606594728112));
    }
}

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    unsigned long long ra = *reinterpret_cast<unsigned long long *>(&a);
    unsigned long long rb = *reinterpret_cast<unsigned long long *>(&b);
    unsigned long long rc = *reinterpret_cast<unsigned long long *>(&c);
    unsigned long long rd;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2 *>(&rd);
}

// The reference's private logsumexp over the C components of one state (mixture_gaussian.py:141-155).
template <int C>
__device__ __forceinline__ float own_lse(const float (&l)[C]) {
    if (C == 1) return l[0];
    float m = l[0];
#pragma unroll
    for (int c = 1; c < C; ++c) m = fmaxf(m, l[c]);
    if (isinf(m)) m = 0.f;
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) s += expf(l[c] - m);
    return logf(fmaxf(s, 1e-8f)) + m;
}

constexpr int EM_THREADS = 64;
constexpr int EM_FRAMES = 2;                       // frames per thread
constexpr int EM_TILE = EM_THREADS * EM_FRAMES;    // frames per CTA tile

template <int C, int NPAIR>
__global__ void __launch_bounds__(EM_THREADS) gmm_emission_fp32_kernel(EmisParams p) {
    extern __shared__ __align__(16) float smem_f[];
    if (p.skip_if_one != nullptr && *p.skip_if_one == 1.f) return;
    const int D = p.D, K = p.K, NP = p.NP;
    const int pitch = D + 1;
    float4 *prm = reinterpret_cast<float4 *>(smem_f);         // [D][NPAIR]
    float *cst = smem_f + (size_t)D * NPAIR * 4;               // [2*NPAIR]
    float *xs = cst + 2 * NPAIR;                               // [EM_TILE][D+1]
    const int tid = threadIdx.x;

    const float4 *pk4 = reinterpret_cast<const float4 *>(p.packed);
    for (int i = tid; i < D * NPAIR; i += EM_THREADS) {
        const int d = i / NPAIR, pp = i % NPAIR;
        prm[i] = (pp < NP) ? __ldg(pk4 + (size_t)d * NP + pp) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int i = tid; i < 2 * NPAIR; i += EM_THREADS) cst[i] = (i < 2 * NP) ? __ldg(p.packed + (size_t)D * NP * 4 + i) : -INFINITY;

    const int64_t n_tiles = (p.n_frames + EM_TILE - 1) / EM_TILE;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t base = tile * EM_TILE;
        const int nf = (int)min((int64_t)EM_TILE, p.n_frames - base);
        __syncthreads();                                        // previous tile fully consumed; params visible
        const float *xg = p.x + base * D;
        if ((D & 3) == 0 && (((uintptr_t)xg) & 15) == 0) {
            const float4 *xg4 = reinterpret_cast<const float4 *>(xg);
            const int D4 = D >> 2;
            for (int i = tid; i < nf * D4; i += EM_THREADS) {
                const int f = i / D4, d4 = i % D4;
                float4 v = __ldg(xg4 + i);
                float *dst = xs + f * pitch + d4 * 4;
                dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
            }
        } else {
            for (int i = tid; i < nf * D; i += EM_THREADS) xs[(i / D) * pitch + (i % D)] = __ldg(xg + i);
        }
        __syncthreads();

        float2 acc[EM_FRAMES][NPAIR];
#pragma unroll
        for (int f = 0; f < EM_FRAMES; ++f)
#pragma unroll
            for (int pp = 0; pp < NPAIR; ++pp) acc[f][pp] = make_float2(0.f, 0.f);
        const float *x0 = xs + tid * pitch;
        const float *x1 = xs + (tid + EM_THREADS) * pitch;
        for (int d = 0; d < D; ++d) {
            const float xa = x0[d], xb = x1[d];
            const float2 xa2 = make_float2(xa, xa), xb2 = make_float2(xb, xb);
            const float4 *row = prm + d * NPAIR;
#pragma unroll
            for (int pp = 0; pp < NPAIR; ++pp) {
                const float4 q = row[pp];
                const float2 s2 = make_float2(q.x, q.y), m2 = make_float2(q.z, q.w);
                const float2 ua = ffma2(xa2, s2, m2);
                acc[0][pp] = ffma2(ua, ua, acc[0][pp]);
                const float2 ub = ffma2(xb2, s2, m2);
                acc[1][pp] = ffma2(ub, ub, acc[1][pp]);
            }
        }
        constexpr int KMAX = (2 * NPAIR) / C;
#pragma unroll
        for (int f = 0; f < EM_FRAMES; ++f) {
            const int64_t frame = base + tid + f * EM_THREADS;
            if (frame < p.n_frames) {
                float *o = p.logb + frame * K;
#pragma unroll
                for (int k = 0; k < KMAX; ++k) {
                    if (k < K) {
                        float l[C];
#pragma unroll
                        for (int c = 0; c < C; ++c) {
                            const int slot = k * C + c;
                            const float a = (slot & 1) ? acc[f][slot >> 1].y : acc[f][slot >> 1].x;
                            l[c] = fmaf(-0.5f, a, cst[slot]);
                        }
                        o[k] = own_lse<C>(l);
                    }
                }
            }
        }
    }
}

// Fallback for shapes the register-tiled kernel does not cover: one thread per (frame, state).
__global__ void __launch_bounds__(128) gmm_emission_generic_kernel(EmisParams p) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p.skip_if_one != nullptr && *p.skip_if_one == 1.f) return;
    if (idx >= p.n_frames * p.K) return;
    const int64_t n = idx / p.K;
    const int k = (int)(idx % p.K);
    const int D = p.D, C = p.C, NP = p.NP;
    const float *xn = p.x + n * D;
    const float *cst = p.packed + (size_t)D * NP * 4;
    float m = -INFINITY, first = 0.f;
    // two passes over the components keep the reference's max-then-sum order without a local array
    for (int pass = 0; pass < 2; ++pass) {
        float s = 0.f;
        for (int c = 0; c < C; ++c) {
            const int kc = k * C + c, pr = kc >> 1, hi = kc & 1;
            float a = 0.f;
            for (int d = 0; d < D; ++d) {
                const float sd = __ldg(p.packed + ((size_t)d * NP + pr) * 4 + hi);
                const float nm = __ldg(p.packed + ((size_t)d * NP + pr) * 4 + 2 + hi);
                const float u = fmaf(xn[d], sd, nm);
                a = fmaf(u, u, a);
            }
            const float l = fmaf(-0.5f, a, cst[kc]);
            if (pass == 0) { m = fmaxf(m, l); first = l; }
            else s += expf(l - m);
        }
        if (pass == 0) {
            if (C == 1) { p.logb[idx] = first; return; }
            if (isinf(m)) m = 0.f;
        } else {
            p.logb[idx] = logf(fmaxf(s, 1e-8f)) + m;
        }
    }
}

template <int C, int NPAIR>
static int launch_emission(const EmisParams &p, cudaStream_t s) {
    const size_t smem = ((size_t)p.D * NPAIR * 4 + 2 * NPAIR + (size_t)EM_TILE * (p.D + 1)) * sizeof(float);
    if (smem > 220 * 1024) return 1;   // caller falls back to the generic kernel
    cudaError_t e = cudaFuncSetAttribute(gmm_emission_fp32_kernel<C, NPAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_error(HMMB200_ELAUNCH, "emission smem opt-in: %s", cudaGetErrorString(e));
    int dev = 0, sms = 148, per_sm = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gmm_emission_fp32_kernel<C, NPAIR>, EM_THREADS, smem);
    if (per_sm < 1) per_sm = 1;
    const int64_t n_tiles = (p.n_frames + EM_TILE - 1) / EM_TILE;
    const int grid = (int)min((int64_t)sms * per_sm, n_tiles);
    gmm_emission_fp32_kernel<C, NPAIR><<<grid, EM_THREADS, smem, s>>>(p);
    return check_launch("gmm_emission_fp32_kernel");
}

template <int C>
static int launch_emission_c(const EmisParams &p, cudaStream_t s) {
    const int KC = p.K * p.C;
    if (KC <= 12) return launch_emission<C, 6>(p, s);
    if (KC <= 24) return launch_emission<C, 12>(p, s);
    if (KC <= 48) return launch_emission<C, 24>(p, s);
    return 1;
}

}  // namespace hmmb200

using namespace hmmb200;

static size_t fp32_section_floats(int K, int C, int D) {
    const size_t NP = ((size_t)K * C + 1) / 2;
    return ((size_t)D * NP * 4 + 2 * NP + 3) & ~(size_t)3;          // keeps the tensor-core section 16-byte aligned
}

HMMB200_EXPORT size_t hmmb200_gmm_packed_floats(int K, int C, int D) {
    if (K <= 0 || C <= 0 || D <= 0) return 0;
    return fp32_section_floats(K, C, D) + tc_floats(K, C, D);
}

HMMB200_EXPORT int hmmb200_gmm_pack_f32(const float *means, const float *log_vars, float log_var_scale,
                                        const float *log_weights, int K, int C, int D, float *packed, void *stream) {
    if (K <= 0 || C <= 0 || D <= 0) return set_error(HMMB200_EINVAL, "gmm_pack: bad shape K=%d C=%d D=%d", K, C, D);
    if (!means || !log_vars || !packed) return set_error(HMMB200_EINVAL, "gmm_pack: null argument");
    if (!log_weights && C != 1) return set_error(HMMB200_EINVAL, "gmm_pack: log_weights may be NULL only when C == 1");
    if (int rc = require_sm100()) return rc;
    const int KC = K * C, NP = (KC + 1) / 2;
    gmm_pack_kernel<<<2 * NP, 32, 0, (cudaStream_t)stream>>>(means, log_vars, log_var_scale, log_weights, KC, D, NP, packed);
    if (int rc = check_launch("gmm_pack_kernel")) return rc;
    if (tc_shape_ok(K, C, D))
        return launch_pack_tc(means, log_vars, log_var_scale, log_weights, K, C, D, packed + fp32_section_floats(K, C, D),
                              (cudaStream_t)stream);
    return HMMB200_OK;
}

// 1 if the packed parameters take the tensor-core kernel (shape supported AND every weight inside the fp16 range), 0 if not.
// SYNCHRONISES the stream (it reads the flag the pack kernel wrote): call it once after packing, never per batch.
HMMB200_EXPORT int hmmb200_gmm_pack_on_tensor_cores(const float *packed, int K, int C, int D, void *stream) {
    if (!packed || K <= 0 || C <= 0 || D <= 0) return set_error(HMMB200_EINVAL, "gmm_pack_on_tensor_cores: bad argument");
    if (!tc_shape_ok(K, C, D)) return 0;
    float flag = 0.f;
    cudaError_t e = cudaMemcpyAsync(&flag, packed + fp32_section_floats(K, C, D), sizeof(float), cudaMemcpyDeviceToHost, (cudaStream_t)stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)stream);
    if (e != cudaSuccess) return set_error(HMMB200_ELAUNCH, "gmm_pack_on_tensor_cores: %s", cudaGetErrorString(e));
    return flag != 0.f ? 1 : 0;
}

static int gmm_emission_impl(const float *x, const float *packed, int64_t n_frames, int K, int C, int D, float *logb, void *stream,
                             int tc_known) {
    return gmm_emission_dispatch(x, packed, n_frames, K, C, D, logb, nullptr, (cudaStream_t)stream, tc_known, nullptr);
}

HMMB200_EXPORT int hmmb200_gmm_emission_f32(const float *x, const float *packed, int64_t n_frames, int K, int C, int D,
                                            float *logb, void *stream) {
    return gmm_emission_impl(x, packed, n_frames, K, C, D, logb, stream, 0);
}

// Same, for a caller that has checked hmmb200_gmm_pack_on_tensor_cores() == 1 for `packed`: only the tcgen05 kernel is
// launched (the plain entry point also enqueues the fp32 kernel, which then exits at once on a device-side flag).
HMMB200_EXPORT int hmmb200_gmm_emission_tc_f32(const float *x, const float *packed, int64_t n_frames, int K, int C, int D,
                                               float *logb, void *stream) {
    return gmm_emission_impl(x, packed, n_frames, K, C, D, logb, stream, 1);
}

namespace hmmb200 {
int gmm_emission_dispatch(const float *x, const float *packed, int64_t n_frames, int K, int C, int D, float *logb, float *comp,
                          cudaStream_t s, int tc_known, const float **tc_flag_out) {
    if (tc_flag_out) *tc_flag_out = nullptr;
    if (n_frames < 0 || K <= 0 || C <= 0 || D <= 0) return set_error(HMMB200_EINVAL, "gmm_emission: bad shape");
    if (n_frames == 0) return HMMB200_OK;
    if (!x || !packed || !logb) return set_error(HMMB200_EINVAL, "gmm_emission: null argument");
    if (((uintptr_t)packed & 15) != 0) return set_error(HMMB200_EINVAL, "gmm_emission: packed must be 16-byte aligned");
    if (int rc = require_sm100()) return rc;
    EmisParams p;
    p.x = x; p.packed = packed; p.n_frames = n_frames; p.K = K; p.C = C; p.D = D; p.NP = (K * C + 1) / 2; p.logb = logb;
    p.skip_if_one = nullptr;
    // tensor-core path first; it declines on the device (flag = 0) when the parameters leave the fp16 range, in which
    // case the fp32 kernel below does the work.  Exactly one of the two kernels computes.
    if (tc_shape_ok(K, C, D) && (((uintptr_t)x) & 15) == 0) {
        const float *tc = packed + fp32_section_floats(K, C, D);
        int trc = launch_emission_tc(x, tc, packed, n_frames, K, C, D, logb, s, comp);
        if (trc < 0) return trc;
        if (trc == 0) {
            if (tc_flag_out) *tc_flag_out = tc;
            if (tc_known) return HMMB200_OK;
            p.skip_if_one = tc;
        }
    }
    int rc = 1;
    switch (C) {
        case 1: rc = launch_emission_c<1>(p, s); break;
        case 2: rc = launch_emission_c<2>(p, s); break;
        case 3: rc = launch_emission_c<3>(p, s); break;
        case 4: rc = launch_emission_c<4>(p, s); break;
        default: rc = 1; break;
    }
    if (rc <= 0) return rc;
    const int64_t total = n_frames * K;
    const int threads = 128;
    gmm_emission_generic_kernel<<<(unsigned)((total + threads - 1) / threads), threads, 0, s>>>(p);
    return check_launch("gmm_emission_generic_kernel");
}
}  // namespace hmmb200
