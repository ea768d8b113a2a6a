// LayerNorm over the channel axis of window tokens, forward and backward, fused with the residual adds
// that surround it in the block (reference swin_block.py:216 attn_norm, :222 `x + shortcut`, :227 mlp_norm).
// HBM-bound row kernels for small C (48..768): a row is handled by a sub-warp group of G lanes, each lane
// owning 4-element vectors, statistics by shuffle reduction in fp32 (two-pass over registers: exact mean,
// then centred variance -- matches torch's numerics closely).  gamma/beta stay fp32 (no per-call casts).
//   forward : s = x (+ res) ; y = (s - mean) * rstd * gamma + beta ; saves mean, rstd   (also writes s if res)
//   backward: dx = rstd * (g - mean(g) - xhat * mean(g * xhat)) (+ dres), g = dy * gamma ;
//             dgamma += sum_rows dy * xhat ; dbeta += sum_rows dy   (register partials -> smem -> fp32 atomics)
#include "common.cuh"

namespace pwa {

constexpr int kLnThreads = 256;
constexpr int kLnMaxNV = 8;   // vectors of 4 per lane -> C <= 32 * 8 * 4 = 1024

template <typename T> struct Vec4;
template <> struct Vec4<float> {
  static __device__ __forceinline__ void load(const float* p, float (&v)[4]) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <> struct Vec4<__nv_bfloat16> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[4]) {
    const uint2 t = *reinterpret_cast<const uint2*>(p);
    const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&t.x);
    const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&t.y);
    v[0] = __low2float(a); v[1] = __high2float(a); v[2] = __low2float(b); v[3] = __high2float(b);
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[4]) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
    uint2 t;
    t.x = *reinterpret_cast<uint32_t*>(&a);
    t.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = t;
  }
};

template <int G>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <typename T, int G, int NV>
__global__ void __launch_bounds__(kLnThreads) ln_fwd_kernel(const T* __restrict__ x, const T* __restrict__ res,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            T* __restrict__ sum_out, T* __restrict__ y,
                                                            float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                            long rows, int C, float eps) {
  constexpr int RPB = kLnThreads / G;                  // rows per block-iteration
  const int gl = threadIdx.x % G, gr = threadIdx.x / G;
  const int nvec = C / 4;
  float gm[NV][4], bt[NV][4];
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int vi = gl + k * G;
    if (vi < nvec) {
      Vec4<float>::load(gamma + vi * 4, gm[k]);
      Vec4<float>::load(beta + vi * 4, bt[k]);
    }
  }
  const float invC = 1.f / (float)C;
  // the loop bound is uniform per CTA so that the sub-warp shuffles always run with all 32 lanes
  for (long base = (long)blockIdx.x * RPB; base < rows; base += (long)gridDim.x * RPB) {
    const long row = base + gr;
    const bool live = row < rows;
    float v[NV][4];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int vi = gl + k * G;
#pragma unroll
      for (int e = 0; e < 4; ++e) v[k][e] = 0.f;
      if (vi < nvec && live) {
        Vec4<T>::load(x + row * C + vi * 4, v[k]);
        if (res != nullptr) {
          float r[4];
          Vec4<T>::load(res + row * C + vi * 4, r);
#pragma unroll
          for (int e = 0; e < 4; ++e) v[k][e] += r[e];
          // the residual sum is stored in the I/O dtype and the statistics use the STORED value
          Vec4<T>::store(sum_out + row * C + vi * 4, v[k]);
          if (sizeof(T) == 2) {
#pragma unroll
            for (int e = 0; e < 4; ++e) v[k][e] = to_f32(from_f32<T>(v[k][e]));
          }
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) s += v[k][e];
      }
    }
    const float mean = group_sum<G>(s) * invC;
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int vi = gl + k * G;
      if (vi < nvec) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float d = v[k][e] - mean;
          q = fmaf(d, d, q);
        }
      }
    }
    const float rstd = rsqrtf(group_sum<G>(q) * invC + eps);
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int vi = gl + k * G;
      if (vi < nvec && live) {
        float o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) o[e] = fmaf((v[k][e] - mean) * rstd, gm[k][e], bt[k][e]);
        Vec4<T>::store(y + row * C + vi * 4, o);
      }
    }
    if (gl == 0 && live) {
      mean_out[row] = mean;
      rstd_out[row] = rstd;
    }
  }
}

template <typename T, int G, int NV>
__global__ void __launch_bounds__(kLnThreads) ln_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ x,
                                                            const float* __restrict__ gamma, const float* __restrict__ mean_in,
                                                            const float* __restrict__ rstd_in, const T* __restrict__ dres,
                                                            T* __restrict__ dx, float* __restrict__ dgamma,
                                                            float* __restrict__ dbeta, long rows, int C) {
  constexpr int RPB = kLnThreads / G;
  extern __shared__ float red[];                        // [2][C]
  const int gl = threadIdx.x % G, gr = threadIdx.x / G;
  const int nvec = C / 4;
  for (int i = threadIdx.x; i < 2 * C; i += kLnThreads) red[i] = 0.f;
  float gm[NV][4], ag[NV][4], ab[NV][4];
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int vi = gl + k * G;
    if (vi < nvec) Vec4<float>::load(gamma + vi * 4, gm[k]);
#pragma unroll
    for (int e = 0; e < 4; ++e) ag[k][e] = ab[k][e] = 0.f;
  }
  const float invC = 1.f / (float)C;
  for (long base = (long)blockIdx.x * RPB; base < rows; base += (long)gridDim.x * RPB) {
    const long row = base + gr;
    const bool live = row < rows;
    const float mean = live ? mean_in[row] : 0.f, rstd = live ? rstd_in[row] : 0.f;
    float g[NV][4], xh[NV][4];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int vi = gl + k * G;
#pragma unroll
      for (int e = 0; e < 4; ++e) g[k][e] = xh[k][e] = 0.f;
      if (vi < nvec && live) {
        float d[4], xv[4];
        Vec4<T>::load(dy + row * C + vi * 4, d);
        Vec4<T>::load(x + row * C + vi * 4, xv);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          xh[k][e] = (xv[e] - mean) * rstd;
          g[k][e] = d[e] * gm[k][e];
          s1 += g[k][e];
          s2 = fmaf(g[k][e], xh[k][e], s2);
          ag[k][e] = fmaf(d[e], xh[k][e], ag[k][e]);
          ab[k][e] += d[e];
        }
      }
    }
    s1 = group_sum<G>(s1) * invC;
    s2 = group_sum<G>(s2) * invC;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int vi = gl + k * G;
      if (vi < nvec && live) {
        float o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) o[e] = rstd * (g[k][e] - s1 - xh[k][e] * s2);
        if (dres != nullptr) {
          float r[4];
          Vec4<T>::load(dres + row * C + vi * 4, r);
#pragma unroll
          for (int e = 0; e < 4; ++e) o[e] += r[e];
        }
        Vec4<T>::store(dx + row * C + vi * 4, o);
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int vi = gl + k * G;
    if (vi < nvec) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        atomicAdd(&red[vi * 4 + e], ag[k][e]);
        atomicAdd(&red[C + vi * 4 + e], ab[k][e]);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += kLnThreads) {
    atomicAdd(&dgamma[i], red[i]);
    atomicAdd(&dbeta[i], red[C + i]);
  }
}

template <typename T, int G, int NV>
static int ln_launch(bool fwd, const void* a, const void* b, const float* gamma, const float* beta_or_mean, const float* rstd,
                     const void* res, void* o1, void* o2, float* f1, float* f2, long rows, int C, float eps, cudaStream_t st) {
  constexpr int RPB = kLnThreads / G;
  long blocks = (rows + RPB - 1) / RPB;
  const long cap = fwd ? 148L * 16 : 148L * 4;        // backward: fewer CTAs -> fewer global atomics
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  if (fwd) {
    ln_fwd_kernel<T, G, NV><<<(unsigned)blocks, kLnThreads, 0, st>>>((const T*)a, (const T*)res, gamma, beta_or_mean, (T*)o1,
                                                                      (T*)o2, f1, f2, rows, C, eps);
  } else {
    ln_bwd_kernel<T, G, NV><<<(unsigned)blocks, kLnThreads, 2 * C * sizeof(float), st>>>(
        (const T*)a, (const T*)b, gamma, beta_or_mean, rstd, (const T*)res, (T*)o1, f1, f2, rows, C);
  }
  PWA_CUDA_OK(cudaGetLastError());
  return PWA_OK;
}

template <typename T>
static int ln_dispatch(bool fwd, const void* a, const void* b, const float* gamma, const float* bm, const float* rstd,
                       const void* res, void* o1, void* o2, float* f1, float* f2, long rows, int C, float eps, cudaStream_t st) {
  const int nvec = C / 4;
#define LN_CASE(G, NV) return ln_launch<T, G, NV>(fwd, a, b, gamma, bm, rstd, res, o1, o2, f1, f2, rows, C, eps, st)
  if (nvec <= 4) LN_CASE(4, 1);
  if (nvec <= 8) LN_CASE(8, 1);
  if (nvec <= 16) LN_CASE(16, 1);
  if (nvec <= 32) LN_CASE(32, 1);
  if (nvec <= 64) LN_CASE(32, 2);
  if (nvec <= 128) LN_CASE(32, 4);
  if (nvec <= 256) LN_CASE(32, 8);
#undef LN_CASE
  set_error("layer norm: C=%d too large (max 1024)", C);
  return PWA_ERR_UNSUPPORTED;
}

}  // namespace pwa

using namespace pwa;

extern "C" int pwa_ln_fwd(const void* x, const void* res, const float* gamma, const float* beta, void* sum_out, void* y,
                          float* mean, float* rstd, int64_t rows, int C, float eps, int dtype, void* stream) {
  PWA_CHECK_ARG(x && gamma && beta && y && mean && rstd, "pwa_ln_fwd: null pointer");
  PWA_CHECK_ARG(res == nullptr || sum_out != nullptr, "pwa_ln_fwd: residual given without sum_out");
  PWA_CHECK_ARG(rows >= 0 && C > 0 && C % 4 == 0, "pwa_ln_fwd: need C %% 4 == 0 (C=%d)", C);
  PWA_CHECK_ARG(dtype == PWA_F32 || dtype == PWA_BF16, "pwa_ln_fwd: bad dtype %d", dtype);
  if (rows == 0) return PWA_OK;
  cudaStream_t st = (cudaStream_t)stream;
  return dtype == PWA_F32 ? ln_dispatch<float>(true, x, nullptr, gamma, beta, nullptr, res, sum_out, y, mean, rstd, rows, C, eps, st)
                          : ln_dispatch<__nv_bfloat16>(true, x, nullptr, gamma, beta, nullptr, res, sum_out, y, mean, rstd, rows, C, eps, st);
}

extern "C" int pwa_ln_bwd(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd,
                          const void* dres, void* dx, float* dgamma, float* dbeta, int64_t rows, int C, int dtype,
                          void* stream) {
  PWA_CHECK_ARG(dy && x && gamma && mean && rstd && dx && dgamma && dbeta, "pwa_ln_bwd: null pointer");
  PWA_CHECK_ARG(rows >= 0 && C > 0 && C % 4 == 0, "pwa_ln_bwd: need C %% 4 == 0 (C=%d)", C);
  PWA_CHECK_ARG(dtype == PWA_F32 || dtype == PWA_BF16, "pwa_ln_bwd: bad dtype %d", dtype);
  cudaStream_t st = (cudaStream_t)stream;
  PWA_CUDA_OK(cudaMemsetAsync(dgamma, 0, (size_t)C * 4, st));
  PWA_CUDA_OK(cudaMemsetAsync(dbeta, 0, (size_t)C * 4, st));
  if (rows == 0) return PWA_OK;
  return dtype == PWA_F32 ? ln_dispatch<float>(false, dy, x, gamma, mean, rstd, dres, dx, nullptr, dgamma, dbeta, rows, C, 0.f, st)
                          : ln_dispatch<__nv_bfloat16>(false, dy, x, gamma, mean, rstd, dres, dx, nullptr, dgamma, dbeta, rows, C, 0.f, st);
}
