// Column sum of fp32 partials: dst[j] = sum_s src[s][j].
//
// The weight gradients of the block's Linear layers (reference swin_block.py:141-143, window_attention.py:28-32; torch
// autograd there) are computed here as a token-split batched GEMM: dW_s = dy_s^T x_s for ~72 slices s of the token axis
// (fp32 partials [S][Cout*Cin], 0.6-8 MB), which streams dy and x at 4-4.8 TB/s where one split-K GEMM reached 1.7-3.3.
// This kernel is the final sum over the slices.  torch's generic reduce kernel needs 7-10 us for it; all loads of a
// thread are independent here (S/8 rows each), so the kernel is one load round trip long.
#include "common.cuh"

namespace pwa {

namespace {

constexpr int kRowGroups = 8;

__global__ void __launch_bounds__(32 * kRowGroups) colsum_f32_kernel(const float* __restrict__ src, float* __restrict__ dst, int S, long n) {
  __shared__ float red[kRowGroups][33];
  const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
  const long c = (long)blockIdx.x * 32 + lane;
  float a = 0.f;
  if (c < n) {
#pragma unroll 4
    for (int s = g; s < S; s += kRowGroups) a += __ldg(src + (size_t)s * n + c);
  }
  red[g][lane] = a;
  __syncthreads();
  if (g == 0 && c < n) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < kRowGroups; ++k) t += red[k][lane];
    dst[c] = t;
  }
}

}  // namespace

}  // namespace pwa

using namespace pwa;

extern "C" int pwa_colsum_f32(const float* src, float* dst, int S, int64_t n, void* stream) {
  PWA_CHECK_ARG(src != nullptr && dst != nullptr, "pwa_colsum_f32: null pointer");
  PWA_CHECK_ARG(S >= 1 && n >= 1, "pwa_colsum_f32: S=%d n=%lld", S, (long long)n);
  const long blocks = (n + 31) / 32;
  colsum_f32_kernel<<<(unsigned)blocks, 32 * kRowGroups, 0, (cudaStream_t)stream>>>(src, dst, S, (long)n);
  PWA_CUDA_OK(cudaGetLastError());
  return PWA_OK;
}
