// Throughput micro-benchmarks for the building blocks of the attention kernels (run on a B200; test
// infrastructure, not part of libpwa_b200.so).  Prints per-SM rates in units / SM clock:
//   tmem_ld   : tcgen05.ld 32x32b.x32 bytes/clk/SM at 1..4 resident CTAs (4 warps each)
//   tmem_st   : tcgen05.st 32x32b.x16
//   ex2 f32 / ex2 bf16x2 / ex2 f16x2 : MUFU results/clk/SM
//   cvt bf16x2, fma-pipe polynomial exp2
//   mma       : cycles per tcgen05.mma for the shapes the attention kernels issue
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "tc_common.cuh"

using namespace pwa::tc;

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(1);                                                                     \
    }                                                                              \
  } while (0)

__device__ __forceinline__ uint32_t smid() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(r));
  return r;
}

// ---------------------------------------------------------------- TMEM ld / st
template <int MODE>  // 0 = ld x32, 1 = st x16, 2 = ld x32 + st x16 (softmax pattern)
__global__ void __launch_bounds__(128) tmem_kernel(long long* cyc, int iters, uint32_t* sink) {
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(&tmem_base_s, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t trow = tmem_base_s + ((uint32_t)(warp * 32) << 16);
  uint32_t acc = 0;
  uint32_t r[32];
  uint32_t pk[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) pk[i] = threadIdx.x + i;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (MODE == 0 || MODE == 2) {
        tmem_ld32(trow + c * 32, r);
        tmem_wait_ld();
        acc ^= r[0] ^ r[31];
      }
      if (MODE == 1 || MODE == 2) {
        tmem_st16(trow + c * 16, pk);
      }
    }
    if (MODE != 0) tmem_wait_st();
  }
  long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678u) sink[0] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base_s, 128);
}

// ---------------------------------------------------------------- MUFU / ALU
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t ex2bf2(uint32_t x) {
  uint32_t y;
  asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(x));
  return y;
}
__device__ __forceinline__ uint32_t ex2h2(uint32_t x) {
  uint32_t y;
  asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x));
  return y;
}
__device__ __forceinline__ uint32_t cvtbf2(float lo, float hi) {
  uint32_t y;
  asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(y) : "f"(hi), "f"(lo));
  return y;
}
// Cody-Waite + degree-3 polynomial exp2 on the FMA/ALU pipes (x <= 0)
__device__ __forceinline__ float exp2_poly(float x) {
  x = fmaxf(x, -126.f);
  const float fl = floorf(x);                 // FRND
  const float f = x - fl;                     // [0,1)
  float p = fmaf(f, 0.0555041086f, 0.2402265069f);
  p = fmaf(p, f, 0.6931471805f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + ((int)fl << 23));
}

template <int MODE>
__global__ void __launch_bounds__(256) alu_kernel(long long* cyc, int iters, float* sink, float seed) {
  constexpr int U = 8;
  float a[U];
  uint32_t b[U];
#pragma unroll
  for (int i = 0; i < U; ++i) {
    a[i] = -seed * (float)(threadIdx.x + i + 1) * 1e-3f;
    b[i] = 0xBF80BF80u + threadIdx.x + i;   // bf16 pair of ~-1
  }
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < U; ++i) {
      if (MODE == 0) a[i] = ex2f(a[i]) - 1.5f;                      // MUFU + FADD
      if (MODE == 1) b[i] = ex2bf2(b[i]) ^ 0x80008000u;             // MUFU(bf16x2) + LOP
      if (MODE == 2) b[i] = ex2h2(b[i]) ^ 0x80008000u;              // MUFU(f16x2) + LOP
      if (MODE == 3) b[i] = cvtbf2(a[i], __uint_as_float(b[i])) + 1;  // cvt pack
      if (MODE == 4) a[i] = exp2_poly(a[i]) - 1.5f;                 // polynomial exp2
      if (MODE == 5) {                                               // cvt pack + MUFU bf16x2 (the candidate inner loop)
        b[i] = ex2bf2(cvtbf2(a[i], a[(i + 1) % U]));
        a[i] = __uint_as_float(b[i]) * -1e-30f;
      }
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < U; ++i) s += a[i] + __uint_as_float(b[i]);
  if (s == 1.2345f) sink[0] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// ---------------------------------------------------------------- MMA shapes
// mode 0: SS M=128 N=n K=16 ; mode 1: TS (A from TMEM) M=128 N=n K=16, B MN-major
__global__ void __launch_bounds__(128) mma_kernel(long long* cyc, int n_mma, int N, int mode) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 32 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  fence_proxy_async_smem();
  __syncwarp();
  if (warp == 0) tmem_alloc(&tmem_base_s, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  long long t0 = 0, t1 = 0;
  if (tid == 0) {
    const uint32_t idesc = make_idesc_bf16(128, N, 0, (mode & 1) ? 1 : 0);
    const uint64_t da = make_smem_desc(smem_u32(smem), 128 * 16, 128);
    const uint64_t db = make_smem_desc(smem_u32(smem) + 8192, 256 * 16, 128);
    t0 = clock64();
    for (int i = 0; i < n_mma; ++i) {
      if (mode == 0) mma_ss(tmem, da, db, idesc, i > 0);
      else if (mode == 1) mma_ts(tmem + 64, tmem + (i & 7) * 8, db, idesc, i > 0);
      else if (mode == 2) mma_ss(tmem + (i & 3) * 64, da, db, idesc, i > 3);            // 4 independent accumulators (N <= 64)
      else mma_ts(tmem + 64 + (i & 3) * 48, tmem + (i & 7) * 8, db, idesc, i > 3);     // TS, 4 accumulators (N <= 48)
    }
    mma_commit(&bar);
  }
  __syncwarp();
  mbar_wait(&bar, 0);
  if (tid == 0) {
    t1 = clock64();
    cyc[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

// several independent MMA streams issued by different warps of ONE CTA (tmem_cols = 128 * nothing: alloc given)
__global__ void __launch_bounds__(128) mma_multi_kernel(long long* cyc, int n_mma, int N, int n_streams, int ts, int cols) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar[4];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 32 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1);
    fence_mbar_init();
  }
  fence_proxy_async_smem();
  __syncwarp();
  if (warp == 0) tmem_alloc(&tmem_base_s, cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  long long t0 = clock64();
  if ((tid & 31) == 0 && warp < n_streams) {
    const uint32_t idesc = make_idesc_bf16(128, N, 0, ts ? 1 : 0);
    const uint64_t da = make_smem_desc(smem_u32(smem), 128 * 16, 128);
    const uint64_t db = make_smem_desc(smem_u32(smem) + 8192, 256 * 16, 128);
    const uint32_t d = tmem + 32 + warp * (cols - 32) / 4;   // A (packed) at cols 0..31, accumulators after
    for (int i = 0; i < n_mma; ++i) {
      if (!ts) mma_ss(d, da, db, idesc, i > 0);
      else mma_ts(d, tmem + (i & 3) * 8, db, idesc, i > 0);
    }
    mma_commit(&bar[warp]);
  }
  __syncwarp();
  if (warp < n_streams) mbar_wait(&bar[warp], 0);
  __syncthreads();
  if (tid == 0) cyc[blockIdx.x] = clock64() - t0;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, cols);
}

static void report(const char* name, const std::vector<long long>& c, double units_per_cta, int ctas_per_sm) {
  long long mx = 0;
  double avg = 0;
  for (auto v : c) { mx = v > mx ? v : mx; avg += (double)v; }
  avg /= c.size();
  printf("%-34s ctas/SM %d  cycles avg %.0f max %lld  -> %.2f units/clk/SM\n", name, ctas_per_sm, avg, mx,
         units_per_cta * ctas_per_sm / avg);
}

int main(int argc, char** argv) {
  const bool only_mma = argc > 1;
  long long* dc;
  uint32_t* dsink;
  CK(cudaMalloc(&dc, 148 * 16 * sizeof(long long)));
  CK(cudaMalloc(&dsink, 64));
  std::vector<long long> hc;
  auto fetch = [&](int n) {
    hc.resize(n);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(hc.data(), dc, n * sizeof(long long), cudaMemcpyDeviceToHost));
  };
  const int iters = 2000;
  for (int occ = 1; occ <= 4 && !only_mma; ++occ) {
    const int grid = 148 * occ;
    tmem_kernel<0><<<grid, 128>>>(dc, iters, dsink);
    fetch(grid);
    report("tmem_ld 32x32b.x32 (bytes)", hc, (double)iters * 4 * 4 * 4096, occ);
    tmem_kernel<1><<<grid, 128>>>(dc, iters, dsink);
    fetch(grid);
    report("tmem_st 32x32b.x16 (bytes)", hc, (double)iters * 4 * 4 * 2048, occ);
    tmem_kernel<2><<<grid, 128>>>(dc, iters, dsink);
    fetch(grid);
    report("tmem ld x32 + st x16 (ld bytes)", hc, (double)iters * 4 * 4 * 4096, occ);
  }
  const char* names[6] = {"ex2.f32 + FADD (results)", "ex2.bf16x2 + LOP (results)", "ex2.f16x2 + LOP (results)",
                          "cvt.bf16x2 + IADD (results)", "poly exp2 (results)", "cvt+ex2.bf16x2+FMUL (results)"};
  for (int occ = 1; occ <= 8 && !only_mma; occ *= 2) {
    const int grid = 148 * occ;
    const int it2 = 4000;
    for (int mode = 0; mode < 6; ++mode) {
      switch (mode) {
        case 0: alu_kernel<0><<<grid, 256>>>(dc, it2, (float*)dsink, 1.f); break;
        case 1: alu_kernel<1><<<grid, 256>>>(dc, it2, (float*)dsink, 1.f); break;
        case 2: alu_kernel<2><<<grid, 256>>>(dc, it2, (float*)dsink, 1.f); break;
        case 3: alu_kernel<3><<<grid, 256>>>(dc, it2, (float*)dsink, 1.f); break;
        case 4: alu_kernel<4><<<grid, 256>>>(dc, it2, (float*)dsink, 1.f); break;
        case 5: alu_kernel<5><<<grid, 256>>>(dc, it2, (float*)dsink, 1.f); break;
      }
      fetch(grid);
      const double per_thread = (double)it2 * 8 * ((mode == 1 || mode == 2 || mode == 3 || mode == 5) ? 2 : 1);
      report(names[mode], hc, per_thread * 256, occ);
    }
  }
  CK(cudaFuncSetAttribute(mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 1024));
  const int shapes[][2] = {{0, 128}, {0, 16}, {1, 16}, {2, 16}, {2, 64}, {3, 16}, {3, 48}};
  for (auto& s : shapes) {
    for (int occ = 1; occ <= 4; occ *= 2) {
      const int n_mma = 512;
      mma_kernel<<<148 * occ, 128, 32 * 1024>>>(dc, n_mma, s[1], s[0]);
      fetch(148 * occ);
      long long mx = 0;
      double avg = 0;
      for (auto v : hc) { mx = v > mx ? v : mx; avg += (double)v; }
      avg /= hc.size();
      printf("mma %s M=128 N=%3d K=16: ctas/SM %d  %.1f clk per MMA (avg), %.1f (max)  -> %.0f FLOP/clk/SM\n",
             (s[0] == 0 ? "SS" : s[0] == 1 ? "TS" : s[0] == 2 ? "SS4acc" : "TS4acc"), s[1], occ, avg / n_mma, (double)mx / n_mma, 2.0 * 128 * s[1] * 16 * n_mma * occ / avg);
    }
  }
  CK(cudaFuncSetAttribute(mma_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 1024));
  for (int ts = 0; ts < 2; ++ts)
    for (int cols = 128; cols <= 256; cols *= 2)
      for (int ns = 1; ns <= 4; ns *= 2)
        for (int occ = 1; occ <= 512 / cols; occ *= 2) {
          const int n_mma = 512, N = 16;
          mma_multi_kernel<<<148 * occ, 128, 32 * 1024>>>(dc, n_mma, N, ns, ts, cols);
          fetch(148 * occ);
          double avg = 0;
          for (auto v : hc) avg += (double)v;
          avg /= hc.size();
          printf("multi %s N=16 tmem_cols %d: ctas/SM %d streams/CTA %d -> %.1f clk per MMA per stream, %.1f clk per MMA per SM\n",
                 ts ? "TS" : "SS", cols, occ, ns, avg / n_mma, avg / n_mma / (ns * occ));
        }
  return 0;
}
