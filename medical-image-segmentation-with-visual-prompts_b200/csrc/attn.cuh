// Parameter block shared by the attention kernels and the C-ABI dispatcher.
#pragma once
#include "common.cuh"

namespace pwa {

struct AttnParams {
  const void *q, *k, *v, *kp, *vp;
  const float *th, *tw, *td, *tok;
  const uint8_t* ids;
  void* out;
  float* lse;
  const void* dout;
  void *dq, *dk, *dv;
  float *dkp, *dvp, *dth, *dtw, *dtd, *dtok, *delta;
  int B, P, C, heads, I, N, NK;
  int ldq, ldp;      // row strides (elements) of q/k/v (+ dq/dk/dv) and of kp/vp; out/dout/dkp/dvp use C
  int wh, ww, wd;
  float scale;
  int debug;         // PWA_TIMELINE=1: CTA 0 writes clock64 stamps into the (otherwise unused) delta buffer
};

// fp32-math CUDA-core kernels (attn_f32.cu)
int attn_f32_forward(const AttnParams& p, int dtype, cudaStream_t st);
int attn_f32_backward(const AttnParams& p, int dtype, cudaStream_t st);

// bf16 tcgen05 / TMEM kernels (attn_tc.cu)
bool attn_tc_supported(const AttnParams& p, int dtype);
int attn_tc_forward(const AttnParams& p, cudaStream_t st);
bool attn_tc_bwd_supported(const AttnParams& p, int dtype);   // attn_tc_bwd.cu
int attn_tc_backward(const AttnParams& p, cudaStream_t st);

}  // namespace pwa
