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
  // attention dropout (window_attention.py:57): 0 = off.  An element (query n, key j) is dropped iff its byte of a
  // counter-based hash is < drop_thresh (probability drop_thresh / 256); kept probabilities are scaled by inv_keep.
  uint32_t drop_thresh;
  float inv_keep;
  const uint32_t* drop_seed;   // DEVICE pointer to two 32-bit seed words (graph-replay safe), or null
  uint32_t seed_host[2];       // used when drop_seed is null
  unsigned int* work;          // forward: per-head window counters (zeroed by the launcher), or null = static round-robin
  int debug;         // PWA_TIMELINE=1: CTA 0 writes clock64 stamps into the (otherwise unused) delta buffer
};

// ---- dropout mask: 32 hash bits per 2x2 block (query pair, key pair), one byte per element, so that a thread that
// walks along keys (forward, dQ) and one that walks along queries (dK/dV) both amortise one hash over two elements
__device__ __forceinline__ uint32_t drop_mix(uint32_t x) {
  x ^= x >> 16; x *= 0x21f0aaadu; x ^= x >> 15; x *= 0x735a2d97u; x ^= x >> 15;
  return x;
}
// state of query pair (n >> 1) of (sample*window, head); NH = number of query pairs per window
__device__ __forceinline__ uint32_t drop_row_state(uint32_t s0, uint32_t s1, uint32_t bw, uint32_t heads, uint32_t head, uint32_t NH,
                                                  uint32_t n) {
  return drop_mix(s0 + ((bw * heads + head) * NH + (n >> 1)) * 0x85EBCA77u) ^ s1;
}
__device__ __forceinline__ uint32_t drop_block_bits(uint32_t row_state, uint32_t j) {
  return drop_mix(row_state ^ ((j >> 1) * 0x9E3779B1u));
}
__device__ __forceinline__ bool drop_keep(uint32_t bits, uint32_t n, uint32_t j, uint32_t thresh) {
  return ((bits >> (8u * ((n & 1u) * 2u + (j & 1u)))) & 0xffu) >= thresh;
}

// fp32-math CUDA-core kernels (attn_f32.cu)
int attn_f32_forward(const AttnParams& p, int dtype, cudaStream_t st);
int attn_f32_backward(const AttnParams& p, int dtype, cudaStream_t st);

// bf16 tcgen05 / TMEM kernels (attn_tc.cu)
bool attn_tc_supported(const AttnParams& p, int dtype);
int attn_tc_forward(const AttnParams& p, cudaStream_t st);
bool attn_tc_bwd_supported(const AttnParams& p, int dtype);   // attn_tc_bwd.cu
int attn_tc_backward(const AttnParams& p, cudaStream_t st);

}  // namespace pwa
