// (b) fused prompted window attention forward on the 5th-gen tensor cores (tcgen05 + TMEM), bf16 I/O.
//
// One CTA = 128 threads = one fixed head; it walks over (sample, window) pairs.  Per window the CTA stages
// this head's Q / K / V slices in shared memory in UMMA canonical (no-swizzle) layouts, then for each of the
// two 128-row query tiles and each key block (content 0-127, content 128-255, prompt 0-I):
//     S[128 x nk]  = Q'.K'^T            tcgen05.mma kind::f16, both operands from smem, fp32 accum in TMEM
//     P            = exp2(mask(S) - m)  each thread owns one query row = one TMEM lane (tcgen05.ld 32x32b),
//                                       writes P back as packed bf16 over the consumed S columns (tcgen05.st)
//     O[128 x dh] += P.V                tcgen05.mma with A = P straight from TMEM, B = V (MN-major) from smem
// and merges the key blocks with a running (max, sum, O) per row in registers.
//
// Relative-position bias is folded INTO the QK^T MMA: bias[n][m] = Th[ih][jh] + Tw[iw][jw] + Td[id][jd] is a
// sum of three one-hot x table products, so Q' = [q | onehot_d(n) | 0 ;; onehot_h(n) | onehot_w(n)] and
// K' = [k | Td[.][jd]/scale | 0 ;; Th[.][jh]/scale | Tw[.][jw]/scale] give S = q.k + bias/scale.  The tensor
// pipe has >8x headroom at head_dim 12 (the kernel is bound by MUFU/ALU softmax work), so the extra k-step is
// free while it removes two FADDs + table lookups per logit from the CUDA cores.  Prompt keys use
// [tok/scale x wh | 0] so every query row picks up tok[i].  The one-hot / table halves do not depend on the
// window, so they are built once per CTA and stay resident in smem, as do the prompt-independent constants.
//
// The shift mask is multiplicative and applied BEFORE softmax (window_attention.py:54-56): masked logits
// become exactly 0 and still get weight exp(0 - max).  It is evaluated from uint8 region ids in smem.
//
// Four CTAs are resident per SM (128 TMEM columns and ~46 KB smem each at head_dim 12): while one CTA waits
// for its MMAs or stages the next window, the others keep the MUFU/ALU pipes busy, so no intra-CTA
// warp-specialised pipeline is needed.
#include "attn.cuh"
#include "tc_common.cuh"

namespace pwa {
using namespace tc;

namespace {

constexpr int kRows = 128;        // query rows per tile = TMEM lanes = threads per CTA
constexpr int kN = 256;           // content tokens per window (two query tiles, two content key blocks)
constexpr int kTmemCols = 128;
constexpr int kOCol = 64;         // O accumulator lives in columns [64, 64 + DHP) of the S region

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int DH> struct Cfg {
  static constexpr int DHP = (DH + 15) / 16 * 16;          // V / O width (PV MMA N)
  static constexpr int NDC = DHP / 8;                      // 16-byte chunks per V row
};

struct TcSmem {
  uint32_t q, k, v, qaug, kaug, ids, total;                // byte offsets
};

__host__ __device__ inline TcSmem tc_layout(int KS, int DHP, int NKT) {
  TcSmem s;
  uint32_t o = 0;
  s.q = o; o += KS * 2 * kN * 16;
  s.k = o; o += KS * 2 * NKT * 16;
  s.v = o; o += NKT * DHP * 2;
  s.qaug = o; o += 2 * kN * 16;
  s.kaug = o; o += 2 * NKT * 16;
  s.ids = o; o += kN;
  s.total = o;
  return s;
}

template <int DH>
__device__ __forceinline__ void load_row(const __nv_bfloat16* src, __nv_bfloat16 (&dst)[DH]) {
  if constexpr (DH % 4 == 0) {
    const uint2* s2 = reinterpret_cast<const uint2*>(src);
    uint2* d2 = reinterpret_cast<uint2*>(dst);
#pragma unroll
    for (int i = 0; i < DH / 4; ++i) d2[i] = __ldg(s2 + i);
  } else {
#pragma unroll
    for (int i = 0; i < DH; ++i) dst[i] = src[i];
  }
}

// staged K-dim layout of one head: [real DH | wd extra columns | zero pad] -> KS k-steps of 16
template <int DH, int KS>
__device__ __forceinline__ void store_chunks(uint8_t* base, uint32_t chunk_stride, int row, const __nv_bfloat16 (&real)[DH],
                                             const __nv_bfloat16* extra, int n_extra) {
#pragma unroll
  for (int c = 0; c < KS * 2; ++c) {
    __align__(16) __nv_bfloat16 tmp[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int col = c * 8 + e;
      const int x = col - DH;
      __nv_bfloat16 v = __float2bfloat16(0.f);
      if (col < DH) v = real[col < DH ? col : 0];
      else if (x < 4 && x < n_extra) v = extra[x & 3];
      tmp[e] = v;
    }
    *reinterpret_cast<uint4*>(base + c * chunk_stride + row * 16) = *reinterpret_cast<const uint4*>(tmp);
  }
}

template <int DH, bool MASKED>
__global__ void __launch_bounds__(kRows, 4) attn_fwd_tc_kernel(AttnParams p) {
  constexpr int DHP = Cfg<DH>::DHP, NDC = Cfg<DH>::NDC;
  constexpr int KS = (DH + 4 + 15) / 16;                   // wd <= 4 extra columns ride in the padding
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5;
  const int NKT = kN + p.I;
  const TcSmem L = tc_layout(KS, DHP, NKT);
  uint8_t* Qs = smem + L.q;
  uint8_t* Ks = smem + L.k;
  uint8_t* Vs = smem + L.v;
  uint8_t* Qa = smem + L.qaug;
  uint8_t* Ka = smem + L.kaug;
  uint8_t* ids_s = smem + L.ids;
  const int head = blockIdx.x % p.heads;
  const float inv_scale = 1.f / p.scale;
  const float c2 = p.scale * 1.4426950408889634f;          // logits -> log2 domain
  const __nv_bfloat16 one = __float2bfloat16(1.f), zero = __float2bfloat16(0.f);

  // ---- once per CTA: window-independent halves of Q' and K' ----
  for (int n = tid; n < kN; n += kRows) {
    const int iw = (n / p.wd) % p.ww, ih = n / (p.wd * p.ww);
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      __align__(16) __nv_bfloat16 tmp[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int col = c * 8 + e;
        tmp[e] = (col < p.wh) ? (col == ih ? one : zero) : ((col - p.wh < p.ww && col - p.wh == iw) ? one : zero);
      }
      *reinterpret_cast<uint4*>(Qa + c * (kN * 16) + n * 16) = *reinterpret_cast<const uint4*>(tmp);
    }
  }
  for (int j = tid; j < NKT; j += kRows) {
    const bool content = j < kN;
    const int jw = (j / p.wd) % p.ww, jh = j / (p.wd * p.ww);
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      __align__(16) __nv_bfloat16 tmp[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int col = c * 8 + e;
        float v = 0.f;
        if (content) {
          if (col < p.wh) v = p.th[(head * p.wh + col) * p.wh + jh];
          else if (col - p.wh < p.ww) v = p.tw[(head * p.ww + (col - p.wh)) * p.ww + jw];
        } else if (col < p.wh) {
          v = p.tok[head * p.I + (j - kN)];
        }
        tmp[e] = __float2bfloat16(v * inv_scale);
      }
      *reinterpret_cast<uint4*>(Ka + c * (NKT * 16) + j * 16) = *reinterpret_cast<const uint4*>(tmp);
    }
  }
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  __syncwarp();
  if (warp == 0) tmem_alloc(&tmem_base_s, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);   // this warp's 32 lanes
  uint32_t phase = 0;

  const uint32_t idescS128 = make_idesc_bf16(128, 128, 0, 0);
  const uint32_t idescSP = make_idesc_bf16(128, p.I > 0 ? p.I : 16, 0, 0);
  const uint32_t idescPV = make_idesc_bf16(128, DHP, 0, 1);
  const int n_kb = p.I > 0 ? 3 : 2;
  const int n_pairs = p.B * p.P;
  const int stride = gridDim.x / p.heads;

  for (int bw = blockIdx.x / p.heads; bw < n_pairs; bw += stride) {
    const int b = bw / p.P, win = bw - b * p.P;
    // ---- stage this (window, head): Q, K (content + prompt rows), V, region ids ----
    __nv_bfloat16 extra[4];
    for (int n = tid; n < kN; n += kRows) {
      __nv_bfloat16 row[DH];
      load_row<DH>((const __nv_bfloat16*)p.q + ((size_t)bw * kN + n) * p.ldq + head * DH, row);
      const int id_ = n % p.wd;
#pragma unroll
      for (int u = 0; u < 4; ++u) extra[u] = (u == id_) ? one : zero;
      store_chunks<DH, KS>(Qs, kN * 16, n, row, extra, p.wd);
    }
    for (int j = tid; j < NKT; j += kRows) {
      const bool content = j < kN;
      const size_t off = content ? ((size_t)bw * kN + j) * p.ldq + head * DH : ((size_t)b * p.I + (j - kN)) * p.ldp + head * DH;
      __nv_bfloat16 row[DH];
      load_row<DH>((const __nv_bfloat16*)(content ? p.k : p.kp) + off, row);
      const int jd = j % p.wd;
#pragma unroll
      for (int u = 0; u < 4; ++u)
        extra[u] = (content && u < p.wd) ? __float2bfloat16(p.td[(head * p.wd + u) * p.wd + jd] * inv_scale) : zero;
      store_chunks<DH, KS>(Ks, NKT * 16, j, row, extra, p.wd);
      load_row<DH>((const __nv_bfloat16*)(content ? p.v : p.vp) + off, row);
#pragma unroll
      for (int dc = 0; dc < NDC; ++dc) {
        __align__(16) __nv_bfloat16 tmp[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) tmp[e] = (dc * 8 + e < DH) ? row[dc * 8 + e] : zero;
        *reinterpret_cast<uint4*>(Vs + (j >> 3) * (NDC * 128) + dc * 128 + (j & 7) * 16) =
            *reinterpret_cast<const uint4*>(tmp);
      }
    }
    if (MASKED)
      for (int i = tid; i < kN / 4; i += kRows)
        reinterpret_cast<uint32_t*>(ids_s)[i] = reinterpret_cast<const uint32_t*>(p.ids + (size_t)win * kN)[i];
    fence_proxy_async_smem();
    __syncthreads();

    for (int mt = 0; mt < 2; ++mt) {
      const int rown = mt * kRows + tid;
      const uint32_t rid = MASKED ? ids_s[rown] : 0;
      float m_run = -1e30f, l_run = 0.f;
      float o_run[DHP];
#pragma unroll
      for (int d = 0; d < DHP; ++d) o_run[d] = 0.f;

      for (int kb = 0; kb < n_kb; ++kb) {
        const int nk = kb < 2 ? 128 : p.I;
        // ---- S = Q'.K'^T ----
        if (tid == 0) {
          tc_fence_after();
          const uint32_t idesc = kb < 2 ? idescS128 : idescSP;
#pragma unroll
          for (int ks = 0; ks < KS; ++ks) {
            const uint64_t da = make_smem_desc(smem_u32(Qs) + ks * 2 * (kN * 16) + mt * (kRows * 16), kN * 16, 128);
            const uint64_t db = make_smem_desc(smem_u32(Ks) + ks * 2 * (NKT * 16) + kb * (128 * 16), NKT * 16, 128);
            mma_ss(tmem, da, db, idesc, ks > 0);
          }
          const uint64_t da = make_smem_desc(smem_u32(Qa) + mt * (kRows * 16), kN * 16, 128);
          const uint64_t db = make_smem_desc(smem_u32(Ka) + kb * (128 * 16), NKT * 16, 128);
          mma_ss(tmem, da, db, idesc, 1);
          mma_commit(&bar);
        }
        __syncwarp();
        mbar_wait(&bar, phase);
        phase ^= 1;
        tc_fence_after();

        // ---- pass 1: row max of the masked logits ----
        const bool do_mask = MASKED && kb < 2;
        float mx = -1e30f;
        for (int c = 0; c < nk / 32; ++c) {
          uint32_t r[32];
          tmem_ld32(trow + c * 32, r);
          tmem_wait_ld();
          const uint32_t* idw = reinterpret_cast<const uint32_t*>(ids_s + kb * 128 + c * 32);
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const uint32_t w = do_mask ? idw[g] : 0;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float s = __uint_as_float(r[g * 4 + e]);
              if (do_mask && ((w >> (8 * e)) & 0xffu) != rid) s = 0.f;
              mx = fmaxf(mx, s);
            }
          }
        }
        const float m_new = fmaxf(m_run, mx);
        const float mb = m_new * c2;
        const float e0 = fast_exp2(-mb);                         // weight of every masked (zeroed) logit
        // ---- pass 2: P = exp2(logit - max), packed bf16 back into TMEM ----
        float sum = 0.f;
        for (int c = 0; c < nk / 32; ++c) {
          uint32_t r[32];
          tmem_ld32(trow + c * 32, r);
          tmem_wait_ld();
          const uint32_t* idw = reinterpret_cast<const uint32_t*>(ids_s + kb * 128 + c * 32);
          uint32_t pk[16];
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const uint32_t w = do_mask ? idw[g] : 0;
            float pv[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float s = __uint_as_float(r[g * 4 + e]);
              float pe = fast_exp2(fmaf(s, c2, -mb));
              if (do_mask && ((w >> (8 * e)) & 0xffu) != rid) pe = e0;
              pv[e] = pe;
              sum += pe;
            }
            pk[g * 2] = pack_bf16(pv[0], pv[1]);
            pk[g * 2 + 1] = pack_bf16(pv[2], pv[3]);
          }
          tmem_st16(trow + c * 16, pk);
        }
        const float alpha = fast_exp2((m_run - m_new) * c2);
        l_run = l_run * alpha + sum;
        m_run = m_new;
        tmem_wait_st();
        tc_fence_before();
        __syncthreads();

        // ---- O_blk = P.V ----
        if (tid == 0) {
          tc_fence_after();
          for (int t = 0; t < nk / 16; ++t) {
            const uint64_t dv = make_smem_desc(smem_u32(Vs) + ((kb * 128 + t * 16) >> 3) * (NDC * 128), NDC * 128, 128);
            mma_ts(tmem + kOCol, tmem + t * 8, dv, idescPV, t > 0);
          }
          mma_commit(&bar);
        }
        __syncwarp();
        mbar_wait(&bar, phase);
        phase ^= 1;
        tc_fence_after();
#pragma unroll
        for (int dq = 0; dq < DHP / 16; ++dq) {
          uint32_t o[16];
          tmem_ld16(trow + kOCol + dq * 16, o);
          tmem_wait_ld();
#pragma unroll
          for (int d = 0; d < 16; ++d) o_run[dq * 16 + d] = fmaf(o_run[dq * 16 + d], alpha, __uint_as_float(o[d]));
        }
        tc_fence_before();
        __syncthreads();   // everyone has drained O / P before the next S MMA overwrites the columns
      }

      // ---- epilogue: normalise, write bf16 output row slice and log-sum-exp ----
      const float inv = 1.f / l_run;
      __nv_bfloat16* og = (__nv_bfloat16*)p.out + ((size_t)bw * kN + rown) * p.C + head * DH;
      if constexpr (DH % 4 == 0) {
#pragma unroll
        for (int d = 0; d < DH; d += 4) {
          uint2 v;
          v.x = pack_bf16(o_run[d] * inv, o_run[d + 1] * inv);
          v.y = pack_bf16(o_run[d + 2] * inv, o_run[d + 3] * inv);
          *reinterpret_cast<uint2*>(og + d) = v;
        }
      } else {
#pragma unroll
        for (int d = 0; d < DH; ++d) og[d] = __float2bfloat16(o_run[d] * inv);
      }
      p.lse[((size_t)bw * p.heads + head) * kN + rown] = m_run * p.scale + __logf(l_run);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, kTmemCols);
}

template <int DH>
int launch_tc(const AttnParams& p, cudaStream_t st) {
  constexpr int DHP = Cfg<DH>::DHP;
  constexpr int KS = (DH + 4 + 15) / 16;
  const int NKT = kN + p.I;
  const TcSmem L = tc_layout(KS, DHP, NKT);
  const size_t smem = L.total;
  int grid = 148 * 4;
  grid -= grid % p.heads;
  const int need = p.B * p.P * p.heads;
  if (grid > need) grid = need;
  if (grid < p.heads) grid = p.heads;
  auto kern = p.ids ? attn_fwd_tc_kernel<DH, true> : attn_fwd_tc_kernel<DH, false>;
  PWA_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, kRows, smem, st>>>(p);
  PWA_CUDA_OK(cudaGetLastError());
  return PWA_OK;
}

}  // namespace

bool attn_tc_supported(const AttnParams& p, int dtype) {
  if (dtype != PWA_BF16) return false;
  if (p.N != kN) return false;                                   // two 128-row tiles / two 128-key content blocks
  if (p.wd > 4 || p.wh + p.ww > 16) return false;                // one-hot bias columns must fit the layout
  if (p.I % 32 != 0 || p.I > 128) return false;                  // prompt block = one MMA of N = I, read in 32-column chunks
  const int dh = p.C / p.heads;
  if (!(dh == 12 || dh == 24 || dh == 48 || dh == 6 || dh == 3)) return false;
  const TcSmem L = tc_layout((dh + 4 + 15) / 16, (dh + 15) / 16 * 16, kN + p.I);
  return L.total <= 200 * 1024;
}

int attn_tc_forward(const AttnParams& p, cudaStream_t st) {
  switch (p.C / p.heads) {
    case 3: return launch_tc<3>(p, st);
    case 6: return launch_tc<6>(p, st);
    case 12: return launch_tc<12>(p, st);
    case 24: return launch_tc<24>(p, st);
    case 48: return launch_tc<48>(p, st);
  }
  set_error("tcgen05 attention: head_dim %d not instantiated", p.C / p.heads);
  return PWA_ERR_UNSUPPORTED;
}

}  // namespace pwa
