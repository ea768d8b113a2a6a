// (c) backward of the fused prompted window attention on tcgen05 tensor cores + TMEM, bf16 I/O.
//
// One CTA = 256 threads = one fixed head, walking over (sample, window) pairs; two CTAs per SM.
// Everything is computed in the TRANSPOSED orientation: the 128 TMEM lanes are KEYS (one key block: content
// 0-127, content 128-255, prompt tokens) and the TMEM columns are query rows, 64 at a time:
//     S^T [128k x 64r] = K'.Q'^T      dP^T [128k x 64r] = V.dO^T          (SS MMAs, fp32 accum in TMEM)
//     P^T = exp2(mask(S^T)*c - lse)   g^T = mask * P^T * (dP^T - delta)    (one thread per key, 32 rows each;
//                                                                           the two warpgroups split the columns)
//     dV  [128k x dh] += P^T.dO       dK' [128k x dh'] += g^T.Q'           (TS MMAs: A = bf16 P^T / g^T in TMEM)
//     dKaug[128k x 16] += g^T.[onehot_h | onehot_w]                        (= relative-position-bias table
//                                                                           gradients, accumulated in TMEM over
//                                                                           ALL windows the CTA processes)
//     dQ' [128r x dh'] += g.K'                                             (A = g^T staged to smem as an MN-major
//                                                                           operand, B = K' MN-major)
// Because lse and delta = rowsum(dO*O) are known, no row-wise reduction is needed and the exponentials are
// evaluated exactly once per (query, key) pair -- the same MUFU work as the forward.  Per logit the CUDA cores
// issue FFMA + MUFU + FMUL + 2 x 1/2 F2FP: `- delta` rides in two spare K columns of the dP^T MMA (bf16 hi/lo
// split, V' columns = 1) and the multiplicative shift mask is applied on the PACKED bf16 pairs with one PRMT
// each (masked P^T -> exp(-lse) of that query, masked g^T -> 0; selector table as in attn_tc.cu).
// tcgen05.mma costs ~100 clk of latency per instruction, but streams issued by different warps overlap
// (csrc/ubench.cu): the dV, dK', dKaug and dQ' chains of a unit are issued by four different warps, a fifth
// issues the next unit's S^T / dP^T as soon as the chains that read the packed P^T / g^T columns have retired.
// One shared-memory copy of each operand serves both roles it plays: the [chunk][row][16 B] layout is the
// canonical no-swizzle K-major layout of a [rows x dh] operand AND the MN-major layout of its transpose.
// Prompt-token dK/dV are reduced over windows with fp32 atomics; bias-table gradients leave the kernel once
// per CTA.  Semantics follow the reference autograd of window_attention.py:49-58 (mask multiplicative,
// pre-softmax: masked logits are 0, keep weight exp(-lse), and pass no gradient to q.k or the bias).
#include "attn.cuh"
#include "tc_common.cuh"

namespace pwa {
using namespace tc;

namespace {

constexpr int kN = 256;
constexpr int kThreadsB = 256;
constexpr int kIds = 28;          // region ids 0..26 and 100 (-> 27), see pwa_region_ids

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
  return d;
}
__device__ __forceinline__ int id_slot(uint32_t id) { return id < (uint32_t)(kIds - 1) ? (int)id : kIds - 1; }

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct BwdSmem {
  uint32_t q, k, v, dO, qaug, kaug, g, lse2, delta, wp, sel, ids, gth, gtw, gtd, gtok, total;
};

__host__ __device__ inline BwdSmem bwd_layout(int KS, int DHP, int NKT, int wh, int ww, int wd, int I, bool masked) {
  BwdSmem s;
  uint32_t o = 0;
  s.q = o; o += KS * 2 * kN * 16;
  s.k = o; o += KS * 2 * (kN + 128) * 16;       // key-side operands always hold 3 x 128 rows: the prompt block is issued as M = 128
  s.v = o; o += (DHP / 8) * (kN + 128) * 16;
  s.dO = o; o += (DHP / 8) * kN * 16;
  s.qaug = o; o += 2 * kN * 16;
  s.kaug = o; o += 2 * (kN + 128) * 16;
  s.g = o; o += 128 * 128 * 2;
  s.lse2 = o; o += kN * 4;
  s.delta = o; o += kN * 4;
  s.wp = o; o += kN * 2;                         // bf16 exp(-lse) per query (value of a masked P entry)
  s.sel = o; o += masked ? kIds * (kN / 4) * 4 : 0;   // PRMT selectors [id slot][4 tokens], as in attn_tc.cu
  s.ids = o; o += kN;
  s.gth = o; o += wh * wh * 4;
  s.gtw = o; o += ww * ww * 4;
  s.gtd = o; o += wd * wd * 4;
  s.gtok = o; o += (I + 4) * 4;
  s.total = (o + 15) & ~15u;
  return s;
}

template <int DH>
__device__ __forceinline__ void load_row_b(const __nv_bfloat16* src, __nv_bfloat16 (&dst)[DH]) {
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

// [real DH | n_extra extra columns | zero pad] -> NCH chunks of 8 columns, chunk c of row r at base + c*stride + r*16
template <int DH, int NCH>
__device__ __forceinline__ void store_chunks_b(uint8_t* base, uint32_t chunk_stride, int row, const __nv_bfloat16 (&real)[DH],
                                               const __nv_bfloat16* extra, int n_extra) {
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
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
__global__ void __launch_bounds__(kThreadsB, (DH <= 12 ? 2 : 1)) attn_bwd_tc_kernel(AttnParams p, uint32_t tmem_cols) {
  constexpr int DHP = (DH + 15) / 16 * 16;
  constexpr int KS = (DH + 4 + 15) / 16;
  constexpr int DKC = KS * 16;                     // staged K' / Q' width = dK' / dQ' accumulator width
  constexpr bool FOLD = (DHP - DH) >= 2;           // -delta rides in two spare K columns of the dP^T MMA
  // TMEM column map
  constexpr uint32_t cST = 0, cDPT = 64, cDV = 128, cDK = cDV + DHP, cDQ = cDK + DKC, cAUG = cDQ + 2 * DKC;
  // mbarriers: S = scores of a unit ready; V / K / A / Q = dV / dK' / dKaug / dQ' chain of a unit retired
  enum { bS = 0, bV = 1, bK = 2, bA = 3, bQ = 4 };

  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar[5];
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, wg = tid >> 7, lane = tid & 31;
  const int lane_row = tid & 127;                  // TMEM lane owned by this thread (key in S^T, row in dQ)
  const int NKT = kN + p.I;
  const int NKR = kN + 128;                        // rows allocated for key-side operands (prompt block issued as M = 128)
  const BwdSmem L = bwd_layout(KS, DHP, NKT, p.wh, p.ww, p.wd, p.I, MASKED);
  uint8_t* Qs = smem + L.q;
  uint8_t* Ks = smem + L.k;
  uint8_t* Vs = smem + L.v;
  uint8_t* dOs = smem + L.dO;
  uint8_t* Qa = smem + L.qaug;
  uint8_t* Ka = smem + L.kaug;
  uint8_t* Gs = smem + L.g;
  float* lse2_s = reinterpret_cast<float*>(smem + L.lse2);
  float* delta_s = reinterpret_cast<float*>(smem + L.delta);
  __nv_bfloat16* wp_s = reinterpret_cast<__nv_bfloat16*>(smem + L.wp);
  uint32_t* sel_s = reinterpret_cast<uint32_t*>(smem + L.sel);
  uint8_t* ids_s = smem + L.ids;
  float* gth_s = reinterpret_cast<float*>(smem + L.gth);
  float* gtw_s = reinterpret_cast<float*>(smem + L.gtw);
  float* gtd_s = reinterpret_cast<float*>(smem + L.gtd);
  float* gtok_s = reinterpret_cast<float*>(smem + L.gtok);

  const int head = blockIdx.x % p.heads;
  const float inv_scale = 1.f / p.scale;
  const float c2 = p.scale * 1.4426950408889634f;
  const __nv_bfloat16 one = __float2bfloat16(1.f), zero = __float2bfloat16(0.f);

  // ---- once per CTA: zero everything the MMAs may touch beyond the staged rows, window-independent operands ----
  for (uint32_t i = tid; i < L.total / 16; i += kThreadsB) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  for (int n = tid; n < kN; n += kThreadsB) {
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
  for (int j = tid; j < NKT; j += kThreadsB) {
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
      *reinterpret_cast<uint4*>(Ka + c * (NKR * 16) + j * 16) = *reinterpret_cast<const uint4*>(tmp);
    }
  }
  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < 5; ++i) mbar_init(&bar[i], 1);
    fence_mbar_init();
  }
  __syncwarp();
  if (warp == 0) tmem_alloc(&tmem_base_s, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t phS = 0, phC = 0, phQ = 0;                           // parities: scores, the three per-unit chains, dQ'
  bool chains_pending = false;                                  // (issuer thread) chains of the previous unit not yet awaited

  const uint32_t idescT = make_idesc_bf16(128, 64, 0, 0);       // S^T, dP^T : A K-major, B K-major, N = 64 rows
  const uint32_t idescDV = make_idesc_bf16(128, DHP, 0, 1);     // dV  : A tmem, B = dO MN-major
  const uint32_t idescDK = make_idesc_bf16(128, DKC, 0, 1);     // dK' : A tmem, B = Q' MN-major
  const uint32_t idescAUG = make_idesc_bf16(128, 16, 0, 1);     // dKaug
  const uint32_t idescDQ = make_idesc_bf16(128, DKC, 1, 1);     // dQ' : A = g smem MN-major, B = K' MN-major
  const int n_kb = p.I > 0 ? 3 : 2;
  const int n_units = n_kb * 4;
  const int n_pairs = p.B * p.P;
  const int stride = gridDim.x / p.heads;
  float acc_d[2][4];                                            // dTd contributions of this thread's keys
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int u = 0; u < 4; ++u) acc_d[a][u] = 0.f;
  bool first_window = true;

  // issue S^T and dP^T for unit (kb, mt, hf)
  auto issue_scores = [&](int kb, int mt, int hf) {
    const uint32_t qrow = (uint32_t)(mt * 128 + hf * 64);
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      const uint64_t da = make_smem_desc(smem_u32(Ks) + ks * 2 * (NKR * 16) + kb * (128 * 16), NKR * 16, 128);
      const uint64_t db = make_smem_desc(smem_u32(Qs) + ks * 2 * (kN * 16) + qrow * 16, kN * 16, 128);
      mma_ss(tmem + cST, da, db, idescT, ks > 0);
    }
    {
      const uint64_t da = make_smem_desc(smem_u32(Ka) + kb * (128 * 16), NKR * 16, 128);
      const uint64_t db = make_smem_desc(smem_u32(Qa) + qrow * 16, kN * 16, 128);
      mma_ss(tmem + cST, da, db, idescT, 1);
    }
#pragma unroll
    for (int ks = 0; ks < DHP / 16; ++ks) {
      const uint64_t da = make_smem_desc(smem_u32(Vs) + ks * 2 * (NKR * 16) + kb * (128 * 16), NKR * 16, 128);
      const uint64_t db = make_smem_desc(smem_u32(dOs) + ks * 2 * (kN * 16) + qrow * 16, kN * 16, 128);
      mma_ss(tmem + cDPT, da, db, idescT, ks > 0);
    }
    mma_commit(&bar[bS]);
  };

  long long* tl = reinterpret_cast<long long*>(p.delta);
  int tli = 0;
  const bool rec = p.debug && blockIdx.x == 0 && (tid == 0 || tid == 128 || tid == 255);
  const int tlb = tid == 0 ? 0 : (tid == 128 ? 2048 : 4096);
#define STAMP(tag) do { if (rec && tli < 1000) { tl[tlb + 2 * tli] = clock64(); tl[tlb + 2 * tli + 1] = (tag); ++tli; } } while (0)
  for (int bw = blockIdx.x / p.heads; bw < n_pairs; bw += stride) {
    const int b = bw / p.P, win = bw - b * p.P;
    STAMP(1);
    // ---- stage this (window, head) ----
    {
      const int n = tid;                                         // one query row per thread
      const size_t goff = ((size_t)bw * kN + n) * p.C + head * DH;
      __nv_bfloat16 row[DH], orow[DH], extra[4];
      load_row_b<DH>((const __nv_bfloat16*)p.q + ((size_t)bw * kN + n) * p.ldq + head * DH, row);
      const int id_ = n % p.wd;
#pragma unroll
      for (int u = 0; u < 4; ++u) extra[u] = (u == id_) ? one : zero;
      store_chunks_b<DH, KS * 2>(Qs, kN * 16, n, row, extra, p.wd);
      load_row_b<DH>((const __nv_bfloat16*)p.dout + goff, row);
      load_row_b<DH>((const __nv_bfloat16*)p.out + goff, orow);
      float dl = 0.f;
#pragma unroll
      for (int d = 0; d < DH; ++d) dl = fmaf(__bfloat162float(row[d]), __bfloat162float(orow[d]), dl);
      // dO' = [dO | -delta (bf16 hi, lo)]: with V' = [V | 1 1] the dP^T MMA yields dP - delta directly
      extra[0] = __float2bfloat16(-dl);
      extra[1] = __float2bfloat16(-dl - __bfloat162float(extra[0]));
      store_chunks_b<DH, DHP / 8>(dOs, kN * 16, n, row, extra, FOLD ? 2 : 0);
      delta_s[n] = dl;
      const float l2 = p.lse[((size_t)bw * p.heads + head) * kN + n] * 1.4426950408889634f;
      lse2_s[n] = l2;
      wp_s[n] = __float2bfloat16(fast_exp2(-l2));
    }
    for (int j = tid; j < NKT; j += kThreadsB) {
      const bool content = j < kN;
      const size_t off = content ? ((size_t)bw * kN + j) * p.ldq + head * DH : ((size_t)b * p.I + (j - kN)) * p.ldp + head * DH;
      __nv_bfloat16 row[DH], extra[4];
      load_row_b<DH>((const __nv_bfloat16*)(content ? p.k : p.kp) + off, row);
      const int jd = j % p.wd;
#pragma unroll
      for (int u = 0; u < 4; ++u)
        extra[u] = (content && u < p.wd) ? __float2bfloat16(p.td[(head * p.wd + u) * p.wd + jd] * inv_scale) : zero;
      store_chunks_b<DH, KS * 2>(Ks, NKR * 16, j, row, extra, p.wd);
      load_row_b<DH>((const __nv_bfloat16*)(content ? p.v : p.vp) + off, row);
      extra[0] = one;
      extra[1] = one;
      store_chunks_b<DH, DHP / 8>(Vs, NKR * 16, j, row, extra, FOLD ? 2 : 0);
    }
    if (MASKED)
      for (int i = tid; i < kN / 4; i += kThreadsB)
        reinterpret_cast<uint32_t*>(ids_s)[i] = reinterpret_cast<const uint32_t*>(p.ids + (size_t)win * kN)[i];
    STAMP(2);
    fence_proxy_async_smem();
    __syncthreads();
    STAMP(3);
    if (warp == 4 && lane == 0) {
      // the packed g^T columns of the previous window's last unit must have been consumed (dKaug is never awaited elsewhere)
      if (chains_pending) {
        mbar_wait(&bar[bV], phC ^ 1);
        mbar_wait(&bar[bK], phC ^ 1);
        mbar_wait(&bar[bA], phC ^ 1);
        chains_pending = false;
      }
      tc_fence_after();
      issue_scores(0, 0, 0);
    }
    if (MASKED) {
      // PRMT selectors: word w of id slot s covers tokens 4w..4w+3 = packed pairs 2w (low half) and 2w+1 (high half);
      // a kept bf16 takes its own bytes (nibbles 1,0 / 3,2), a masked one the bytes of the second operand (5,4 / 7,6)
      for (int i = tid; i < kIds * (kN / 4); i += kThreadsB) {
        const int s = i / (kN / 4), w = i - s * (kN / 4);
        const uint32_t idw = reinterpret_cast<const uint32_t*>(ids_s)[w];
        uint32_t sel = 0;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const bool keep = id_slot((idw >> (8 * e)) & 0xffu) == s;
          const uint32_t nib = (e & 1) ? (keep ? 0x32u : 0x76u) : (keep ? 0x10u : 0x54u);
          sel |= nib << (((e & 1) ? 8 : 0) + ((e >> 1) ? 16 : 0));
        }
        sel_s[i] = sel;
      }
      __syncthreads();
    }

    for (int unit = 0; unit < n_units; ++unit) {
      const int kb = unit >> 2, u = unit & 3, mt = u >> 1, hf = u & 1;
      const int nk = kb < 2 ? 128 : p.I;                         // valid keys in this block
      const bool warp_ok = (warp & 3) * 32 < nk;                 // (nk is a multiple of 32: whole warps are valid or not)
      const bool do_mask = MASKED && kb < 2;
      STAMP(10 + unit);
      __syncwarp();
      mbar_wait(&bar[bS], phS);
      phS ^= 1;
      tc_fence_after();
      STAMP(100);
      // dQ'(previous query tile) reads Gs: it must have retired before this unit's g^T overwrites the tile
      if (hf == 0 && unit > 0) {
        mbar_wait(&bar[bQ], phQ);
        phQ ^= 1;
      }
      if (warp_ok) {
        // ---- this thread: key = lane_row, rows r0 .. r0+31 ----
        const int r0 = mt * 128 + hf * 64 + wg * 32;
        uint32_t s[32], dp[32];
        tmem_ld32(trow + cST + wg * 32, s);
        tmem_ld32(trow + cDPT + wg * 32, dp);
        tmem_wait_ld();
        STAMP(101);
        uint32_t pk[16], gk[16];
#pragma unroll
        for (int q4 = 0; q4 < 8; ++q4) {
          const float4 l4 = *reinterpret_cast<const float4*>(lse2_s + r0 + q4 * 4);
          const float lv[4] = {l4.x, l4.y, l4.z, l4.w};
          float dv[4] = {0.f, 0.f, 0.f, 0.f};
          if (!FOLD) {
            const float4 d4 = *reinterpret_cast<const float4*>(delta_s + r0 + q4 * 4);
            dv[0] = d4.x; dv[1] = d4.y; dv[2] = d4.z; dv[3] = d4.w;
          }
          float pv[4], gv[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int r = q4 * 4 + e;
            pv[e] = fast_exp2(fmaf(__uint_as_float(s[r]), c2, -lv[e]));
            gv[e] = pv[e] * (FOLD ? __uint_as_float(dp[r]) : __uint_as_float(dp[r]) - dv[e]);
          }
          pk[q4 * 2] = pack_bf16(pv[0], pv[1]);
          pk[q4 * 2 + 1] = pack_bf16(pv[2], pv[3]);
          gk[q4 * 2] = pack_bf16(gv[0], gv[1]);
          gk[q4 * 2 + 1] = pack_bf16(gv[2], gv[3]);
        }
        if (do_mask) {
          const uint32_t cid = ids_s[kb * 128 + lane_row];
          const uint4* sp = reinterpret_cast<const uint4*>(sel_s + id_slot(cid) * (kN / 4) + r0 / 4);
          const uint4* wpp = reinterpret_cast<const uint4*>(wp_s + r0);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const uint4 s4 = sp[h];
            const uint32_t sw[4] = {s4.x, s4.y, s4.z, s4.w};
            const uint4 wa = wpp[h * 2], wb = wpp[h * 2 + 1];
            const uint32_t ww[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
            for (int w = 0; w < 4; ++w) {
              pk[h * 8 + w * 2] = prmt(pk[h * 8 + w * 2], ww[w * 2], sw[w]);
              pk[h * 8 + w * 2 + 1] = prmt(pk[h * 8 + w * 2 + 1], ww[w * 2 + 1], sw[w] >> 16);
              gk[h * 8 + w * 2] = prmt(gk[h * 8 + w * 2], 0u, sw[w]);
              gk[h * 8 + w * 2 + 1] = prmt(gk[h * 8 + w * 2 + 1], 0u, sw[w] >> 16);
            }
          }
        }
        STAMP(102);
        tmem_st16(trow + cST + wg * 32, pk);                     // packed over this warpgroup's own consumed columns
        tmem_st16(trow + cDPT + wg * 32, gk);
        // g^T -> smem as the MN-major A operand of dQ = g.K : [row group of 8][key group of 8][key%8][16 B]
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int rg = (hf * 64 + wg * 32) / 8 + q;
          *reinterpret_cast<uint4*>(Gs + rg * 2048 + lane_row * 16) = make_uint4(gk[q * 4], gk[q * 4 + 1], gk[q * 4 + 2], gk[q * 4 + 3]);
        }
        tmem_wait_st();
      }
      STAMP(103);
      fence_proxy_async_smem();
      tc_fence_before();
      STAMP(104);
      __syncthreads();
      STAMP(105);
      if (lane == 0 && warp < 5) {
        tc_fence_after();
        const uint32_t qrow = (uint32_t)(mt * 128 + hf * 64);
        const uint32_t acc0 = u > 0;
        // K = 64 rows = 4 k-steps; A = packed bf16 in TMEM: rows 0-31 live at cols 0-15, rows 32-63 at cols 32-47
        if (warp == 0) {                                         // dV += P^T.dO
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const uint32_t acol = (t >> 1) * 32 + (t & 1) * 8;
            const uint64_t bdo = make_smem_desc(smem_u32(dOs) + (qrow + t * 16) * 16, 128, kN * 16);
            mma_ts(tmem + cDV, tmem + cST + acol, bdo, idescDV, acc0 | (t > 0));
          }
          mma_commit(&bar[bV]);
          STAMP(108);
        } else if (warp == 1) {                                  // dK' += g^T.Q'
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const uint32_t acol = (t >> 1) * 32 + (t & 1) * 8;
            const uint64_t bq = make_smem_desc(smem_u32(Qs) + (qrow + t * 16) * 16, 128, kN * 16);
            mma_ts(tmem + cDK, tmem + cDPT + acol, bq, idescDK, acc0 | (t > 0));
          }
          mma_commit(&bar[bK]);
        } else if (warp == 2) {                                  // dKaug += g^T.Qaug   (accumulates over all windows)
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const uint32_t acol = (t >> 1) * 32 + (t & 1) * 8;
            const uint64_t bqa = make_smem_desc(smem_u32(Qa) + (qrow + t * 16) * 16, 128, kN * 16);
            mma_ts(tmem + cAUG + kb * 16, tmem + cDPT + acol, bqa, idescAUG, (!first_window) | acc0 | (t > 0));
          }
          mma_commit(&bar[bA]);
        } else if (warp == 3) {
          if (hf == 1) {                                         // dQ'[mt] += g[128 rows x nk keys] . K'[kb]
            for (int t = 0; t < nk / 16; ++t) {
              const uint64_t da = make_smem_desc(smem_u32(Gs) + t * 256, 128, 2048);
              const uint64_t db = make_smem_desc(smem_u32(Ks) + (kb * 128 + t * 16) * 16, 128, NKR * 16);
              mma_ss(tmem + cDQ + mt * DKC, da, db, idescDQ, (kb > 0) | (t > 0));
            }
            mma_commit(&bar[bQ]);
          }
        } else {                                                 // warp 4: next unit's scores, once P^T / g^T are consumed
          if (unit + 1 < n_units) {
            mbar_wait(&bar[bV], phC);
            mbar_wait(&bar[bK], phC);
            mbar_wait(&bar[bA], phC);
            STAMP(106);
            tc_fence_after();
            const int nu = unit + 1;
            issue_scores(nu >> 2, (nu & 3) >> 1, nu & 1);
            STAMP(107);
          } else {
            chains_pending = true;
          }
        }
      }
      if (u == 3) {
        // ---- key block done: drain dV (warpgroup 0) and dK' (warpgroup 1) ----
        __syncwarp();
        mbar_wait(&bar[wg == 0 ? bV : bK], phC);
        tc_fence_after();
        const int key = kb * 128 + lane_row;
        const bool key_ok = lane_row < nk;
        if (wg == 0) {
          float dv[DHP];
#pragma unroll
          for (int dq = 0; dq < DHP / 16; ++dq) {
            uint32_t o[16];
            tmem_ld16(trow + cDV + dq * 16, o);
            tmem_wait_ld();
#pragma unroll
            for (int d = 0; d < 16; ++d) dv[dq * 16 + d] = __uint_as_float(o[d]);
          }
          if (key_ok) {
            if (kb < 2) {
              __nv_bfloat16* g = (__nv_bfloat16*)p.dv + ((size_t)bw * kN + key) * p.ldq + head * DH;
#pragma unroll
              for (int d = 0; d < DH; ++d) g[d] = __float2bfloat16(dv[d]);
            } else {
              float* g = p.dvp + ((size_t)b * p.I + lane_row) * p.C + head * DH;
#pragma unroll
              for (int d = 0; d < DH; ++d) atomicAdd(g + d, dv[d]);
            }
          }
        } else {
          float dk[DKC];
#pragma unroll
          for (int dq = 0; dq < DKC / 16; ++dq) {
            uint32_t o[16];
            tmem_ld16(trow + cDK + dq * 16, o);
            tmem_wait_ld();
#pragma unroll
            for (int d = 0; d < 16; ++d) dk[dq * 16 + d] = __uint_as_float(o[d]);
          }
          if (key_ok) {
            if (kb < 2) {
              __nv_bfloat16* g = (__nv_bfloat16*)p.dk + ((size_t)bw * kN + key) * p.ldq + head * DH;
#pragma unroll
              for (int d = 0; d < DH; ++d) g[d] = __float2bfloat16(dk[d] * p.scale);
#pragma unroll
              for (int uu = 0; uu < 4; ++uu) acc_d[kb][uu] += dk[DH + uu];   // d K'[DH+u] = sum_rows(id==u) g = dTd[u][jd]
            } else {
              float* g = p.dkp + ((size_t)b * p.I + lane_row) * p.C + head * DH;
#pragma unroll
              for (int d = 0; d < DH; ++d) atomicAdd(g + d, dk[d] * p.scale);
            }
          }
        }
        STAMP(109);
        tc_fence_before();       // ordered before the next unit's __syncthreads -> next block's chains (which restart dV / dK')
      }
      phC ^= 1;
    }
    // ---- all key blocks done: dQ' tiles (warpgroup w drains query tile w) ----
    {
      __syncwarp();
      mbar_wait(&bar[bQ], phQ);
      phQ ^= 1;
      tc_fence_after();
      float dq[DKC];
#pragma unroll
      for (int c = 0; c < DKC / 16; ++c) {
        uint32_t o[16];
        tmem_ld16(trow + cDQ + wg * DKC + c * 16, o);
        tmem_wait_ld();
#pragma unroll
        for (int d = 0; d < 16; ++d) dq[c * 16 + d] = __uint_as_float(o[d]);
      }
      __nv_bfloat16* g = (__nv_bfloat16*)p.dq + ((size_t)bw * kN + wg * 128 + lane_row) * p.ldq + head * DH;
#pragma unroll
      for (int d = 0; d < DH; ++d) g[d] = __float2bfloat16(dq[d] * p.scale);
    }
    STAMP(4);
    first_window = false;
    tc_fence_before();
    __syncthreads();
  }

  // ---- once per CTA: bias-table gradients ----
  // dKaug[key][u] = sum over all rows/windows of g * onehot: columns [0,wh) -> dTh[u][jh(key)], [wh,wh+ww) -> dTw[u][jw(key)];
  // for prompt keys the wh replicated columns sum to dtok[i].  The /scale of K'aug and the *scale of dS cancel.
  if (!first_window) {
    if (warp == 4 && lane == 0 && chains_pending) {
      mbar_wait(&bar[bA], phC ^ 1);
      chains_pending = false;
    }
    __syncthreads();
    tc_fence_after();
    for (int kb = wg; kb < n_kb; kb += 2) {
      uint32_t o[16];
      tmem_ld16(trow + cAUG + kb * 16, o);
      tmem_wait_ld();
      const int key = kb * 128 + lane_row;
      if (kb < 2) {
        const int jw = (key / p.wd) % p.ww, jh = key / (p.wd * p.ww);
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          if (c < p.wh) atomicAdd(&gth_s[c * p.wh + jh], __uint_as_float(o[c]));
          else if (c - p.wh < p.ww) atomicAdd(&gtw_s[(c - p.wh) * p.ww + jw], __uint_as_float(o[c]));
        }
      } else if (lane_row < p.I) {
        float t = 0.f;
#pragma unroll
        for (int c = 0; c < 16; ++c)
          if (c < p.wh) t += __uint_as_float(o[c]);
        atomicAdd(&gtok_s[lane_row], t);
      }
    }
    if (wg == 1) {
#pragma unroll
      for (int kb = 0; kb < 2; ++kb) {
        const int jd = (kb * 128 + lane_row) % p.wd;
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (u < p.wd) atomicAdd(&gtd_s[u * p.wd + jd], acc_d[kb][u]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (!first_window) {
    for (int i = tid; i < p.wh * p.wh; i += kThreadsB) atomicAdd(&p.dth[head * p.wh * p.wh + i], gth_s[i]);
    for (int i = tid; i < p.ww * p.ww; i += kThreadsB) atomicAdd(&p.dtw[head * p.ww * p.ww + i], gtw_s[i]);
    for (int i = tid; i < p.wd * p.wd; i += kThreadsB) atomicAdd(&p.dtd[head * p.wd * p.wd + i], gtd_s[i]);
    for (int i = tid; i < p.I; i += kThreadsB) atomicAdd(&p.dtok[head * p.I + i], gtok_s[i]);
  }
  if (warp == 0) tmem_dealloc(tmem, tmem_cols);
}

template <int DH>
int launch_bwd_tc(const AttnParams& p, cudaStream_t st) {
  constexpr int DHP = (DH + 15) / 16 * 16;
  constexpr int KS = (DH + 4 + 15) / 16;
  const int NKT = kN + p.I;
  const BwdSmem L = bwd_layout(KS, DHP, NKT, p.wh, p.ww, p.wd, p.I, p.ids != nullptr);
  const size_t smem = L.total;
  const uint32_t need_cols = 128 + DHP + 3 * KS * 16 + 48;
  const uint32_t cols = need_cols <= 256 ? 256 : 512;
  const int per_sm = (cols == 256 && smem <= 110 * 1024) ? 2 : 1;
  int grid = 148 * per_sm;
  grid -= grid % p.heads;
  if (grid < p.heads) grid = p.heads;
  const int need = p.B * p.P * p.heads;
  if (grid > need) grid = need;
  auto kern = p.ids ? attn_bwd_tc_kernel<DH, true> : attn_bwd_tc_kernel<DH, false>;
  PWA_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, kThreadsB, smem, st>>>(p, cols);
  PWA_CUDA_OK(cudaGetLastError());
  return PWA_OK;
}

}  // namespace

bool attn_tc_bwd_supported(const AttnParams& p, int dtype) {
  if (!attn_tc_supported(p, dtype)) return false;
  const int dh = p.C / p.heads;
  const int KS = (dh + 4 + 15) / 16, DHP = (dh + 15) / 16 * 16;
  if (128 + DHP + 3 * KS * 16 + 48 > 512) return false;
  return bwd_layout(KS, DHP, kN + p.I, p.wh, p.ww, p.wd, p.I, true).total <= 220 * 1024;
}

int attn_tc_backward(const AttnParams& p, cudaStream_t st) {
  switch (p.C / p.heads) {
    case 3: return launch_bwd_tc<3>(p, st);
    case 6: return launch_bwd_tc<6>(p, st);
    case 12: return launch_bwd_tc<12>(p, st);
    case 24: return launch_bwd_tc<24>(p, st);
    case 48: return launch_bwd_tc<48>(p, st);
  }
  set_error("tcgen05 attention backward: head_dim %d not instantiated", p.C / p.heads);
  return PWA_ERR_UNSUPPORTED;
}

}  // namespace pwa
