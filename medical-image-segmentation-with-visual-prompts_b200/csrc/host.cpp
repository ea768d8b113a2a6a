// Host-only part of the C ABI: error channel, window geometry, region ids, index maps.
// No CUDA calls here, so these entry points work (and are tested) on a machine without a GPU.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <vector>

#include "../../include/pwa.h"

namespace pwa {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace pwa

extern "C" int pwa_version(void) { return PWA_VERSION; }
extern "C" const char* pwa_last_error(void) { return pwa::g_err; }

// swin_block.py:146-164 (padding rule) and :265-270 (effective shift).
extern "C" int pwa_geometry(const int32_t dims[3], const int32_t ws[3], const int32_t shift_cfg[3], pwa_geom* g) {
  if (!dims || !ws || !shift_cfg || !g) {
    pwa::set_error("pwa_geometry: null argument");
    return PWA_ERR_ARG;
  }
  memset(g, 0, sizeof(*g));
  bool need_pad = false;
  for (int a = 0; a < 3; ++a) {
    if (dims[a] <= 0 || ws[a] <= 0 || shift_cfg[a] < 0 || shift_cfg[a] >= ws[a]) {
      pwa::set_error("pwa_geometry: bad axis %d: dim %d window %d shift %d", a, dims[a], ws[a], shift_cfg[a]);
      return PWA_ERR_ARG;
    }
    if (dims[a] % ws[a] != 0) need_pad = true;
  }
  g->P = 1;
  g->N = 1;
  for (int a = 0; a < 3; ++a) {
    g->dims[a] = dims[a];
    g->ws[a] = ws[a];
    // a shift survives only on axes strictly larger than the window (unpadded size)
    g->shift[a] = dims[a] <= ws[a] ? 0 : shift_cfg[a];
    int lo = 0, hi = 0;
    if (need_pad) {  // every axis is padded once any axis needs it; a divisible axis grows by a whole window
      int r = ws[a] - dims[a] % ws[a];
      lo = r / 2;
      hi = r - lo;
    }
    g->pads[2 * a] = lo;
    g->pads[2 * a + 1] = hi;
    g->sp[a] = dims[a] + lo + hi;
    g->nwin[a] = g->sp[a] / ws[a];
    g->data_lo[a] = hi;  // F.pad(x, reversed(paddings)) swaps each (lo,hi) pair
    g->crop_lo[a] = lo;
    g->P *= g->nwin[a];
    g->N *= ws[a];
    if (g->shift[a] > 0) g->masked = 1;
    if (lo + hi > 0) g->padded = 1;
  }
  return PWA_OK;
}

// Per-axis region index in the rolled frame.  The reference fills three Python slices in order,
// later fills overwriting earlier ones: [0,-w) -> 0, [-w,-s) -> 1, [-s,None) -> 2; for s == 0 the
// last slice is the whole axis (swin_block.py:320-334).
static std::vector<int> axis_region(int sp, int w, int s) {
  std::vector<int> reg(sp, 0);
  auto norm = [sp](int v) { return v < 0 ? (v + sp < 0 ? 0 : v + sp) : (v > sp ? sp : v); };
  int b0 = norm(0), e0 = norm(-w);
  for (int i = b0; i < e0; ++i) reg[i] = 0;
  int b1 = norm(-w), e1 = (s == 0) ? 0 : norm(-s);  // slice(-w, -0) == slice(-w, 0): empty
  for (int i = b1; i < e1; ++i) reg[i] = 1;
  int b2 = (s == 0) ? 0 : norm(-s);                  // slice(-0, None) == whole axis
  for (int i = b2; i < sp; ++i) reg[i] = 2;
  return reg;
}

extern "C" int pwa_region_ids(const pwa_geom* g, uint8_t* ids) {
  if (!g || !ids) {
    pwa::set_error("pwa_region_ids: null argument");
    return PWA_ERR_ARG;
  }
  std::vector<int> reg[3];
  for (int a = 0; a < 3; ++a) reg[a] = axis_region(g->sp[a], g->ws[a], g->shift[a]);
  const int P1 = g->nwin[0], P2 = g->nwin[1], P3 = g->nwin[2];
  size_t o = 0;
  for (int p1 = 0; p1 < P1; ++p1)
    for (int p2 = 0; p2 < P2; ++p2)
      for (int p3 = 0; p3 < P3; ++p3)
        for (int t1 = 0; t1 < g->ws[0]; ++t1)
          for (int t2 = 0; t2 < g->ws[1]; ++t2)
            for (int t3 = 0; t3 < g->ws[2]; ++t3) {
              // strided windows: rolled-frame coordinate = token index * #windows + window index
              const int c[3] = {t1 * P1 + p1, t2 * P2 + p2, t3 * P3 + p3};
              int id = 9 * reg[0][c[0]] + 3 * reg[1][c[1]] + reg[2][c[2]];
              if (g->padded) {
                bool in = true;
                for (int a = 0; a < 3; ++a)
                  in = in && c[a] >= g->pads[2 * a] && c[a] < g->sp[a] - g->pads[2 * a + 1];
                if (in) id = 100;
              }
              ids[o++] = (uint8_t)id;
            }
  return PWA_OK;
}

extern "C" int pwa_index_map(const pwa_geom* g, int which, int32_t* map) {
  if (!g || !map || (which != 0 && which != 1)) {
    pwa::set_error("pwa_index_map: bad argument");
    return PWA_ERR_ARG;
  }
  const int32_t* lo = which == 0 ? g->data_lo : g->crop_lo;
  const int P1 = g->nwin[0], P2 = g->nwin[1], P3 = g->nwin[2];
  size_t o = 0;
  for (int p1 = 0; p1 < P1; ++p1)
    for (int p2 = 0; p2 < P2; ++p2)
      for (int p3 = 0; p3 < P3; ++p3)
        for (int t1 = 0; t1 < g->ws[0]; ++t1)
          for (int t2 = 0; t2 < g->ws[1]; ++t2)
            for (int t3 = 0; t3 < g->ws[2]; ++t3) {
              const int c[3] = {t1 * P1 + p1, t2 * P2 + p2, t3 * P3 + p3};
              int src[3];
              bool ok = true;
              for (int a = 0; a < 3; ++a) {
                src[a] = (c[a] + g->shift[a]) % g->sp[a] - lo[a];  // roll(-s): out[i] = in[(i+s) mod S]
                ok = ok && src[a] >= 0 && src[a] < g->dims[a];
              }
              map[o++] = ok ? (src[0] * g->dims[1] + src[1]) * g->dims[2] + src[2] : -1;
            }
  return PWA_OK;
}

// PRMT selector table of the shift mask for the tcgen05 attention kernels, per window: [P][28][N/4] uint32 in HOST
// memory.  Row-id slot s (ids 0..26, 100 -> 27), word w covers keys 4w..4w+3 = packed bf16 pairs 2w (low half of the
// word) and 2w+1 (high half): a key of the row's own region keeps its bytes (nibbles 1,0 / 3,2), any other key takes
// those of the second PRMT operand (5,4 / 7,6).  It depends only on the geometry (swin_block.py:312-364 through
// pwa_region_ids), so it is built once and cached next to the region ids; the forward kernel then fetches a window's
// 7 KB with one bulk copy instead of rebuilding it for every (window, head).
extern "C" int pwa_attn_sel_table(const uint8_t* ids, int P, int N, uint32_t* table) {
  if (!ids || !table || P <= 0 || N <= 0 || N % 4 != 0) {
    pwa::set_error("pwa_attn_sel_table: bad argument (P=%d N=%d)", P, N);
    return PWA_ERR_ARG;
  }
  const int W = N / 4;
  for (int p = 0; p < P; ++p)
    for (int s = 0; s < 28; ++s)
      for (int w = 0; w < W; ++w) {
        uint32_t sel = 0;
        for (int e = 0; e < 4; ++e) {
          const uint32_t id = ids[(size_t)p * N + 4 * w + e];
          const bool keep = (int)(id < 27u ? id : 27u) == s;
          const uint32_t nib = (e & 1) ? (keep ? 0x32u : 0x76u) : (keep ? 0x10u : 0x54u);
          sel |= nib << (((e & 1) ? 8 : 0) + ((e >> 1) ? 16 : 0));
        }
        table[((size_t)p * 28 + s) * W + w] = sel;
      }
  return PWA_OK;
}
