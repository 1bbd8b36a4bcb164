// I-picture kernels of the B200 HEVC encoder (sm_100a).
//
//   k_intra_modes   mode decision for every CU of the picture at once: one CTA per 16x16 block
//                   (8x8 CUs where 16 does not fit the picture), one thread per sample, 35-mode
//                   SAD search (H.265 8.4.4.2) predicted from the
//                   SOURCE picture's neighbours with the standard's availability / substitution /
//                   smoothing rules.  No dependency on any reconstruction, so fully parallel.
//   k_intra_frame   reconstruction wavefront: intra prediction needs the reconstructed left /
//                   above / above-right neighbours, so CTUs run as a wavefront -- one CTA per CTU,
//                   CTU indices handed out in raster order by an atomic ticket (every CTU a CTA
//                   waits for is already running or done: no deadlock), a per-row progress counter
//                   published with a release fence.  Inside a CTU the CUs are coded in z-order;
//                   the three planes of a CU go through prediction, DCT, quantisation, inverse and
//                   reconstruction (rows K5, K6) concurrently: 256 + 64 + 64 threads.
#include <algorithm>

#include "hevc_intra.cuh"
#include "hevc_kernels.h"

namespace b200 {

namespace {

constexpr int kModeThreads = 256;     // k_intra_modes: one thread per luma sample of a 16x16 CU
constexpr int kReconThreads = 384;    // k_intra_frame: 256 luma + 64 Cb + 64 Cr samples of a 16x16 CU

// ---- mode decision, whole picture in parallel ----------------------------------------------------

__device__ void decide_cu(ModeShared &sh, int16_t (*resid)[256], const FrameParams &fp, const uint8_t *src, CuInfo *cu, int x0, int y0, int log2)
{
  const int n = 1 << log2, t = threadIdx.x;
  intra_search_cu(sh, fp, src, x0, y0, log2, fp.intra_satd ? resid : nullptr);
  if (t == 0) {
    CuInfo ci;
    ci.mvx = 0; ci.mvy = 0; ci.log2_size = (uint8_t)log2; ci.pred_mode = 1; ci.intra_mode = (uint8_t)sh.best_mode;
    ci.cbf = 0; ci.skip = 0; ci.merge_idx = 0xff; ci.mvp_idx = 0; ci.qp = 0;
      ci.ref_idx = 0; ci.chroma_mode = ci.intra_mode; ci.tu_log2 = ci.log2_size < 5 ? ci.log2_size : 5; ci.flags = 0;
    const int n8 = n >> 3;
    for (int j = 0; j < n8; j++)
      for (int i = 0; i < n8; i++) cu[(size_t)((y0 >> 3) + j) * fp.w8 + (x0 >> 3) + i] = ci;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kModeThreads)
k_intra_modes(FrameParams fp, const uint8_t *__restrict__ src, CuInfo *__restrict__ cu)
{
  __shared__ ModeShared sh;
  __shared__ int16_t resid[2][256];
  const int bw = (fp.w + 15) >> 4;
  const int x0 = (blockIdx.x % bw) * 16, y0 = (blockIdx.x / bw) * 16;
  if (x0 + 16 <= fp.w && y0 + 16 <= fp.h) {
    decide_cu(sh, resid, fp, src, cu, x0, y0, 4);
  } else {
    for (int q = 0; q < 4; q++) {
      int x1 = x0 + 8 * (q & 1), y1 = y0 + 8 * (q >> 1);
      if (x1 < fp.w && y1 < fp.h) decide_cu(sh, resid, fp, src, cu, x1, y1, 3);
    }
  }
}

// ---- reconstruction wavefront ----------------------------------------------------------------------

struct ReconSharedI {
  uint8_t rec_y[64 * 64], rec_c[2][32 * 32];      // reconstruction of the current CTU
  uint8_t src_y[64 * 64], src_c[2][32 * 32];      // source samples of the current CTU
  RefSet rs[3];
  CtuBorder bd;                                   // row above / column left of the CTU (gather_refs)
  uint32_t cuw[64][2];                            // words 1 and 3 of the CTU's cu map entries, raster order of 8x8 units
  int nz[3], ctu;
  int8_t dct[32][32], dctT[32][32];
  uint8_t pred[384];
  int16_t a[384], b[384];
};

// One CU: the three planes side by side.  Thread t: plane p (0: t < 256, 1: 256..319, 2: 320..383
// for a 16x16 CU), sample (x,y) of that plane's n_p x n_p block.
__device__ void recon_cu(ReconSharedI &sh, const FrameParams &fp, const uint8_t *src, uint8_t *rec, int16_t *levels,
                         CuInfo *cu, int cx, int cy, int x0, int y0, int log2)
{
  constexpr bool kDecode = false;          // the decoder has its own walk (k_intra_decode below)
  const int t = threadIdx.x;
  const size_t ysz = (size_t)fp.w * fp.h;
  const unsigned cur = coding_order_i(fp, x0, y0);
  const int nl = 1 << log2, nc = nl >> 1, ll = nl * nl, lc = nc * nc;
  // plane of this thread
  int p, li;
  if (t < ll) { p = 0; li = t; }
  else if (t < ll + lc) { p = 1; li = t - ll; }
  else if (t < ll + 2 * lc) { p = 2; li = t - ll - lc; }
  else { p = -1; li = 0; }
  const int l2 = p == 0 ? log2 : log2 - 1, n = 1 << l2;
  const int y = li >> l2, x = li & (n - 1);
  const int pw = p == 0 ? fp.w : fp.w >> 1;
  const size_t poff = p <= 0 ? 0 : ysz + (p == 2 ? ysz / 4 : 0);
  const int bx = p == 0 ? x0 : x0 >> 1, by = p == 0 ? y0 : y0 >> 1;
  const int qp = p == 0 ? qp_at(fp, x0, y0) : qp_c_at(fp, x0, y0);
  const int mode = (sh.cuw[(((y0 - cy) >> 3) << 3) + ((x0 - cx) >> 3)][0] >> 16) & 0xff;
  // neighbours: groups of 128 threads gather one plane each (4n+1 <= 65 samples, +1 for the DC sum)
  {
    const int g = t >> 7, gt = t & 127;                    // g = plane whose neighbours this thread helps with
    const int gn = g == 0 ? nl : nc;
    const size_t goff = g == 0 ? 0 : ysz + (g == 2 ? ysz / 4 : 0);
    const int gpw = g == 0 ? fp.w : fp.w >> 1;
    const int gx = g == 0 ? x0 : x0 >> 1, gy = g == 0 ? y0 : y0 >> 1;
    if (t < 3) sh.nz[t] = kDecode ? ((__ldcg(&cu[(size_t)(y0 >> 3) * fp.w8 + (x0 >> 3)].cbf) >> t) & 1) : 0;
    const uint8_t *tile = g == 0 ? sh.rec_y : sh.rec_c[g - 1];
    gather_refs(sh.rs[g], fp, rec + goff, gpw, g, gx, gy, gn, cur, gt, tile, g == 0 ? 64 : 32, g == 0 ? cx : cx >> 1, g == 0 ? cy : cy >> 1,
                g == 0 ? sh.bd.top_y : sh.bd.top_c[g - 1], g == 0 ? sh.bd.left_y : sh.bd.left_c[g - 1]);
    __syncthreads();
    substitute_refs(sh.rs[g], gn, gt);
    __syncthreads();
    filter_refs(sh.rs[g], gn, gt);
    __syncthreads();
  }
  const int nshift = 5 - l2;
  int16_t *a = sh.a + (p <= 0 ? 0 : ll + (p == 2 ? lc : 0));
  int16_t *b = sh.b + (p <= 0 ? 0 : ll + (p == 2 ? lc : 0));
  int pr = 0;
  if (p >= 0) pr = intra_pixel(sh.rs[p].sub, sh.rs[p].filt, n, l2, mode, p, sh.rs[p].dc, x, y);
  if (!kDecode) {
    if (p >= 0) {
      const int T = p == 0 ? 64 : 32;
      const uint8_t *st = p == 0 ? sh.src_y : sh.src_c[p - 1];
      int s = st[(by + y - (p == 0 ? cy : cy >> 1)) * T + bx + x - (p == 0 ? cx : cx >> 1)];
      a[li] = (int16_t)(s - pr);
    }
    __syncthreads();
    if (p >= 0) {
      int acc = 0, kk = x << nshift;
      for (int i = 0; i < n; i++) acc += sh.dctT[i][kk] * a[y * n + i];
      int s1 = l2 - 1;
      b[li] = (int16_t)((acc + (1 << (s1 - 1))) >> s1);
    }
    __syncthreads();
  }
  if (p >= 0) {
    const int qper = qp / 6, qrem = qp % 6;
    int lvl;
    if (!kDecode) {
      int acc = 0;
      const int8_t *c = sh.dct[y << nshift];
      for (int j = 0; j < n; j++) acc += c[j] * b[j * n + x];
      int s2 = l2 + 6;
      int coef = (acc + (1 << (s2 - 1))) >> s2;
      int qbits = 14 + qper + (7 - l2);
      unsigned add = (unsigned)(fp.is_idr ? 171 : 85) << (qbits - 9);
      unsigned av = ((unsigned)abs(coef) * sl_quant_scale(fp.scaling, qrem, l2, p, x, y) + add) >> qbits;
      lvl = (int)min(av, 32767u);
      if (coef < 0) lvl = -lvl;
      levels[poff + (size_t)(by + y) * pw + bx + x] = (int16_t)lvl;
      if (lvl) sh.nz[p] = 1;
    } else {
      lvl = sh.nz[p] ? levels[poff + (size_t)(by + y) * pw + bx + x] : 0;
    }
    int bd = l2 + 3;
    long long d = ((long long)lvl * (sl_factor(fp.scaling, l2, p, x, y) * c_level_scale[qrem])) << qper;
    d = (d + (1LL << (bd - 1))) >> bd;
    a[li] = (int16_t)max(-32768LL, min(32767LL, d));
  }
  __syncthreads();
  const int nz = p >= 0 ? sh.nz[p] : 0;
  if (nz) {
    int acc = 0;
    for (int k = 0; k < n; k++) acc += sh.dct[k << nshift][y] * a[k * n + x];
    b[li] = (int16_t)clip3(-32768, 32767, (acc + 64) >> 7);
  }
  __syncthreads();
  if (p >= 0) {
    int v = pr;
    if (nz) {
      int acc = 0;
      for (int k = 0; k < n; k++) acc += sh.dct[k << nshift][x] * b[y * n + k];
      v = clip8(pr + clip3(-32768, 32767, (acc + 2048) >> 12));
    }
    __stcg(rec + poff + (size_t)(by + y) * pw + bx + x, (uint8_t)v);
    const int T = p == 0 ? 64 : 32;
    uint8_t *rt = p == 0 ? sh.rec_y : sh.rec_c[p - 1];
    rt[(by + y - (p == 0 ? cy : cy >> 1)) * T + bx + x - (p == 0 ? cx : cx >> 1)] = (uint8_t)v;
  }
  const int n8 = nl >> 3;
  if (!kDecode && t < n8 * n8) {
    int cbf = (sh.nz[0] ? 1 : 0) | (sh.nz[1] ? 2 : 0) | (sh.nz[2] ? 4 : 0);
    __stcg(&cu[(size_t)((y0 >> 3) + t / n8) * fp.w8 + (x0 >> 3) + t % n8].cbf, (uint8_t)cbf);
  }
  // no device-scope fence here: later CUs of this CTU read these samples from the shared-memory
  // tile (ordered by the barrier); other CTUs only after the fence + progress update at CTU end
  __syncthreads();
}

__global__ void __launch_bounds__(kReconThreads, 2)     // <= 85 registers: a CTA then fits beside two motion-search CTAs
k_intra_frame(FrameParams fp, const uint8_t *__restrict__ src, uint8_t *rec, int16_t *levels, CuInfo *cu,
              int *ticket, const int *__restrict__ order)
{
  constexpr bool kDecode = false;
  __shared__ ReconSharedI sh;
  const int t = threadIdx.x;
  // P picture: the intra CUs (if any) follow the inter reconstruction; a picture without any is done
  if (!fp.is_idr && *(volatile int *)fp.any_intra == 0) return;
  for (int i = t; i < 1024; i += kReconThreads) {
    ((int8_t *)sh.dct)[i] = c_dct32[i >> 5][i & 31];
    ((int8_t *)sh.dctT)[i] = c_dct32[i & 31][i >> 5];
  }
  // Persistent CTAs: the wavefront is at most min(rows, ceil(cols/2)) CTUs wide, so a small grid
  // that keeps taking tickets does the same work as one CTA per CTU without parking hundreds of
  // waiting CTAs (384 threads x ~150 registers each) on SMs that other streams' kernels could use.
  const int nctu = fp.ctb_cols * fp.ctb_rows;
  const size_t ysz = (size_t)fp.w * fp.h;
  for (;;) {
    __syncthreads();
    if (t == 0) sh.ctu = atomicAdd(ticket, 1);
    __syncthreads();
    if (sh.ctu >= nctu) break;
    // tickets follow the wavefront (anti-diagonals c + 2r), not raster order, so that the CTUs a
    // small resident grid holds at any time are the ones that can actually run concurrently;
    // every dependency (left, above-left, above, above-right) still has a smaller ticket
    const int ctu = order[sh.ctu];
    const int row = ctu / fp.ctb_cols, col = ctu - row * fp.ctb_cols;
    const int cx = col * kCtb, cy = row * kCtb;
    // does this CTU hold intra CUs at all?  (always in an I picture)
    int has = 1;
    if (!fp.is_idr) {
      int mine = 0;
      if (t < 64) {
        const int x8 = (cx >> 3) + z_to_x(t), y8 = (cy >> 3) + z_to_y(t);
        if (x8 < fp.w8 && y8 < fp.h8) mine = cu[(size_t)y8 * fp.w8 + x8].pred_mode == 1;
      }
      has = __syncthreads_or(mine);
    }
    if (has) {
      if (t == 0) {
        // One flag per CTU (a CTU without intra CUs completes at once, so "everything before column c
        // of this row" no longer follows from its left neighbour): wait for the four CTUs whose
        // samples intra prediction can read.
        volatile int *d = fp.ctu_done;
        if (col > 0) while (d[ctu - 1] == 0) __nanosleep(32);
        if (row > 0) {
          if (col > 0) while (d[ctu - fp.ctb_cols - 1] == 0) __nanosleep(32);
          while (d[ctu - fp.ctb_cols] == 0) __nanosleep(32);
          if (col + 1 < fp.ctb_cols) while (d[ctu - fp.ctb_cols + 1] == 0) __nanosleep(32);
        }
        __threadfence();
      }
      __syncthreads();
      if (!fp.is_idr) {
        // the inter CUs of this CTU were reconstructed by the kernel before: bring the CTU's samples
        // into the tile that neighbour gathering reads
        for (int i = t; i < 64 * 16; i += kReconThreads) {
          int y = cy + (i >> 4), x = cx + 4 * (i & 15);
          ((uint32_t *)sh.rec_y)[i] = (y < fp.h && x < fp.w) ? __ldcg((const uint32_t *)(rec + (size_t)y * fp.w + x)) : 0u;
        }
        for (int i = t; i < 2 * 32 * 8; i += kReconThreads) {
          int c = i >> 8, j = i & 255;
          int y = (cy >> 1) + (j >> 3), x = (cx >> 1) + 4 * (j & 7);
          const uint8_t *pl = rec + ysz + (c ? ysz / 4 : 0);
          ((uint32_t *)sh.rec_c[c])[j] = (y < (fp.h >> 1) && x < (fp.w >> 1)) ? __ldcg((const uint32_t *)(pl + (size_t)y * (fp.w >> 1) + x)) : 0u;
        }
      }
      load_ctu_border(sh.bd, fp, rec, cx, cy, t, kReconThreads);
      if (t < 64) {
        const int x8 = (cx >> 3) + (t & 7), y8 = (cy >> 3) + (t >> 3);
        uint32_t w1 = 0, w3 = 0;
        if (x8 < fp.w8 && y8 < fp.h8) {
          const uint32_t *e = (const uint32_t *)(cu + (size_t)y8 * fp.w8 + x8);
          w1 = __ldcg(e + 1); w3 = __ldcg(e + 3);
        }
        sh.cuw[t][0] = w1; sh.cuw[t][1] = w3;
      }
      if (!kDecode) {
        for (int i = t; i < 64 * 16; i += kReconThreads) {
          int y = cy + (i >> 4), x = cx + 4 * (i & 15);
          ((uint32_t *)sh.src_y)[i] = (y < fp.h && x < fp.w) ? __ldg((const uint32_t *)(src + (size_t)y * fp.w + x)) : 0u;
        }
        for (int i = t; i < 2 * 32 * 8; i += kReconThreads) {
          int c = i >> 8, j = i & 255;
          int y = (cy >> 1) + (j >> 3), x = (cx >> 1) + 4 * (j & 7);
          const uint8_t *pl = src + ysz + (c ? ysz / 4 : 0);
          ((uint32_t *)sh.src_c[c])[j] = (y < (fp.h >> 1) && x < (fp.w >> 1)) ? __ldg((const uint32_t *)(pl + (size_t)y * (fp.w >> 1) + x)) : 0u;
        }
      }
      __syncthreads();
      // z-order walk over the sixteen 16x16 positions of the CTU: an intra CU of 16x16, or the 8x8
      // intra CUs among its four quarters (picture edges of an I picture; anywhere in a foreign stream)
      for (int z16 = 0; z16 < 16; z16++) {
        int x0 = cx + 16 * ((z16 & 1) | ((z16 >> 1) & 2)), y0 = cy + 16 * (((z16 >> 1) & 1) | ((z16 >> 2) & 2));
        if (x0 >= fp.w || y0 >= fp.h) continue;
        // prediction and reconstruction go transform unit by transform unit (8.4.4.1): 16x16 units
        // here, else the 8x8 units among the four quarters
        const uint32_t *u = sh.cuw[(((y0 - cy) >> 3) << 3) + ((x0 - cx) >> 3)];     // [0]: log2_size | pred_mode << 8 | ...; [1]: ... | tu_log2 << 16
        const int u_tu = (u[1] >> 16) & 0xff;
        if (u_tu == 4) {
          if (((u[0] >> 8) & 0xff) == 1) recon_cu(sh, fp, src, rec, levels, cu, cx, cy, x0, y0, 4);
        } else if (u_tu <= 3) {
          for (int q = 0; q < 4; q++) {
            int x1 = x0 + 8 * (q & 1), y1 = y0 + 8 * (q >> 1);
            if (x1 >= fp.w || y1 >= fp.h) continue;
            const uint32_t *v = sh.cuw[(((y1 - cy) >> 3) << 3) + ((x1 - cx) >> 3)];
            if (((v[0] >> 8) & 0xff) == 1 && ((v[1] >> 16) & 0xff) == 3) recon_cu(sh, fp, src, rec, levels, cu, cx, cy, x1, y1, 3);
          }
        }
      }
      __threadfence();
    }
    __syncthreads();
    if (t == 0) atomicExch(&fp.ctu_done[ctu], 1);
  }
}


// ---- decoder: intra CUs of any conforming stream -------------------------------------------------
// Same wavefront over CTUs as k_intra_frame; inside a CTU the walk follows the cu map the parser
// wrote: transform units in z order, each predicted from the reconstruction before it (8.4.4.1) --
// luma blocks of 4x4 (DST-VII, four per 8x8 unit, NxN CUs with a mode per block) up to 32x32 (with
// strong smoothing when the SPS says so), chroma blocks of half the size with their own mode.  The
// three planes of a unit go side by side: threads 0..255 luma, 256..319 Cb, 320..383 Cr, up to four
// samples per thread.

struct RefSetD { uint8_t raw[132], sub[132], filt[132], av[132]; unsigned avm[5]; int dc; };     // 4 * 32 + 1 neighbours; avm = availability bits

struct DecSharedI {
  uint8_t rec_y[64 * 64], rec_c[2][32 * 32];      // reconstruction of the current CTU
  uint32_t cu[64][4];                             // its cu map entries, z order (log2_size 0 = outside the picture)
  RefSetD rs[3];
  CtuBorder bd;                                   // row above / column left of the CTU
  int ctu;
  int8_t dct[32][32];
  int16_t a[1024 + 2 * 256], b[1024 + 2 * 256];   // a 32x32 luma block and two 16x16 chroma blocks
};

static __constant__ int8_t c_dst4[4][4] = {{29, 55, 74, 84}, {74, 74, 0, -74}, {84, -29, -74, 55}, {55, -84, 74, -29}};

__device__ __forceinline__ unsigned coding_order4(const FrameParams &fp, int x, int y)
{
  return coding_order_i(fp, x, y) * 4u + (unsigned)((((y >> 2) & 1) << 1) | ((x >> 2) & 1));
}

// One transform unit.  Luma block of 1 << l2y at luma (lx, ly) with mode_y; when with_c, the chroma blocks
// of 1 << l2c at chroma (lx >> 1 & ~3.., see caller) = (ccx, ccy) with mode_c.  cbf: bit 0 luma, 1 Cb, 2 Cr.
__device__ void dec_tu(DecSharedI &sh, const FrameParams &fp, uint8_t *rec, const int16_t *levels, int cx, int cy,
                       int lx, int ly, int l2y, int mode_y, bool with_c, int ccx, int ccy, int l2c, int mode_c, int cbf)
{
  const int t = threadIdx.x;
  const size_t ysz = (size_t)fp.w * fp.h;
  const int p = t < 256 ? 0 : (t < 320 ? 1 : 2);
  const int gi = p == 0 ? t : (p == 1 ? t - 256 : t - 320), gs = p == 0 ? 256 : 64;
  const bool act = p == 0 || with_c;
  const int l2 = p == 0 ? l2y : l2c, n = 1 << l2, cnt = 4 * n + 1;
  const int bx = p == 0 ? lx : ccx, by = p == 0 ? ly : ccy, sft = p ? 1 : 0;
  const int pw = fp.w >> sft;
  const size_t poff = p == 0 ? 0 : ysz + (p == 2 ? ysz / 4 : 0);
  const int T = p == 0 ? 64 : 32, tx0 = cx >> sft, ty0 = cy >> sft;
  uint8_t *tile = p == 0 ? sh.rec_y : sh.rec_c[p - 1];
  RefSetD &rs = sh.rs[p];
  const int mode = p == 0 ? mode_y : mode_c;
  const bool nz = (cbf >> p) & 1;
  const bool dst = p == 0 && l2 == 2;
  {
    // neighbours (8.4.4.2.2): available = inside the picture and earlier in decoding order, at 4x4 granularity.
    // The availability bits go to rs.avm by warp ballot (a warp belongs to one plane group and the loop
    // bound is the group's: every lane reaches the ballot).
    const unsigned cur = coding_order4(fp, bx << sft, by << sft);
    for (int base = 0; base < cnt; base += gs) {
      const int k = base + gi;
      bool ok = false;
      if (act && k < cnt) {
      int x, y;
      if (k < 2 * n) { x = bx - 1; y = by + 2 * n - 1 - k; }
      else if (k == 2 * n) { x = bx - 1; y = by - 1; }
      else { x = bx + (k - 2 * n - 1); y = by - 1; }
      const int ax = x << sft, ay = y << sft;
      ok = ax >= 0 && ay >= 0 && ax < fp.w && ay < fp.h && coding_order4(fp, ax, ay) < cur;
      uint8_t v = 0;
      if (ok) {
        // inside the CTU: its tile; else the border fetched once per CTU (every available neighbour lies there)
        const int ux = x - tx0, uy = y - ty0;
        const uint8_t *top = p == 0 ? sh.bd.top_y : sh.bd.top_c[p - 1], *left = p == 0 ? sh.bd.left_y : sh.bd.left_c[p - 1];
        if (ux >= 0 && uy >= 0 && ux < T && uy < T) v = tile[uy * T + ux];
        else if (uy == -1 && ux >= -1 && ux < 2 * T) v = top[ux + 1];
        else if (ux == -1 && uy >= 0 && uy < T) v = left[uy];
        else v = __ldcg(rec + poff + (size_t)y * pw + x);
      }
      rs.av[k] = ok; rs.raw[k] = v;
      }
      const unsigned bal = __ballot_sync(0xffffffffu, ok);
      if (act && (gi & 31) == 0 && k < 160) rs.avm[k >> 5] = bal;
    }
  }
  __syncthreads();
  if (act)
    for (int k = gi; k < cnt; k += gs) {
      const int j = substitute_from(rs.avm, (cnt + 31) >> 5, k);
      rs.sub[k] = j >= 0 ? rs.raw[j] : 128;
    }
  __syncthreads();
  if (act) {
    // [1 2 1] smoothing, or the bi-linear one for nearly flat 32x32 luma neighbourhoods (8.4.4.2.3)
    const int corner = rs.sub[2 * n];
    const bool strong = p == 0 && n == 32 && fp.strong_intra && abs(corner + rs.sub[4 * n] - 2 * rs.sub[3 * n]) < 8 &&
                        abs(corner + rs.sub[0] - 2 * rs.sub[n]) < 8;
    for (int k = gi; k < cnt; k += gs) {
      int v;
      if (k == 0 || k == cnt - 1 || (strong && k == 2 * n)) v = rs.sub[k];
      else if (strong) {
        const int i = k < 2 * n ? 2 * n - 1 - k : k - 2 * n - 1;
        v = ((63 - i) * corner + (i + 1) * rs.sub[k < 2 * n ? 0 : 4 * n] + 32) >> 6;
      } else v = (rs.sub[k - 1] + 2 * rs.sub[k] + rs.sub[k + 1] + 2) >> 2;
      rs.filt[k] = (uint8_t)v;
    }
    if (gi == gs - 1) {
      int sum = n;
      for (int i = 0; i < n; i++) sum += rs.sub[2 * n + 1 + i] + rs.sub[2 * n - 1 - i];
      rs.dc = sum >> (l2 + 1);
    }
  }
  __syncthreads();
  int16_t *a = sh.a + (p == 0 ? 0 : (p == 1 ? 1024 : 1280));
  int16_t *b = sh.b + (p == 0 ? 0 : (p == 1 ? 1024 : 1280));
  const int nshift = 5 - l2;
  int pr[4] = {0, 0, 0, 0};
  if (act) {
    const int qp = p == 0 ? qp_at(fp, lx, ly) : qp_c_at(fp, lx, ly, p);
    const int qper = qp / 6, lscale = c_level_scale[qp % 6], bd = l2 + 3;
    int s = 0;
    for (int i = gi; i < n * n; i += gs, s++) {
      const int y = i >> l2, x = i & (n - 1);
      pr[s] = intra_pixel(rs.sub, rs.filt, n, l2, mode, p, rs.dc, x, y);
      if (nz) {
        const int lvl = levels[poff + (size_t)(by + y) * pw + bx + x];
        long long d = ((long long)lvl * (sl_factor(fp.scaling, l2, p, x, y) * lscale)) << qper;
        d = (d + (1LL << (bd - 1))) >> bd;
        a[i] = (int16_t)max(-32768LL, min(32767LL, d));
      }
    }
  }
  __syncthreads();
  if (act && nz)
    for (int i = gi; i < n * n; i += gs) {
      const int y = i >> l2, x = i & (n - 1);
      int acc = 0;
      for (int k = 0; k < n; k++) acc += (dst ? c_dst4[k][y] : sh.dct[k << nshift][y]) * a[k * n + x];
      b[i] = (int16_t)clip3(-32768, 32767, (acc + 64) >> 7);
    }
  __syncthreads();
  if (act) {
    int s = 0;
    for (int i = gi; i < n * n; i += gs, s++) {
      const int y = i >> l2, x = i & (n - 1);
      int v = pr[s];
      if (nz) {
        int acc = 0;
        for (int k = 0; k < n; k++) acc += (dst ? c_dst4[k][x] : sh.dct[k << nshift][x]) * b[y * n + k];
        v = clip8(v + clip3(-32768, 32767, (acc + 2048) >> 12));
      }
      __stcg(rec + poff + (size_t)(by + y) * pw + bx + x, (uint8_t)v);
      tile[(by + y - ty0) * T + bx + x - tx0] = (uint8_t)v;
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kReconThreads)
k_intra_decode(FrameParams fp, uint8_t *rec, const int16_t *__restrict__ levels, const CuInfo *cu, int *ticket,
               const int *__restrict__ order)
{
  __shared__ DecSharedI sh;
  const int t = threadIdx.x;
  if (!fp.is_idr && *(volatile int *)fp.any_intra == 0) return;
  for (int i = t; i < 1024; i += kReconThreads) ((int8_t *)sh.dct)[i] = c_dct32[i >> 5][i & 31];
  const int nctu = fp.ctb_cols * fp.ctb_rows;
  const size_t ysz = (size_t)fp.w * fp.h;
  for (;;) {
    __syncthreads();
    if (t == 0) sh.ctu = atomicAdd(ticket, 1);
    __syncthreads();
    if (sh.ctu >= nctu) break;
    const int ctu = order[sh.ctu];
    const int row = ctu / fp.ctb_cols, col = ctu - row * fp.ctb_cols;
    const int cx = col * kCtb, cy = row * kCtb;
    int mine = 0;
    if (t < 64) {
      const int x8 = (cx >> 3) + z_to_x(t), y8 = (cy >> 3) + z_to_y(t);
      uint32_t w0 = 0, w1 = 0, w2 = 0, w3 = 0;
      if (x8 < fp.w8 && y8 < fp.h8) {
        const uint32_t *s = (const uint32_t *)(cu + (size_t)y8 * fp.w8 + x8);
        w0 = __ldcg(s); w1 = __ldcg(s + 1); w2 = __ldcg(s + 2); w3 = __ldcg(s + 3);
        mine = ((w1 >> 8) & 0xff) == 1 && (w1 & 0xff) != 0;
      } else {
        w1 = 0;
      }
      sh.cu[t][0] = w0; sh.cu[t][1] = w1; sh.cu[t][2] = w2; sh.cu[t][3] = w3;
    }
    const int has = __syncthreads_or(mine);
    if (has) {
      if (t == 0) {
        volatile int *d = fp.ctu_done;
        if (col > 0) while (d[ctu - 1] == 0) __nanosleep(32);
        if (row > 0) {
          if (col > 0) while (d[ctu - fp.ctb_cols - 1] == 0) __nanosleep(32);
          while (d[ctu - fp.ctb_cols] == 0) __nanosleep(32);
          if (col + 1 < fp.ctb_cols) while (d[ctu - fp.ctb_cols + 1] == 0) __nanosleep(32);
        }
        __threadfence();
      }
      __syncthreads();
      if (!fp.is_idr) {
        // the inter CUs of this CTU were reconstructed by the kernel before
        for (int i = t; i < 64 * 16; i += kReconThreads) {
          int y = cy + (i >> 4), x = cx + 4 * (i & 15);
          ((uint32_t *)sh.rec_y)[i] = (y < fp.h && x < fp.w) ? __ldcg((const uint32_t *)(rec + (size_t)y * fp.w + x)) : 0u;
        }
        for (int i = t; i < 2 * 32 * 8; i += kReconThreads) {
          int c = i >> 8, j = i & 255;
          int y = (cy >> 1) + (j >> 3), x = (cx >> 1) + 4 * (j & 7);
          const uint8_t *pl = rec + ysz + (c ? ysz / 4 : 0);
          ((uint32_t *)sh.rec_c[c])[j] = (y < (fp.h >> 1) && x < (fp.w >> 1)) ? __ldcg((const uint32_t *)(pl + (size_t)y * (fp.w >> 1) + x)) : 0u;
        }
      }
      load_ctu_border(sh.bd, fp, rec, cx, cy, t, kReconThreads);
      __syncthreads();
      for (int z = 0; z < 64; z++) {
        const CuInfo u = *(const CuInfo *)sh.cu[z];
        if (u.log2_size == 0 || u.pred_mode != 1) continue;
        const int ux = z_to_x(z), uy = z_to_y(z);
        const int x0 = cx + 8 * ux, y0 = cy + 8 * uy;
        if (u.tu_log2 >= 3) {
          const int n8 = 1 << (u.tu_log2 - 3);
          if ((ux & (n8 - 1)) || (uy & (n8 - 1))) continue;              // not the first unit of its transform unit
          dec_tu(sh, fp, rec, levels, cx, cy, x0, y0, u.tu_log2, u.intra_mode, true, x0 >> 1, y0 >> 1, u.tu_log2 - 1,
                 u.chroma_mode, u.cbf & 7);
        } else {
          // four 4x4 luma blocks (each with its own mode in an NxN CU), chroma alongside the first
          const bool nxn = u.flags & 1;
          for (int b = 0; b < 4; b++) {
            const int m = !nxn || b == 0 ? u.intra_mode : (b == 1 ? (u.mvx & 0xff) : (b == 2 ? ((u.mvx >> 8) & 0xff) : (u.mvy & 0xff)));
            dec_tu(sh, fp, rec, levels, cx, cy, x0 + 4 * (b & 1), y0 + 4 * (b >> 1), 2, m, b == 0, x0 >> 1, y0 >> 1, 2,
                   u.chroma_mode, ((u.cbf >> (4 + b)) & 1) | (u.cbf & 6));
          }
        }
      }
      __threadfence();
    }
    __syncthreads();
    if (t == 0) atomicExch(&fp.ctu_done[ctu], 1);
  }
}

}  // namespace

// CTU indices in wavefront order: sorted by (col + 2*row, row).  n = cols*rows entries.
void intra_wavefront_order(int cols, int rows, int *out)
{
  int k = 0;
  for (int w = 0; w <= (cols - 1) + 2 * (rows - 1); w++)
    for (int r = 0; r < rows; r++) {
      int c = w - 2 * r;
      if (c >= 0 && c < cols) out[k++] = r * cols + c;
    }
}

// I picture: wavefront width + slack, never more CTAs than CTUs.  P picture: most CTUs hold no intra
// CU and finish at once, and those that do wait for their four neighbours only, so a wider grid pays.
static int intra_grid(const FrameParams &fp)
{
  int width = std::min(fp.ctb_rows, (fp.ctb_cols + 1) / 2) + 3;
  if (!fp.is_idr) width = std::max(width, 148);
  return std::max(1, std::min(width, fp.ctb_cols * fp.ctb_rows));
}

// `ticket` = 1 int; fp.ctu_done = one int per CTU; both zeroed here.  I pictures: mode decision for the
// whole picture, then the reconstruction wavefront.
cudaError_t launch_intra_frame(const FrameParams &fp, const uint8_t *src, uint8_t *rec, int16_t *levels, CuInfo *cu,
                               int *ticket, const int *order, cudaStream_t s)
{
  cudaError_t e = cudaMemsetAsync(fp.ctu_done, 0, sizeof(int) * fp.ctb_cols * fp.ctb_rows, s);
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(ticket, 0, sizeof(int), s);
  if (e != cudaSuccess) return e;
  const int blocks16 = ((fp.w + 15) >> 4) * ((fp.h + 15) >> 4);
  k_intra_modes<<<blocks16, kModeThreads, 0, s>>>(fp, src, cu);
  k_intra_frame<<<intra_grid(fp), kReconThreads, 0, s>>>(fp, src, rec, levels, cu, ticket, order);
  return cudaGetLastError();
}

// P pictures: the intra CUs the motion search chose (cu map), after the inter reconstruction.  The
// kernel returns at once when *fp.any_intra is 0.
cudaError_t launch_intra_in_p(const FrameParams &fp, const uint8_t *src, uint8_t *rec, int16_t *levels, CuInfo *cu,
                              int *ticket, const int *order, cudaStream_t s)
{
  cudaError_t e = cudaMemsetAsync(fp.ctu_done, 0, sizeof(int) * fp.ctb_cols * fp.ctb_rows, s);
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(ticket, 0, sizeof(int), s);
  if (e != cudaSuccess) return e;
  k_intra_frame<<<intra_grid(fp), kReconThreads, 0, s>>>(fp, src, rec, levels, cu, ticket, order);
  return cudaGetLastError();
}

// Decoder reconstruction of the intra CUs of a picture (modes, cbf and levels from the parser): all
// CUs of an I picture; in a P picture the intra CUs, after launch_inter_decode.
cudaError_t launch_intra_decode(const FrameParams &fp, uint8_t *rec, const int16_t *levels, const CuInfo *cu,
                                int *ticket, const int *order, cudaStream_t s)
{
  cudaError_t e = cudaMemsetAsync(fp.ctu_done, 0, sizeof(int) * fp.ctb_cols * fp.ctb_rows, s);
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(ticket, 0, sizeof(int), s);
  if (e != cudaSuccess) return e;
  k_intra_decode<<<intra_grid(fp), kReconThreads, 0, s>>>(fp, rec, levels, cu, ticket, order);
  return cudaGetLastError();
}

}  // namespace b200
