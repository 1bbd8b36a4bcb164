// Intra prediction building blocks shared by the I-picture kernels (hevc_intra.cu) and the intra
// candidates of P pictures inside the motion-search kernel (hevc_inter.cu); H.265 8.4.4.2.
#pragma once
#include "hevc_device.cuh"

namespace b200 {

// neighbour samples of one block: raw gather, substituted, [1 2 1]-filtered, availability
struct RefSet { uint8_t raw[68], sub[68], filt[68], av[68]; unsigned avm[4]; int dc; };      // avm: availability as bit masks (32 entries per word)

// 8.4.4.2.2 in constant time: index of the entry that entry t takes its value from -- the nearest
// available one at or below t, else the lowest available one (the search "from the bottom-left up");
// -1 when nothing is available.  `avm` = availability bits, `words` of them.
__device__ __forceinline__ int substitute_from(const unsigned *avm, int words, int t)
{
  const int w = t >> 5;
  unsigned m = avm[w] & (0xffffffffu >> (31 - (t & 31)));
  for (int k = w; ; ) {
    if (m) return 32 * k + 31 - __clz(m);
    if (--k < 0) break;
    m = avm[k];
  }
  for (int k = 0; k < words; k++)
    if (avm[k]) return 32 * k + __ffs(avm[k]) - 1;
  return -1;
}

__device__ __forceinline__ unsigned coding_order_i(const FrameParams &fp, int x, int y)
{
  return (unsigned)((y >> kCtbLog2) * fp.ctb_cols + (x >> kCtbLog2)) * 64u + (unsigned)xy_to_z((x >> 3) & 7, (y >> 3) & 7);
}

// prediction of sample (x,y) of an n x n block, H.265 8.4.4.2.4-6.  `u` = substituted
// neighbours, `f` = their [1 2 1]-filtered version; layout: [0..2n-1] left column from the
// bottom up, [2n] corner, [2n+1..4n] top row.
__device__ __forceinline__ int intra_pixel(const uint8_t *u, const uint8_t *f, int n, int log2n, int mode, int cidx,
                                           int dc, int x, int y)
{
  const uint8_t *r = u;
  if (cidx == 0 && mode != 1 && n != 4) {
    int d = min(abs(mode - 26), abs(mode - 10));
    int thres = n == 8 ? 7 : (n == 16 ? 1 : 0);
    if (d > thres) r = f;
  }
#define LEFT(yy) r[2 * n - 1 - (yy)]
#define TOP(xx) r[2 * n + 1 + (xx)]
  if (mode == 0)
    return ((n - 1 - x) * LEFT(y) + (x + 1) * TOP(n) + (n - 1 - y) * TOP(x) + (y + 1) * LEFT(n) + n) >> (log2n + 1);
  if (mode == 1) {
    if (cidx == 0 && n < 32) {
      if (x == 0 && y == 0) return (LEFT(0) + 2 * dc + TOP(0) + 2) >> 2;
      if (y == 0) return (TOP(x) + 3 * dc + 2) >> 2;
      if (x == 0) return (LEFT(y) + 3 * dc + 2) >> 2;
    }
    return dc;
  }
  const int angle = c_intra_angle[mode], inv = c_inv_angle[mode];
  const bool vert = mode >= 18;
  const int a = vert ? x : y, b = vert ? y : x;          // a runs along the reference, b away from it
  const int idx = ((b + 1) * angle) >> 5, fact = ((b + 1) * angle) & 31;
  int i0 = a + idx + 1;
  // ref[i]: i >= 0 -> main side sample i-1 (i = 0 is the corner); i < 0 -> projected side sample
  auto ref_at = [&](int i) -> int {
    if (i >= 0) return vert ? TOP(i - 1) : LEFT(i - 1);
    int s = -1 + ((i * inv + 128) >> 8);
    return vert ? LEFT(s) : TOP(s);
  };
  int v = fact ? ((32 - fact) * ref_at(i0) + fact * ref_at(i0 + 1) + 16) >> 5 : ref_at(i0);
  if (cidx == 0 && n < 32 && a == 0 && angle == 0) {       // modes 26 / 10: first column / row smoothing
    if (vert) v = clip8(TOP(0) + ((LEFT(b) - LEFT(-1)) >> 1));
    else v = clip8(LEFT(0) + ((TOP(b) - TOP(-1)) >> 1));
  }
  return v;
#undef LEFT
#undef TOP
}

// Gather (8.4.4.2.2) the 4n+1 neighbours of the n x n block at (x0,y0) of `plane` (plane c of the
// picture; c > 0 is 4:2:0 chroma) into rs.raw / rs.av.  `first` = index of the first thread of
// the group of >= 4n+1 threads doing this block.  Caller synchronises, then calls finish_refs.
// `tile` (may be NULL): the current CTU's reconstruction of this plane in shared memory, T x T
// samples whose top-left is plane sample (tx0,ty0); neighbours inside it are read from there
// (no L2 round trip on the wavefront's critical path).  `top` / `left` (with `tile`): the row above the
// CTU from x = tx0 - 1 (1 + 2T samples) and the column left of it (T samples), fetched once per CTU --
// every available neighbour outside the tile lies there, so nothing is read from HBM / L2 per CU.
__device__ __forceinline__ void gather_refs(RefSet &rs, const FrameParams &fp, const uint8_t *plane, int pw, int c,
                                            int x0, int y0, int n, unsigned cur_order, int t,
                                            const uint8_t *tile = nullptr, int T = 0, int tx0 = 0, int ty0 = 0,
                                            const uint8_t *top = nullptr, const uint8_t *left = nullptr)
{
  const int cnt = 4 * n + 1, sft = c ? 1 : 0;
  bool avail = false;
  if (t >= 0 && t < cnt) {
    int x, y;
    if (t < 2 * n) { x = x0 - 1; y = y0 + 2 * n - 1 - t; }
    else if (t == 2 * n) { x = x0 - 1; y = y0 - 1; }
    else { x = x0 + (t - 2 * n - 1); y = y0 - 1; }
    int lx = x << sft, ly = y << sft;
    bool ok = lx >= 0 && ly >= 0 && lx < fp.w && ly < fp.h && coding_order_i(fp, lx, ly) < cur_order;
    rs.av[t] = ok;
    avail = ok;
    uint8_t v = 0;
    if (ok) {
      const int lx2 = x - tx0, ly2 = y - ty0;
      if (tile && lx2 >= 0 && ly2 >= 0 && lx2 < T && ly2 < T) v = tile[ly2 * T + lx2];
      else if (top && ly2 == -1 && lx2 >= -1 && lx2 < 2 * T) v = top[lx2 + 1];
      else if (left && lx2 == -1 && ly2 >= 0 && ly2 < T) v = left[ly2];
      else v = __ldcg(plane + (size_t)y * pw + x);
    }
    rs.raw[t] = v;
  }
  // availability bits, one word per warp of the group (whole warps call this: t runs over a multiple of 32)
  const unsigned bal = __ballot_sync(0xffffffffu, avail);
  if (t >= 0 && t < 128 && (t & 31) == 0) rs.avm[t >> 5] = bal;
}
// substitution (8.4.4.2.2) by thread t of the group; caller synchronises afterwards
__device__ __forceinline__ void substitute_refs(RefSet &rs, int n, int t)
{
  const int cnt = 4 * n + 1;
  if (t >= 0 && t < cnt) {
    const int j = substitute_from(rs.avm, (cnt + 31) >> 5, t);
    rs.sub[t] = j >= 0 ? rs.raw[j] : 128;
  }
}
// [1 2 1] smoothing (8.4.4.2.3) and the DC value; caller synchronises afterwards
__device__ __forceinline__ void filter_refs(RefSet &rs, int n, int t)
{
  const int cnt = 4 * n + 1;
  if (t >= 0 && t < cnt)
    rs.filt[t] = (t == 0 || t == cnt - 1) ? rs.sub[t] : (uint8_t)((rs.sub[t - 1] + 2 * rs.sub[t] + rs.sub[t + 1] + 2) >> 2);
  if (t == cnt) {                      // one spare thread of the group sums the DC value
    int s = n;
    for (int i = 0; i < n; i++) s += rs.sub[2 * n + 1 + i] + rs.sub[2 * n - 1 - i];
    rs.dc = s >> (31 - __clz(n) + 1);
  }
}

// ---- 35-mode search of one CU from SOURCE neighbours, by a group of 256 threads (thread t <->
// sample t of the CU).  Result: sh.best_mode / sh.best_cost = SAD + lambda * mode bits (fixed prior:
// the MPM list is unknown in a parallel pass).  All threads of the CTA must call it (barriers).
// The border of a CTU for gather_refs: the row above it (from x = tx0 - 1, 1 + 2T samples) and the column
// left of it (T samples) of the three planes.  Loaded by all threads of the CTA after the CTU's dependencies
// are complete; samples outside the picture are not touched (they are never available).
struct CtuBorder { uint8_t top_y[132], left_y[64], top_c[2][68], left_c[2][32]; };

__device__ __forceinline__ void load_ctu_border(CtuBorder &b, const FrameParams &fp, const uint8_t *rec, int cx, int cy, int t, int nthreads)
{
  const size_t ysz = (size_t)fp.w * fp.h;
  const int cw = fp.w >> 1, chh = fp.h >> 1, ccx = cx >> 1, ccy = cy >> 1;
  for (int i = t; i < 129 + 64 + 2 * (65 + 32); i += nthreads) {
    if (i < 129) {                                      // luma, row above
      const int x = cx - 1 + i;
      if (cy > 0 && x >= 0 && x < fp.w) b.top_y[i] = __ldcg(rec + (size_t)(cy - 1) * fp.w + x);
    } else if (i < 193) {                               // luma, column to the left
      const int y = cy + i - 129;
      if (cx > 0 && y < fp.h) b.left_y[i - 129] = __ldcg(rec + (size_t)y * fp.w + cx - 1);
    } else {
      const int j = i - 193, c = j / 97, k = j - c * 97;
      const uint8_t *pl = rec + ysz + (c ? ysz / 4 : 0);
      if (k < 65) {
        const int x = ccx - 1 + k;
        if (ccy > 0 && x >= 0 && x < cw) b.top_c[c][k] = __ldcg(pl + (size_t)(ccy - 1) * cw + x);
      } else {
        const int y = ccy + k - 65;
        if (ccx > 0 && y < chh) b.left_c[c][k - 65] = __ldcg(pl + (size_t)y * cw + ccx - 1);
      }
    }
  }
}

struct ModeShared {
  RefSet rs;
  unsigned sad[36];
  uint8_t src[256];
  unsigned best_cost;
  int best_mode;
};

// `resid` (may be null): two buffers of 256 residuals.  Given, the modes are compared by the Hadamard SATD
// of the residual over 8x8 tiles, (sum |h| + 2) >> 2 per tile as oracle/hevc_prims.c orc_satd (SURVEY.md
// 8a-K row K2), instead of the SAD: the residuals of a mode go through shared memory to one warp per tile,
// two rows of the tile per lane (rows r and r + 4), so that the first vertical butterfly is in registers
// and the other five stages are warp shuffles.
__device__ __forceinline__ void intra_search_cu(ModeShared &sh, const FrameParams &fp, const uint8_t *src, int x0, int y0, int log2,
                                                int16_t (*resid)[256] = nullptr)
{
  const int n = 1 << log2, t = threadIdx.x;
  const unsigned cur = coding_order_i(fp, x0, y0);
  const int y = t >> log2, x = t & (n - 1);
  const bool act = t < n * n;
  if (t < 36) sh.sad[t] = 0;
  if (act) sh.src[t] = __ldg(src + (size_t)(y0 + y) * fp.w + x0 + x);
  gather_refs(sh.rs, fp, src, fp.w, 0, x0, y0, n, cur, t);
  __syncthreads();
  substitute_refs(sh.rs, n, t);
  __syncthreads();
  filter_refs(sh.rs, n, t);
  __syncthreads();
  if (resid) {
    const int tiles_log2 = log2 - 3, ntile = 1 << (2 * tiles_log2);
    const int lane = t & 31, tile = t >> 5;
    const int tx = (tile & ((1 << tiles_log2) - 1)) * 8 + (lane & 7), ty = (tile >> tiles_log2) * 8 + (lane >> 3);
    for (int mode = 0; mode < 35; mode++) {
      int16_t *d = resid[mode & 1];
      if (act) d[t] = (int16_t)((int)sh.src[t] - intra_pixel(sh.rs.sub, sh.rs.filt, n, log2, mode, 0, sh.rs.dc, x, y));
      __syncthreads();                                   // (the buffer of mode - 1 is still being read: two buffers)
      if (tile < ntile) {
        const int v0 = d[(ty << log2) + tx], v1 = d[((ty + 4) << log2) + tx];
        int a = v0 + v1, b = v0 - v1;
#pragma unroll
        for (int m = 8; m != 4; m = m == 16 ? 1 : m << 1) {          // lane bits 3, 4 (rows), then 0, 1, 2 (columns)
          const int oa = __shfl_xor_sync(0xffffffffu, a, m), ob = __shfl_xor_sync(0xffffffffu, b, m);
          a = (lane & m) ? oa - a : a + oa;
          b = (lane & m) ? ob - b : b + ob;
        }
        {
          const int oa = __shfl_xor_sync(0xffffffffu, a, 4), ob = __shfl_xor_sync(0xffffffffu, b, 4);
          a = (lane & 4) ? oa - a : a + oa;
          b = (lane & 4) ? ob - b : b + ob;
        }
        const unsigned sum = __reduce_add_sync(0xffffffffu, (unsigned)(abs(a) + abs(b)));
        if (lane == 0) atomicAdd(&sh.sad[mode], (sum + 2) >> 2);
      }
    }
  } else {
    for (int mode = 0; mode < 35; mode++) {
      unsigned d = 0;
      if (act) d = (unsigned)abs((int)sh.src[t] - intra_pixel(sh.rs.sub, sh.rs.filt, n, log2, mode, 0, sh.rs.dc, x, y));
      d = __reduce_add_sync(0xffffffffu, d);
      if ((t & 31) == 0 && d) atomicAdd(&sh.sad[mode], d);
    }
  }
  __syncthreads();
  if (t == 0) {
    unsigned best = 0xffffffffu;
    int bm = 0;
    for (int mode = 0; mode < 35; mode++) {
      int bits = (mode == 0 || mode == 1 || mode == 26) ? 2 : 6;
      unsigned cost = sh.sad[mode] + (unsigned)((lambda_q4_at(fp, x0, y0) * bits) >> 4);
      if (cost < best) { best = cost; bm = mode; }
    }
    sh.best_cost = best; sh.best_mode = bm;
  }
  __syncthreads();
}

}  // namespace b200
