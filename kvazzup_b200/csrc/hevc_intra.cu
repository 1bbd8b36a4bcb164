// I-picture kernel of the B200 HEVC encoder (sm_100a).
//
// Intra prediction needs the reconstructed left / above / above-right neighbours, so CTUs run
// as a wavefront: one CTA per CTU, CTU indices handed out in raster order by an atomic ticket
// (so every CTU a CTA waits for is already running or done -- no deadlock), a per-row progress
// counter published with a release fence.  Inside a CTU the CUs (16x16, or 8x8 where 16 does
// not fit the picture) are coded in z-order by all 256 threads: one thread per sample for the
// 35-mode SAD search (H.265 8.4.4.2, rows K4/K10 of SURVEY.md 8a-K), then the transform /
// quantisation / reconstruction of the chosen mode (rows K5, K6) with chroma in derived mode.
#include "hevc_device.cuh"
#include "hevc_kernels.h"

namespace b200 {

namespace {

constexpr int kThreads = 256;

struct IntraShared {
  uint8_t raw[132], sub[132], filt[132], av[132];
  unsigned sad[36];
  int best_mode, dc, ctu, cand[3];
  int8_t dct[32][32], dctT[32][32];
  uint8_t src[256], pred[256];
  int16_t a[256], b[256];
};

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

// Gather + substitute (8.4.4.2.2) + filter (8.4.4.2.3) the neighbours of the n x n block at
// (x0,y0) of plane c.  Result in sh.sub / sh.filt, DC value in sh.dc.
__device__ void prepare_refs(IntraShared &sh, const FrameParams &fp, const uint8_t *rec_plane, int pw, int c,
                             int x0, int y0, int n, unsigned cur_order)
{
  const int t = threadIdx.x, cnt = 4 * n + 1, sft = c ? 1 : 0;
  if (t < cnt) {
    int x, y;
    if (t < 2 * n) { x = x0 - 1; y = y0 + 2 * n - 1 - t; }
    else if (t == 2 * n) { x = x0 - 1; y = y0 - 1; }
    else { x = x0 + (t - 2 * n - 1); y = y0 - 1; }
    int lx = x << sft, ly = y << sft;
    bool ok = lx >= 0 && ly >= 0 && lx < fp.w && ly < fp.h && coding_order_i(fp, lx, ly) < cur_order;
    sh.av[t] = ok;
    sh.raw[t] = ok ? __ldcg(rec_plane + (size_t)y * pw + x) : 0;
  }
  __syncthreads();
  if (t < cnt) {
    int j = t;
    while (j >= 0 && !sh.av[j]) j--;
    if (j < 0) { j = t + 1; while (j < cnt && !sh.av[j]) j++; }
    sh.sub[t] = j < cnt ? sh.raw[j] : 128;
  }
  __syncthreads();
  if (t < cnt)
    sh.filt[t] = (t == 0 || t == cnt - 1) ? sh.sub[t] : (uint8_t)((sh.sub[t - 1] + 2 * sh.sub[t] + sh.sub[t + 1] + 2) >> 2);
  if (t == 0) {
    int s = n;
    for (int i = 0; i < n; i++) s += sh.sub[2 * n + 1 + i] + sh.sub[2 * n - 1 - i];
    sh.dc = s >> (31 - __clz(n) + 1);
  }
  __syncthreads();
}

// residual -> DCT -> Q -> IQ -> IDCT -> reconstruction of one n x n block held in sh.src / sh.pred.
// Writes levels and reconstruction to HBM; returns (to all threads) whether any level is non-zero.
__device__ int tq_block(IntraShared &sh, const FrameParams &fp, int log2n, int qp, int16_t *lev_plane,
                        uint8_t *rec_plane, int pw, int x0, int y0)
{
  const int n = 1 << log2n, nn = n * n, t = threadIdx.x, nshift = 5 - log2n;
  const int y = t >> log2n, x = t & (n - 1);
  const bool act = t < nn;
  if (act) sh.a[t] = (int16_t)((int)sh.src[t] - (int)sh.pred[t]);
  __syncthreads();
  if (act) {
    int acc = 0, kk = x << nshift;
    for (int i = 0; i < n; i++) acc += sh.dctT[i][kk] * sh.a[y * n + i];
    int s1 = log2n - 1;
    sh.b[t] = (int16_t)((acc + (1 << (s1 - 1))) >> s1);
  }
  __syncthreads();
  int lvl = 0;
  if (act) {
    int acc = 0;
    const int8_t *c = sh.dct[y << nshift];
    for (int j = 0; j < n; j++) acc += c[j] * sh.b[j * n + x];
    int s2 = log2n + 6;
    int coef = (acc + (1 << (s2 - 1))) >> s2;
    int qper = qp / 6, qrem = qp % 6;
    int qbits = 14 + qper + (7 - log2n);
    unsigned add = (unsigned)(fp.is_idr ? 171 : 85) << (qbits - 9);
    unsigned a = ((unsigned)abs(coef) * (unsigned)c_quant_scale[qrem] + add) >> qbits;
    lvl = (int)min(a, 32767u);
    if (coef < 0) lvl = -lvl;
    lev_plane[(size_t)(y0 + y) * pw + x0 + x] = (int16_t)lvl;
    int bd = log2n + 3;
    long long d = ((long long)lvl * (16 * c_level_scale[qrem])) << qper;
    d = (d + (1LL << (bd - 1))) >> bd;
    sh.a[t] = (int16_t)max(-32768LL, min(32767LL, d));
  }
  int nz = __syncthreads_or(lvl != 0);
  if (nz) {
    if (act) {
      int acc = 0;
      for (int k = 0; k < n; k++) acc += sh.dct[k << nshift][y] * sh.a[k * n + x];
      sh.b[t] = (int16_t)clip3(-32768, 32767, (acc + 64) >> 7);
    }
    __syncthreads();
  }
  if (act) {
    int pr = sh.pred[t];
    if (nz) {
      int acc = 0;
      for (int k = 0; k < n; k++) acc += sh.dct[k << nshift][x] * sh.b[y * n + k];
      pr = clip8(pr + clip3(-32768, 32767, (acc + 2048) >> 12));
    }
    __stcg(rec_plane + (size_t)(y0 + y) * pw + x0 + x, (uint8_t)pr);
  }
  __syncthreads();
  return nz;
}

__device__ void intra_cu(IntraShared &sh, const FrameParams &fp, const uint8_t *src, uint8_t *rec, int16_t *levels,
                         CuInfo *cu, int x0, int y0, int log2)
{
  const int n = 1 << log2, t = threadIdx.x;
  const size_t ysz = (size_t)fp.w * fp.h;
  const unsigned cur = coding_order_i(fp, x0, y0);
  // most probable modes (8.4.2) -> mode signalling cost
  if (t == 0) {
    int a = 1, b = 1;
    if (x0 > 0) { const CuInfo *nb = &cu[(size_t)(y0 >> 3) * fp.w8 + ((x0 - 1) >> 3)]; if (__ldcg(&nb->pred_mode) == 1) a = __ldcg(&nb->intra_mode); }
    if (y0 > 0 && (y0 & (kCtb - 1))) { const CuInfo *nb = &cu[(size_t)((y0 - 1) >> 3) * fp.w8 + (x0 >> 3)]; if (__ldcg(&nb->pred_mode) == 1) b = __ldcg(&nb->intra_mode); }
    if (a == b) {
      if (a < 2) { sh.cand[0] = 0; sh.cand[1] = 1; sh.cand[2] = 26; }
      else { sh.cand[0] = a; sh.cand[1] = 2 + ((a + 29) % 32); sh.cand[2] = 2 + ((a - 2 + 1) % 32); }
    } else {
      sh.cand[0] = a; sh.cand[1] = b;
      sh.cand[2] = (a != 0 && b != 0) ? 0 : ((a != 1 && b != 1) ? 1 : 26);
    }
  }
  if (t < 36) sh.sad[t] = 0;
  const int y = t >> log2, x = t & (n - 1);
  const bool act = t < n * n;
  if (act) sh.src[t] = __ldg(src + (size_t)(y0 + y) * fp.w + x0 + x);
  prepare_refs(sh, fp, rec, fp.w, 0, x0, y0, n, cur);
  // 35-mode search, one thread per sample
  for (int mode = 0; mode < 35; mode++) {
    unsigned d = 0;
    if (act) d = (unsigned)abs((int)sh.src[t] - intra_pixel(sh.sub, sh.filt, n, log2, mode, 0, sh.dc, x, y));
    d = __reduce_add_sync(0xffffffffu, d);
    if ((t & 31) == 0 && d) atomicAdd(&sh.sad[mode], d);
  }
  __syncthreads();
  if (t == 0) {
    unsigned best = 0xffffffffu;
    int bm = 0;
    for (int mode = 0; mode < 35; mode++) {
      int bits = mode == sh.cand[0] ? 2 : ((mode == sh.cand[1] || mode == sh.cand[2]) ? 3 : 6);
      unsigned cost = sh.sad[mode] + (unsigned)((fp.lambda_q4 * bits) >> 4);
      if (cost < best) { best = cost; bm = mode; }
    }
    sh.best_mode = bm;
  }
  __syncthreads();
  const int mode = sh.best_mode;
  if (act) sh.pred[t] = (uint8_t)intra_pixel(sh.sub, sh.filt, n, log2, mode, 0, sh.dc, x, y);
  __syncthreads();
  int cbf = tq_block(sh, fp, log2, fp.qp, levels, rec, fp.w, x0, y0) ? 1 : 0;
  // chroma, derived mode
  const int nc = n >> 1, lc = log2 - 1, cw = fp.w >> 1;
  for (int c = 1; c < 3; c++) {
    const size_t off = ysz + (c == 2 ? ysz / 4 : 0);
    const int yc = t >> lc, xc = t & (nc - 1);
    const bool actc = t < nc * nc;
    if (actc) sh.src[t] = __ldg(src + off + (size_t)(y0 / 2 + yc) * cw + x0 / 2 + xc);
    prepare_refs(sh, fp, rec + off, cw, c, x0 / 2, y0 / 2, nc, cur);
    if (actc) sh.pred[t] = (uint8_t)intra_pixel(sh.sub, sh.filt, nc, lc, mode, c, sh.dc, xc, yc);
    __syncthreads();
    if (tq_block(sh, fp, lc, fp.qp_c, levels + off, rec + off, cw, x0 / 2, y0 / 2)) cbf |= 1 << c;
  }
  const int n8 = n >> 3;
  if (t < n8 * n8) {
    CuInfo ci;
    ci.mvx = 0; ci.mvy = 0; ci.log2_size = (uint8_t)log2; ci.pred_mode = 1; ci.intra_mode = (uint8_t)mode;
    ci.cbf = (uint8_t)cbf; ci.skip = 0; ci.merge_idx = 0xff; ci.mvp_idx = 0; ci.pad = 0;
    CuInfo *dst = &cu[(size_t)((y0 >> 3) + t / n8) * fp.w8 + (x0 >> 3) + t % n8];
    uint32_t *d32 = (uint32_t *)dst;
    const uint32_t *s32 = (const uint32_t *)&ci;
    __stcg(d32, s32[0]); __stcg(d32 + 1, s32[1]); __stcg(d32 + 2, s32[2]);
  }
  __threadfence();
  __syncthreads();
}

__device__ void intra_tree(IntraShared &sh, const FrameParams &fp, const uint8_t *src, uint8_t *rec, int16_t *levels,
                           CuInfo *cu, int cx, int cy)
{
  // z-order walk over the sixteen 16x16 positions of the CTU; 16x16 that cross the picture edge fall to 8x8
  for (int z16 = 0; z16 < 16; z16++) {
    int x0 = cx + 16 * ((z16 & 1) | ((z16 >> 1) & 2)), y0 = cy + 16 * (((z16 >> 1) & 1) | ((z16 >> 2) & 2));
    if (x0 >= fp.w || y0 >= fp.h) continue;
    if (x0 + 16 <= fp.w && y0 + 16 <= fp.h) {
      intra_cu(sh, fp, src, rec, levels, cu, x0, y0, 4);
    } else {
      for (int q = 0; q < 4; q++) {
        int x1 = x0 + 8 * (q & 1), y1 = y0 + 8 * (q >> 1);
        if (x1 < fp.w && y1 < fp.h) intra_cu(sh, fp, src, rec, levels, cu, x1, y1, 3);
      }
    }
  }
}

__global__ void __launch_bounds__(kThreads)
k_intra_frame(FrameParams fp, const uint8_t *__restrict__ src, uint8_t *rec, int16_t *levels, CuInfo *cu,
              int *progress, int *ticket)
{
  __shared__ IntraShared sh;
  const int t = threadIdx.x;
  for (int i = t; i < 1024; i += kThreads) {
    ((int8_t *)sh.dct)[i] = c_dct32[i >> 5][i & 31];
    ((int8_t *)sh.dctT)[i] = c_dct32[i & 31][i >> 5];
  }
  if (t == 0) sh.ctu = atomicAdd(ticket, 1);
  __syncthreads();
  const int ctu = sh.ctu;
  const int row = ctu / fp.ctb_cols, col = ctu - row * fp.ctb_cols;
  if (t == 0) {
    // left CTU of this row, and the above-right CTU of the row above
    volatile int *p = progress;
    while (col > 0 && p[row] < col) __nanosleep(64);
    if (row > 0) {
      int need = min(col + 2, fp.ctb_cols);
      while (p[row - 1] < need) __nanosleep(64);
    }
    __threadfence();
  }
  __syncthreads();
  intra_tree(sh, fp, src, rec, levels, cu, col * kCtb, row * kCtb);
  __threadfence();
  __syncthreads();
  if (t == 0) atomicExch(&progress[row], col + 1);
}

}  // namespace

cudaError_t launch_intra_frame(const FrameParams &fp, const uint8_t *src, uint8_t *rec, int16_t *levels, CuInfo *cu,
                               int *progress, int *ticket, cudaStream_t s)
{
  cudaError_t e = cudaMemsetAsync(progress, 0, sizeof(int) * fp.ctb_rows, s);
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(ticket, 0, sizeof(int), s);
  if (e != cudaSuccess) return e;
  k_intra_frame<<<fp.ctb_cols * fp.ctb_rows, kThreads, 0, s>>>(fp, src, rec, levels, cu, progress, ticket);
  return cudaGetLastError();
}

}  // namespace b200
