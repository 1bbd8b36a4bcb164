// Device-side building blocks shared by the encoder and decoder kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "hevc_common.h"
#include "hevc_tables.cuh"

namespace b200 {

__device__ __forceinline__ int clip3(int lo, int hi, int v) { return min(max(v, lo), hi); }
__device__ __forceinline__ int clip8(int v) { return min(max(v, 0), 255); }

// Per-CTU quantiser (cu_qp_delta / ROI): tables the host otherwise folds into FrameParams.
static __constant__ uint16_t c_lambda_q4_tab[52] = {   // round(16 * sqrt(0.57 * 2^((qp-12)/3)))
  3, 3, 4, 4, 5, 5, 6, 7, 8, 9, 10, 11, 12, 14, 15, 17, 19, 22, 24, 27, 30, 34, 38, 43, 48, 54, 61, 68,
  77, 86, 97, 108, 122, 137, 153, 172, 193, 217, 244, 273, 307, 344, 387, 434, 487, 547, 614, 689, 773,
  868, 974, 1093};
static __constant__ uint8_t c_chroma_qp_tab[58] = {     // Table 8-10: qPi (0..57) -> QpC
  0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29,
  29, 30, 31, 32, 33, 33, 34, 34, 35, 35, 36, 36, 37, 37, 38, 39, 40, 41, 42, 43, 44, 45, 46, 47, 48, 49, 50, 51};
// luma QP / chroma QP / SAD-domain lambda of the CTU that holds luma sample (x, y)
__device__ __forceinline__ int qp_at(const FrameParams &fp, int x, int y)
{
  return fp.ctu_qp ? fp.ctu_qp[(y >> kCtbLog2) * fp.ctb_cols + (x >> kCtbLog2)] : fp.qp;
}
__device__ __forceinline__ int qp_c_at(const FrameParams &fp, int x, int y)
{
  return fp.ctu_qp ? c_chroma_qp_tab[qp_at(fp, x, y)] : fp.qp_c;
}
// ... of chroma plane c (1 Cb, 2 Cr) with the chroma QP offsets a foreign stream may carry (8.6.1)
__device__ __forceinline__ int qp_c_at(const FrameParams &fp, int x, int y, int c)
{
  const int off = c == 1 ? fp.cb_qp_offset : fp.cr_qp_offset;
  if (!fp.ctu_qp && off == 0) return fp.qp_c;
  return c_chroma_qp_tab[clip3(0, 57, qp_at(fp, x, y) + off)];
}
// Tile-column mode: may a block at x, n wide, use horizontal motion mvx (quarter samples)?  With a
// fractional luma or chroma position ((mvx & 7) != 0) the interpolation reaches up to 4 luma samples
// further on either side.
__device__ __forceinline__ bool mv_allowed(const FrameParams &fp, int x, int n, int mvx)
{
  const int ix = mvx >> 2, m = (mvx & 7) ? 4 : 0;
  if ((fp.mv_edges & 1) && x + ix - m < 0) return false;
  if ((fp.mv_edges & 2) && x + n + ix + m > fp.w) return false;
  return true;
}
// ... and vertical motion mvy for a block at y, n high (mv_edges bit 2 / 3: the top / bottom edge is an
// interior tile edge -- tile rows)
__device__ __forceinline__ bool mv_allowed_v(const FrameParams &fp, int y, int n, int mvy)
{
  const int iy = mvy >> 2, m = (mvy & 7) ? 4 : 0;
  if ((fp.mv_edges & 4) && y + iy - m < 0) return false;
  if ((fp.mv_edges & 8) && y + n + iy + m > fp.h) return false;
  return true;
}
__device__ __forceinline__ int lambda_q4_at(const FrameParams &fp, int x, int y)
{
  return fp.ctu_qp ? c_lambda_q4_tab[qp_at(fp, x, y)] : fp.lambda_q4;
}

// sum of absolute differences of four packed bytes, accumulated (one VABSDIFF4)
__device__ __forceinline__ unsigned sad4_acc(unsigned a, unsigned b, unsigned acc)
{
  unsigned d;
  asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(acc));
  return d;
}

// dot product of four unsigned bytes (a) with four signed bytes (b), accumulated
__device__ __forceinline__ int dp4a_us(unsigned a, unsigned b, int c)
{
  int d;
  asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

__device__ __forceinline__ unsigned pack4(const int8_t *p)
{
  return (unsigned)(uint8_t)p[0] | ((unsigned)(uint8_t)p[1] << 8) | ((unsigned)(uint8_t)p[2] << 16) |
         ((unsigned)(uint8_t)p[3] << 24);
}

// HM xGetComponentBits: exp-Golomb-like length of one mvd component
__device__ __forceinline__ int mv_comp_bits(int v)
{
  unsigned t = v <= 0 ? ((unsigned)(-v) << 1) + 1 : (unsigned)v << 1;
  return 2 * (31 - __clz(t)) + 1;
}
__device__ __forceinline__ unsigned mv_penalty(int lambda_q4, int mvx, int mvy)
{
  return (unsigned)((lambda_q4 * (mv_comp_bits(mvx) + mv_comp_bits(mvy))) >> 4);
}

// z-order index of an 8x8 unit inside its CTB <-> unit coordinates
__device__ __forceinline__ int z_to_x(int z) { return (z & 1) | ((z >> 1) & 2) | ((z >> 2) & 4); }
__device__ __forceinline__ int z_to_y(int z) { return ((z >> 1) & 1) | ((z >> 2) & 2) | ((z >> 3) & 4); }
__device__ __forceinline__ int xy_to_z(int x, int y)
{
  return (x & 1) | ((y & 1) << 1) | ((x & 2) << 1) | ((y & 2) << 2) | ((x & 4) << 2) | ((y & 4) << 3);
}

// Load a (ws x ws) window of a plane into shared memory as bytes, edge-clamped (8.5.3.3.3.1).
// Window pixel (wx,wy) = plane(clamp(x0+wx), clamp(y0+wy)).  x0 must be a multiple of 4 and
// the row pitch of the window (wsw words) >= ws/4.
__device__ __forceinline__ void load_window(const uint8_t *__restrict__ plane, int pw, int ph, int x0, int y0,
                                            int ws, int wsw, uint32_t *s_win)
{
  const int words = (ws + 3) >> 2;
  for (int i = threadIdx.x; i < ws * words; i += blockDim.x) {
    int wy = i / words, wi = i - wy * words;
    int y = clip3(0, ph - 1, y0 + wy), x = x0 + 4 * wi;
    const uint8_t *row = plane + (size_t)y * pw;
    uint32_t v;
    if (x >= 0 && x + 3 < pw) {
      v = __ldg((const uint32_t *)(row + x));
    } else {
      v = (uint32_t)row[clip3(0, pw - 1, x)] | ((uint32_t)row[clip3(0, pw - 1, x + 1)] << 8) |
          ((uint32_t)row[clip3(0, pw - 1, x + 2)] << 16) | ((uint32_t)row[clip3(0, pw - 1, x + 3)] << 24);
    }
    s_win[wy * wsw + wi] = v;
  }
}

// One row of a reference patch: `nwords` words holding plane bytes x .. x + 4 * nwords - 1 of row y,
// coordinates clamped to the picture (8.5.3.3.3.1), for any byte alignment of x.  Rows of the plane
// are word aligned (pw is a multiple of 4).
__device__ __forceinline__ void load_patch_row(const uint8_t *__restrict__ plane, int pw, int ph, int x, int y, int nwords,
                                               uint32_t *dst)
{
  const uint8_t *row = plane + (size_t)clip3(0, ph - 1, y) * pw;
  if (x >= 0 && (x & ~3) + 4 * (nwords + 1) <= pw) {
    const uint32_t *w = (const uint32_t *)(row + (x & ~3));
    const int sh = (x & 3) * 8;
    uint32_t a = __ldg(w);
    for (int i = 0; i < nwords; i++) {
      const uint32_t b = __ldg(w + i + 1);
      dst[i] = __funnelshift_r(a, b, sh);
      a = b;
    }
  } else {
    for (int i = 0; i < nwords; i++) {
      const int xx = x + 4 * i;
      dst[i] = (uint32_t)row[clip3(0, pw - 1, xx)] | ((uint32_t)row[clip3(0, pw - 1, xx + 1)] << 8) |
               ((uint32_t)row[clip3(0, pw - 1, xx + 2)] << 16) | ((uint32_t)row[clip3(0, pw - 1, xx + 3)] << 24);
    }
  }
}

// ---- luma motion compensation: 2 columns x 8 rows at window position (xi,yi), fractional
// phase (fx,fy).  Separable 8-tap, rows first (8.5.3.3.3); exact for zero phases as well
// because the zero-phase filter is {0,0,0,64,0,0,0,0}.  out[r] = pixel(col0) | pixel(col1) << 8.
__device__ __forceinline__ void mc_luma_2x8(const uint32_t *s_win, int wsw, int xi, int yi, int fx, int fy,
                                            unsigned out[8])
{
  const unsigned tl = pack4(c_luma_filter[fx]), th = pack4(c_luma_filter[fx] + 4);
  const int x0 = xi - 3;
  const int xw = x0 >> 2, sh = (x0 & 3) * 8;
  int h0[15], h1[15];
#pragma unroll
  for (int r = 0; r < 15; r++) {
    const uint32_t *row = s_win + (yi - 3 + r) * wsw + xw;
    unsigned a = row[0], b = row[1], c = row[2];
    unsigned p0 = __funnelshift_r(a, b, sh), p1 = __funnelshift_r(b, c, sh), p2 = c >> sh;
    unsigned q0 = __funnelshift_r(p0, p1, 8), q1 = __funnelshift_r(p1, p2, 8);
    h0[r] = dp4a_us(p0, tl, dp4a_us(p1, th, 0));
    h1[r] = dp4a_us(q0, tl, dp4a_us(q1, th, 0));
  }
  // vertical pass by DP2A: the row sums fit int16 (|sum| <= 88 * 255), so two of them and two
  // taps go through one instruction.  Pairs (r, r+1) are packed once for even and once for odd r.
  const unsigned t01 = pack4(c_luma_filter[fy]), t23 = t01 >> 16, t45 = pack4(c_luma_filter[fy] + 4), t67 = t45 >> 16;
  unsigned p0[14], p1[14];                 // p[r] = (h[r] & 0xffff) | (h[r+1] << 16)
#pragma unroll
  for (int r = 0; r < 14; r++) {
    p0[r] = __byte_perm((unsigned)h0[r], (unsigned)h0[r + 1], 0x5410);
    p1[r] = __byte_perm((unsigned)h1[r], (unsigned)h1[r + 1], 0x5410);
  }
#pragma unroll
  for (int r = 0; r < 8; r++) {
    int v0 = __dp2a_lo((int)p0[r], (int)t01, 0);
    v0 = __dp2a_lo((int)p0[r + 2], (int)t23, v0);
    v0 = __dp2a_lo((int)p0[r + 4], (int)t45, v0);
    v0 = __dp2a_lo((int)p0[r + 6], (int)t67, v0);
    int v1 = __dp2a_lo((int)p1[r], (int)t01, 0);
    v1 = __dp2a_lo((int)p1[r + 2], (int)t23, v1);
    v1 = __dp2a_lo((int)p1[r + 4], (int)t45, v1);
    v1 = __dp2a_lo((int)p1[r + 6], (int)t67, v1);
    v0 = clip8(((v0 >> 6) + 32) >> 6);
    v1 = clip8(((v1 >> 6) + 32) >> 6);
    out[r] = (unsigned)v0 | ((unsigned)v1 << 8);
  }
}

// ---- chroma motion compensation: 1 column x 4 rows, 4-tap, eighth-sample phases.
__device__ __forceinline__ void mc_chroma_1x4(const uint32_t *s_win, int wsw, int xi, int yi, int fx, int fy,
                                              unsigned out[4])
{
  const unsigned tp = pack4(c_chroma_filter[fx]);
  const int x0 = xi - 1;
  const int xw = x0 >> 2, sh = (x0 & 3) * 8;
  int hv[7];
#pragma unroll
  for (int r = 0; r < 7; r++) {
    const uint32_t *row = s_win + (yi - 1 + r) * wsw + xw;
    hv[r] = dp4a_us(__funnelshift_r(row[0], row[1], sh), tp, 0);
  }
#pragma unroll
  for (int r = 0; r < 4; r++) {
    int v = 0;
#pragma unroll
    for (int t = 0; t < 4; t++) v += c_chroma_filter[fy][t] * hv[r + t];
    out[r] = (unsigned)clip8(((v >> 6) + 32) >> 6);
  }
}

// ---- transform / quantisation pipeline over a CTU-shaped tile held in shared memory ---------
//
// Geometry: the tile is T x T samples (T = 64 luma, 32 chroma).  unit_log2 is the size of one
// cu-map unit in this plane (3 luma, 2 chroma).  s_org[z] is the z-index of the origin unit of the
// CU covering unit z (0xff = outside the picture); s_log2[z] the CU size (luma log2).  The
// transform block of every CU is the CU itself (luma) / half of it (chroma).
//
// All four 1-D passes are written as  out[a][b] = sum_i M[b][i] * in[a][i]  with the contraction
// index i contiguous in shared memory, so that two int16 samples and four int8 coefficients go
// through one DP2A: the int16 tiles have a row pitch of T+2 (33 words: column accesses hit 32
// different banks), every pass stores its result in the orientation the next pass contracts
// over, and lanes always run along the matrix index b, which makes the data reads broadcasts and
// the coefficient-word reads consecutive.
//
//   forward:  resid[y][x] --H--> tmpT[k][j] --V--> coef[v][k] -> level (natural) , deqT[k][v]
//   inverse:  deqT[x][v] --V--> tmp2[y][x] --H--> residual[y][x] -> reconstruction
struct TileGeom {
  int T;            // tile width (64 / 32)
  int tlog2;        // log2(T)
  int unit_log2;    // 3 / 2
  int chroma;       // 0 / 1: transform size = CU size >> chroma
};

struct TqParams {
  int qp;           // QP of this plane
  int is_idr;       // quantiser offset 171 (I slices) / 85
  const uint8_t *sl;   // FrameParams::scaling
  int matrix;       // scaling matrix of the plane (3 + plane: inter)
};

// Coefficient words.  wc[q][r] = { c32[r][4q .. 4q+3] }: row r of the 32-point matrix; the N-point
// transform uses rows k << (5 - log2 N) and q < N/4.  wct holds, per transform size, the words of
// the TRANSPOSED N-point matrix: wct[off(N) + q*N + y] = { C_N[4q .. 4q+3][y] }.
struct DctWords {
  uint32_t wc[8][32];
  uint32_t wct[4 + 16 + 64 + 256];
};
__device__ __forceinline__ int wct_offset(int log2n) { return log2n == 2 ? 0 : (log2n == 3 ? 4 : (log2n == 4 ? 20 : 84)); }

__device__ __forceinline__ void build_dct_words(DctWords &w)
{
  for (int i = threadIdx.x; i < 8 * 32; i += blockDim.x) {
    int q = i >> 5, r = i & 31;
    w.wc[q][r] = pack4(&c_dct32[r][4 * q]);
  }
  for (int i = threadIdx.x; i < 340; i += blockDim.x) {
    int log2n = i < 4 ? 2 : (i < 20 ? 3 : (i < 84 ? 4 : 5));
    int j = i - wct_offset(log2n), n = 1 << log2n, nshift = 5 - log2n;
    int q = j / n, y = j - q * n;
    int8_t c[4];
    for (int t = 0; t < 4; t++) c[t] = c_dct32[(4 * q + t) << nshift][y];
    w.wct[i] = pack4(c);
  }
}

// sum over n int16 samples (pairs in `data`, 4-byte aligned) times the coefficient words cw[q * stride]
__device__ __forceinline__ int dot_dp2a(const uint32_t *data, const uint32_t *cw, int stride, int n)
{
  int acc = 0;
  for (int q = 0; q < (n >> 2); q++) {
    const int c = (int)cw[q * stride];
    acc = __dp2a_lo((int)data[2 * q], c, acc);
    acc = __dp2a_hi((int)data[2 * q + 1], c, acc);
  }
  return acc;
}

// One sample position (x,y) of the tile -> its transform block; returns false outside the picture.
// `sub`: bit of s_nz[org] that says whether the block has levels -- 0, except for the four 4x4 luma
// blocks of an 8x8 unit whose transform unit is split once more (s_log2 == 2; decoder only), which
// share the unit's entry (their chroma is one 4x4 block per plane, as for any 8x8 unit).
struct TbPos { int ox, oy, n, log2n, org, sub; };
__device__ __forceinline__ bool tb_at(const TileGeom &g, const uint8_t *s_org, const uint8_t *s_log2, int x, int y, TbPos &tb)
{
  int z = xy_to_z(x >> g.unit_log2, y >> g.unit_log2);
  int o = s_org[z];
  if (o == 0xff) return false;
  tb.org = o;
  const int l2 = s_log2[o];
  tb.log2n = max(l2 - g.chroma, 2);
  tb.n = 1 << tb.log2n;
  tb.ox = z_to_x(o) << g.unit_log2;
  tb.oy = z_to_y(o) << g.unit_log2;
  tb.sub = 0;
  if (l2 == 2 && !g.chroma) {
    tb.ox += x & 4; tb.oy += y & 4;
    tb.sub = ((y >> 1) & 2) | ((x >> 2) & 1);
  }
  return true;
}

// Scaling factor m[x][y] (8.6.4.2) of the coefficient at column x, row y of a (1 << log2n)-sized block of
// matrix 0..2 (intra Y / Cb / Cr) or 3..5 (inter); `sl` = FrameParams::scaling (ScalingTable layout), null = flat
__device__ __forceinline__ int sl_factor(const uint8_t *sl, int log2n, int matrix, int x, int y)
{
  if (!sl) return 16;
  const int sid = log2n - 2;
  if (sid >= 2 && (x | y) == 0) return sl[1536 + (sid - 2) * 6 + matrix];
  const int s = max(sid - 1, 0);
  return sl[(sid * 6 + matrix) * 64 + ((y >> s) << (sid ? 3 : 2)) + (x >> s)];
}
// forward quantiser scale of that coefficient (HM: quantCoef = (quantScale << 4) / m)
__device__ __forceinline__ unsigned sl_quant_scale(const uint8_t *sl, int qrem, int log2n, int matrix, int x, int y)
{
  const unsigned s = (unsigned)c_quant_scale[qrem];
  return sl ? (s << 4) / (unsigned)sl_factor(sl, log2n, matrix, x, y) : s;
}

// dscale = m * levelScale[qp % 6] (m = 16 without scaling lists)
__device__ __forceinline__ int16_t dequant_level(int lvl, int log2n, int qper, int dscale)
{
  const int bd = log2n + 3;
  long long d = (((long long)lvl * dscale) << qper);
  d = (d + (1LL << (bd - 1))) >> bd;
  return (int16_t)max(-32768LL, min(32767LL, d));
}

// Forward path.  s_src / s_pred: uint8 tiles (pitch T).  s_a, s_b: int16 tiles of pitch T+2.
// On return s_b holds the levels (natural orientation), s_a the dequantised coefficients
// TRANSPOSED inside each block, s_nz[org] != 0 where the block has a non-zero level.  The caller
// must have zeroed s_nz and synchronised.
__device__ __forceinline__ void forward_tq(const TileGeom g, const TqParams q, const uint8_t *s_src, const uint8_t *s_pred,
                                           const uint8_t *s_org, const uint8_t *s_log2, const DctWords &w,
                                           int16_t *s_a, int16_t *s_b, int *s_nz)
{
  const int T = g.T, P = T + 2, total = T * T;
  for (int p = threadIdx.x; p < total; p += blockDim.x) {
    int y = p >> g.tlog2, x = p & (T - 1);
    s_a[y * P + x] = (int16_t)((int)s_src[p] - (int)s_pred[p]);
  }
  __syncthreads();
  // horizontal pass, lanes along k: tmpT[k][j] = (sum_i C[k][i] * resid[j][i] + rnd) >> (log2n - 1)
  for (int p = threadIdx.x; p < total; p += blockDim.x) {
    int y = p >> g.tlog2, x = p & (T - 1);
    TbPos tb;
    if (!tb_at(g, s_org, s_log2, x, y, tb)) continue;
    const int k = x - tb.ox, j = y - tb.oy, nshift = 5 - tb.log2n;
    const uint32_t *row = (const uint32_t *)(s_a + y * P + tb.ox);
    int acc = dot_dp2a(row, &w.wc[0][k << nshift], 32, tb.n);
    const int s1 = tb.log2n - 1;
    s_b[(tb.oy + k) * P + tb.ox + j] = (int16_t)((acc + (1 << (s1 - 1))) >> s1);
  }
  __syncthreads();
  // vertical pass, lanes along v: coef[v][k] = (sum_j C[v][j] * tmpT[k][j] + rnd) >> (log2n + 6), then Q and IQ
  const int qper = q.qp / 6, qrem = q.qp % 6;
  const int lscale = c_level_scale[qrem];
  int16_t vals[16];
  int cnt = 0;
  for (int p = threadIdx.x; p < total; p += blockDim.x, cnt++) {
    int x = p >> g.tlog2, y = p & (T - 1);           // transposed enumeration: consecutive lanes run down a column
    TbPos tb;
    vals[cnt] = 0;
    if (!tb_at(g, s_org, s_log2, x, y, tb)) continue;
    const int v = y - tb.oy, k = x - tb.ox, nshift = 5 - tb.log2n;
    const uint32_t *row = (const uint32_t *)(s_b + (tb.oy + k) * P + tb.ox);
    int acc = dot_dp2a(row, &w.wc[0][v << nshift], 32, tb.n);
    const int s2 = tb.log2n + 6;
    const int coef = (acc + (1 << (s2 - 1))) >> s2;
    const int qbits = 14 + qper + (7 - tb.log2n);
    const unsigned add = (unsigned)(q.is_idr ? 171 : 85) << (qbits - 9);
    const unsigned scale = sl_quant_scale(q.sl, qrem, tb.log2n, q.matrix, k, v);
    const unsigned a = ((unsigned)abs(coef) * scale + add) >> qbits;    // < 2^32: |coef| <= 2^15, scale < 2^15
    int lvl = (int)min(a, 32767u);
    if (coef < 0) lvl = -lvl;
    vals[cnt] = (int16_t)lvl;
    if (lvl) s_nz[tb.org] = 1;
    s_a[(tb.oy + k) * P + tb.ox + v] = dequant_level(lvl, tb.log2n, qper, sl_factor(q.sl, tb.log2n, q.matrix, k, v) * lscale);   // deqT[k][v]
  }
  __syncthreads();     // all reads of s_b (tmpT) are done; the levels may now overwrite it
  cnt = 0;
  for (int p = threadIdx.x; p < total; p += blockDim.x, cnt++) {
    int x = p >> g.tlog2, y = p & (T - 1);
    s_b[y * P + x] = vals[cnt];
  }
  __syncthreads();
}

// Inverse path: s_a holds dequantised coefficients transposed inside each block (deqT[x][v]);
// the reconstruction goes to s_rec (uint8 tile, pitch T).  s_t is an int16 scratch tile (pitch T+2).
// Blocks with s_nz[org]==0 copy the prediction.
__device__ __forceinline__ void inverse_recon(const TileGeom g, const uint8_t *s_pred, const uint8_t *s_org,
                                              const uint8_t *s_log2, const DctWords &w, const int16_t *s_a,
                                              int16_t *s_t, const int *s_nz, uint8_t *s_rec)
{
  const int T = g.T, P = T + 2, total = T * T;
  // vertical pass, lanes along y: tmp2[y][x] = clip16((sum_v C[v][y] * deqT[x][v] + 64) >> 7)
  for (int p = threadIdx.x; p < total; p += blockDim.x) {
    int x = p >> g.tlog2, y = p & (T - 1);
    TbPos tb;
    if (!tb_at(g, s_org, s_log2, x, y, tb) || !((s_nz[tb.org] >> tb.sub) & 1)) continue;
    const int yy = y - tb.oy, xx = x - tb.ox;
    const uint32_t *row = (const uint32_t *)(s_a + (tb.oy + xx) * P + tb.ox);
    int acc = dot_dp2a(row, &w.wct[wct_offset(tb.log2n) + yy], tb.n, tb.n);
    s_t[y * P + x] = (int16_t)clip3(-32768, 32767, (acc + 64) >> 7);
  }
  __syncthreads();
  // horizontal pass, lanes along x: res[y][x] = (sum_k C[k][x] * tmp2[y][k] + 2048) >> 12
  for (int p = threadIdx.x; p < total; p += blockDim.x) {
    int y = p >> g.tlog2, x = p & (T - 1);
    TbPos tb;
    if (!tb_at(g, s_org, s_log2, x, y, tb)) continue;
    int pr = s_pred[p];
    if ((s_nz[tb.org] >> tb.sub) & 1) {
      const int xx = x - tb.ox;
      const uint32_t *row = (const uint32_t *)(s_t + y * P + tb.ox);
      int acc = dot_dp2a(row, &w.wct[wct_offset(tb.log2n) + xx], tb.n, tb.n);
      pr = clip8(pr + clip3(-32768, 32767, (acc + 2048) >> 12));
    }
    s_rec[p] = (uint8_t)pr;
  }
  __syncthreads();
}

}  // namespace b200
