/* TEST INFRASTRUCTURE ONLY -- scalar restatement of the HEVC block primitives.
 * Written from the ITU-T H.265 text (clause numbers cited per function) and, where the
 * standard is silent (forward transform, quantiser rounding, SATD), from the HM
 * reference-software conventions that Kvazaar's strategy functions also follow
 * (SURVEY.md 8a-K rows K1-K7).  See hevc_tables.h for the pinning status. */
#include "hevc_prims.h"
#include "hevc_tables.h"

#include <stdlib.h>
#include <string.h>

static inline int clip3(int lo, int hi, int v) { return v < lo ? lo : (v > hi ? hi : v); }
static inline int clip8(int v) { return clip3(0, 255, v); }

/* ---- K1 / K2 -------------------------------------------------------------------------- */

uint32_t orc_sad(const uint8_t *a, int sa, const uint8_t *b, int sb, int w, int h)
{
  uint32_t s = 0;
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) s += (uint32_t)abs((int)a[y * sa + x] - (int)b[y * sb + x]);
  return s;
}

static uint32_t had4(const uint8_t *a, int sa, const uint8_t *b, int sb)
{
  int d[16], m[16];
  for (int y = 0; y < 4; y++)
    for (int x = 0; x < 4; x++) d[y * 4 + x] = (int)a[y * sa + x] - (int)b[y * sb + x];
  for (int y = 0; y < 4; y++) {      /* rows */
    int s0 = d[y * 4] + d[y * 4 + 1], s1 = d[y * 4 + 2] + d[y * 4 + 3];
    int t0 = d[y * 4] - d[y * 4 + 1], t1 = d[y * 4 + 2] - d[y * 4 + 3];
    m[y * 4] = s0 + s1; m[y * 4 + 1] = t0 + t1; m[y * 4 + 2] = s0 - s1; m[y * 4 + 3] = t0 - t1;
  }
  uint32_t sum = 0;
  for (int x = 0; x < 4; x++) {      /* columns */
    int s0 = m[x] + m[4 + x], s1 = m[8 + x] + m[12 + x];
    int t0 = m[x] - m[4 + x], t1 = m[8 + x] - m[12 + x];
    sum += abs(s0 + s1) + abs(t0 + t1) + abs(s0 - s1) + abs(t0 - t1);
  }
  return (sum + 1) >> 1;
}

static uint32_t had8(const uint8_t *a, int sa, const uint8_t *b, int sb)
{
  int m[64];
  for (int y = 0; y < 8; y++)
    for (int x = 0; x < 8; x++) m[y * 8 + x] = (int)a[y * sa + x] - (int)b[y * sb + x];
  for (int pass = 0; pass < 2; pass++) {
    int step = pass ? 8 : 1, line = pass ? 1 : 8;
    for (int l = 0; l < 8; l++) {
      int *p = m + l * line;
      for (int len = 1; len < 8; len <<= 1)        /* 3 butterfly stages */
        for (int i = 0; i < 8; i += 2 * len)
          for (int j = i; j < i + len; j++) {
            int u = p[j * step], v = p[(j + len) * step];
            p[j * step] = u + v; p[(j + len) * step] = u - v;
          }
    }
  }
  uint32_t sum = 0;
  for (int i = 0; i < 64; i++) sum += (uint32_t)abs(m[i]);
  return (sum + 2) >> 2;
}

uint32_t orc_satd(const uint8_t *a, int sa, const uint8_t *b, int sb, int w, int h)
{
  uint32_t s = 0;
  if ((w & 7) || (h & 7)) {
    for (int y = 0; y < h; y += 4)
      for (int x = 0; x < w; x += 4) s += had4(a + y * sa + x, sa, b + y * sb + x, sb);
  } else {
    for (int y = 0; y < h; y += 8)
      for (int x = 0; x < w; x += 8) s += had8(a + y * sa + x, sa, b + y * sb + x, sb);
  }
  return s;
}

/* ---- K5: transforms ---------------------------------------------------------------------- */

/* Forward: HM TComTrQuant::xTrMxN -- horizontal pass (shift log2N + bitDepth - 9), then
 * vertical pass (shift log2N + 6), round half up; matrix form of the partial butterflies. */
static void fwd_2d(const int16_t *src, int16_t *dst, int n, int log2n, int (*coef)(int, int, int))
{
  int tmp[32 * 32];
  int s1 = log2n - 1, s2 = log2n + 6;
  for (int j = 0; j < n; j++)             /* row j of the residual */
    for (int k = 0; k < n; k++) {
      int acc = 0;
      for (int i = 0; i < n; i++) acc += coef(n, k, i) * src[j * n + i];
      tmp[j * n + k] = s1 > 0 ? (acc + (1 << (s1 - 1))) >> s1 : acc;
    }
  for (int k = 0; k < n; k++)             /* column k of tmp */
    for (int v = 0; v < n; v++) {
      int acc = 0;
      for (int j = 0; j < n; j++) acc += coef(n, v, j) * tmp[j * n + k];
      dst[v * n + k] = (int16_t)((acc + (1 << (s2 - 1))) >> s2);
    }
}

/* Inverse: H.265 8.6.4.2 -- vertical pass, clip to 16 bit after (x+64)>>7, horizontal
 * pass, then (x + (1<<(bdShift-1))) >> bdShift with bdShift = 20 - bitDepth = 12. */
static void inv_2d(const int16_t *src, int16_t *dst, int n, int (*coef)(int, int, int))
{
  int tmp[32 * 32];
  for (int x = 0; x < n; x++)
    for (int y = 0; y < n; y++) {
      int acc = 0;
      for (int k = 0; k < n; k++) acc += coef(n, k, y) * src[k * n + x];
      tmp[y * n + x] = clip3(-32768, 32767, (acc + 64) >> 7);
    }
  for (int y = 0; y < n; y++)
    for (int x = 0; x < n; x++) {
      int acc = 0;
      for (int k = 0; k < n; k++) acc += coef(n, k, x) * tmp[y * n + k];
      dst[y * n + x] = (int16_t)clip3(-32768, 32767, (acc + 2048) >> 12);
    }
}

static int dst_coef(int n, int k, int i) { (void)n; return orc_dst4[k][i]; }

void orc_fdct(const int16_t *r, int16_t *c, int log2n) { fwd_2d(r, c, 1 << log2n, log2n, orc_dct_coef); }
void orc_idct(const int16_t *c, int16_t *r, int log2n) { inv_2d(c, r, 1 << log2n, orc_dct_coef); }
void orc_fdst4(const int16_t *r, int16_t *c) { fwd_2d(r, c, 4, 2, dst_coef); }
void orc_idst4(const int16_t *c, int16_t *r) { inv_2d(c, r, 4, dst_coef); }

/* ---- K6: quantisation ------------------------------------------------------------------------ */

int orc_chroma_qp(int qp_y) { return orc_chroma_qp_table[clip3(0, 57, qp_y)]; }

/* HM TComTrQuant::xQuant without RDOQ / sign hiding. */
int orc_quant(const int16_t *coeff, int16_t *level, int log2n, int qp, int intra_slice)
{
  int n = 1 << log2n, nz = 0;
  int transform_shift = 15 - 8 - log2n;
  int qbits = 14 + qp / 6 + transform_shift;
  int scale = orc_quant_scales[qp % 6];
  int64_t add = (int64_t)(intra_slice ? 171 : 85) << (qbits - 9);
  for (int i = 0; i < n * n; i++) {
    int c = coeff[i];
    int64_t a = ((int64_t)abs(c) * scale + add) >> qbits;
    int l = (int)(a > 32767 ? 32767 : a);
    if (c < 0) l = -l;
    level[i] = (int16_t)l;
    nz += l != 0;
  }
  return nz;
}

/* H.265 8.6.3 with flat scaling factor m = 16. */
void orc_dequant(const int16_t *level, int16_t *coeff, int log2n, int qp)
{
  int n = 1 << log2n;
  int bd_shift = 8 + log2n - 5;
  int scale = 16 * orc_level_scale[qp % 6];
  for (int i = 0; i < n * n; i++) {
    int64_t v = ((int64_t)level[i] * scale) << (qp / 6);
    v = (v + ((int64_t)1 << (bd_shift - 1))) >> bd_shift;
    coeff[i] = (int16_t)(v < -32768 ? -32768 : (v > 32767 ? 32767 : v));
  }
}

/* ---- K4: intra prediction -------------------------------------------------------------------- */

void orc_intra_predict(const uint8_t *refs_in, int log2n, int mode, int cidx, uint8_t *dst, int ds)
{
  orc_intra_predict2(refs_in, log2n, mode, cidx, 0, dst, ds);
}

void orc_intra_predict2(const uint8_t *refs_in, int log2n, int mode, int cidx, int strong, uint8_t *dst, int ds)
{
  const int n = 1 << log2n;
  uint8_t filt[4 * 32 + 1];
  const uint8_t *refs = refs_in;
  /* 8.4.4.2.3 filtering of neighbouring samples (luma only in 4:2:0) */
  if (cidx == 0 && mode != 1 && n != 4) {
    int d26 = abs(mode - 26), d10 = abs(mode - 10);
    int min_dist = d26 < d10 ? d26 : d10;
    int thres = n == 8 ? 7 : (n == 16 ? 1 : 0);
    if (min_dist > thres) {
      filt[0] = refs_in[0];
      filt[4 * n] = refs_in[4 * n];
      /* strong (bi-linear) smoothing of 32x32 blocks whose neighbours are nearly linear */
      const int corner = refs_in[2 * n];
      if (strong && n == 32 && abs(corner + refs_in[4 * n] - 2 * refs_in[3 * n]) < 8 &&
          abs(corner + refs_in[0] - 2 * refs_in[n]) < 8) {
        filt[2 * n] = (uint8_t)corner;
        for (int i = 0; i < 63; i++) {
          filt[2 * n - 1 - i] = (uint8_t)(((63 - i) * corner + (i + 1) * refs_in[0] + 32) >> 6);          /* left, downwards */
          filt[2 * n + 1 + i] = (uint8_t)(((63 - i) * corner + (i + 1) * refs_in[4 * n] + 32) >> 6);      /* top, rightwards */
        }
      } else {
        for (int i = 1; i < 4 * n; i++) filt[i] = (uint8_t)((refs_in[i - 1] + 2 * refs_in[i] + refs_in[i + 1] + 2) >> 2);
      }
      refs = filt;
    }
  }
  /* accessors in spec coordinates: left(y) = p[-1][y], top(x) = p[x][-1], y/x in -1..2N-1 */
#define LEFT(y) refs[2 * n - 1 - (y)]
#define TOP(x)  refs[2 * n + 1 + (x)]
  if (mode == 0) {                       /* 8.4.4.2.4 planar */
    for (int y = 0; y < n; y++)
      for (int x = 0; x < n; x++)
        dst[y * ds + x] = (uint8_t)(((n - 1 - x) * LEFT(y) + (x + 1) * TOP(n) + (n - 1 - y) * TOP(x) +
                                     (y + 1) * LEFT(n) + n) >> (log2n + 1));
    return;
  }
  if (mode == 1) {                       /* 8.4.4.2.5 DC */
    int sum = n;
    for (int i = 0; i < n; i++) sum += TOP(i) + LEFT(i);
    int dc = sum >> (log2n + 1);
    for (int y = 0; y < n; y++)
      for (int x = 0; x < n; x++) dst[y * ds + x] = (uint8_t)dc;
    if (cidx == 0 && n < 32) {
      dst[0] = (uint8_t)((LEFT(0) + 2 * dc + TOP(0) + 2) >> 2);
      for (int x = 1; x < n; x++) dst[x] = (uint8_t)((TOP(x) + 3 * dc + 2) >> 2);
      for (int y = 1; y < n; y++) dst[y * ds] = (uint8_t)((LEFT(y) + 3 * dc + 2) >> 2);
    }
    return;
  }
  /* 8.4.4.2.6 angular */
  int angle = orc_intra_pred_angle[mode];
  int ref_buf[3 * 32 + 2];
  int *ref = ref_buf + 32;                /* ref[-32 .. 2N] */
  if (mode >= 18) {
    for (int x = 0; x <= n; x++) ref[x] = TOP(x - 1);
    if (angle < 0) {
      int last = (n * angle) >> 5;
      if (last < -1)
        for (int x = last; x <= -1; x++) ref[x] = LEFT(-1 + ((x * orc_inv_angle[mode] + 128) >> 8));
    } else {
      for (int x = n + 1; x <= 2 * n; x++) ref[x] = TOP(x - 1);
    }
    for (int y = 0; y < n; y++) {
      int idx = ((y + 1) * angle) >> 5, fact = ((y + 1) * angle) & 31;
      for (int x = 0; x < n; x++)
        dst[y * ds + x] = (uint8_t)(fact ? ((32 - fact) * ref[x + idx + 1] + fact * ref[x + idx + 2] + 16) >> 5
                                         : ref[x + idx + 1]);
    }
    if (mode == 26 && cidx == 0 && n < 32)
      for (int y = 0; y < n; y++) dst[y * ds] = (uint8_t)clip8(TOP(0) + ((LEFT(y) - LEFT(-1)) >> 1));
  } else {
    for (int x = 0; x <= n; x++) ref[x] = LEFT(x - 1);
    if (angle < 0) {
      int last = (n * angle) >> 5;
      if (last < -1)
        for (int x = last; x <= -1; x++) ref[x] = TOP(-1 + ((x * orc_inv_angle[mode] + 128) >> 8));
    } else {
      for (int x = n + 1; x <= 2 * n; x++) ref[x] = LEFT(x - 1);
    }
    for (int x = 0; x < n; x++) {
      int idx = ((x + 1) * angle) >> 5, fact = ((x + 1) * angle) & 31;
      for (int y = 0; y < n; y++)
        dst[y * ds + x] = (uint8_t)(fact ? ((32 - fact) * ref[y + idx + 1] + fact * ref[y + idx + 2] + 16) >> 5
                                         : ref[y + idx + 1]);
    }
    if (mode == 10 && cidx == 0 && n < 32)
      for (int x = 0; x < n; x++) dst[x] = (uint8_t)clip8(LEFT(0) + ((TOP(x) - TOP(-1)) >> 1));
  }
#undef LEFT
#undef TOP
}

/* ---- K3: fractional sample interpolation (8.5.3.3.3), uni-prediction, 8 bit ----------------- */

static inline int ref_px(const uint8_t *ref, int stride, int pw, int ph, int x, int y)
{
  return ref[clip3(0, ph - 1, y) * stride + clip3(0, pw - 1, x)];
}

static void mc_generic(const uint8_t *ref, int stride, int pw, int ph, int xi, int yi, int w, int h,
                       const int8_t *fx, const int8_t *fy, int taps, int has_fx, int has_fy,
                       uint8_t *dst, int ds)
{
  const int before = taps / 2 - 1;
  if (!has_fx && !has_fy) {                 /* (x << 6 + 32) >> 6 == x */
    for (int y = 0; y < h; y++)
      for (int x = 0; x < w; x++) dst[y * ds + x] = (uint8_t)ref_px(ref, stride, pw, ph, xi + x, yi + y);
    return;
  }
  if (!has_fy) {
    for (int y = 0; y < h; y++)
      for (int x = 0; x < w; x++) {
        int v = 0;
        for (int t = 0; t < taps; t++) v += fx[t] * ref_px(ref, stride, pw, ph, xi + x + t - before, yi + y);
        dst[y * ds + x] = (uint8_t)clip8((v + 32) >> 6);
      }
    return;
  }
  if (!has_fx) {
    for (int y = 0; y < h; y++)
      for (int x = 0; x < w; x++) {
        int v = 0;
        for (int t = 0; t < taps; t++) v += fy[t] * ref_px(ref, stride, pw, ph, xi + x, yi + y + t - before);
        dst[y * ds + x] = (uint8_t)clip8((v + 32) >> 6);
      }
    return;
  }
  /* separable: rows first (shift1 = bitDepth - 8 = 0), then columns (shift2 = 6) */
  int16_t tmp[(64 + 7) * 64];
  for (int r = 0; r < h + taps - 1; r++)
    for (int x = 0; x < w; x++) {
      int a = 0;
      for (int t = 0; t < taps; t++) a += fx[t] * ref_px(ref, stride, pw, ph, xi + x + t - before, yi + r - before);
      tmp[r * w + x] = (int16_t)a;
    }
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      int v = 0;
      for (int t = 0; t < taps; t++) v += fy[t] * tmp[(y + t) * w + x];
      v >>= 6;
      dst[y * ds + x] = (uint8_t)clip8((v + 32) >> 6);   /* 8.5.3.3.4.2: shift1 = 14 - bitDepth */
    }
}

void orc_mc_luma(const uint8_t *ref, int stride, int pw, int ph, int x0, int y0, int w, int h,
                 int mvx, int mvy, uint8_t *dst, int ds)
{
  int fx = mvx & 3, fy = mvy & 3;
  mc_generic(ref, stride, pw, ph, x0 + (mvx >> 2), y0 + (mvy >> 2), w, h,
             orc_luma_filter[fx], orc_luma_filter[fy], 8, fx != 0, fy != 0, dst, ds);
}

void orc_mc_chroma(const uint8_t *ref, int stride, int pw, int ph, int x0, int y0, int w, int h,
                   int mvx, int mvy, uint8_t *dst, int ds)
{
  int fx = mvx & 7, fy = mvy & 7;
  mc_generic(ref, stride, pw, ph, x0 + (mvx >> 3), y0 + (mvy >> 3), w, h,
             orc_chroma_filter[fx], orc_chroma_filter[fy], 4, fx != 0, fy != 0, dst, ds);
}

/* ---- K7: deblocking filter (8.7.2.5), 8 bit, slice offsets 0 --------------------------------- */

void orc_deblock_luma_segment(uint8_t *pix, int xs, int ys, int bs, int qp)
{
  orc_deblock_luma_segment2(pix, xs, ys, bs, qp, 0, 0);
}

/* beta_off / tc_off: slice_beta_offset_div2 / slice_tc_offset_div2 (8.7.2.5.3) */
void orc_deblock_luma_segment2(uint8_t *pix, int xs, int ys, int bs, int qp, int beta_off, int tc_off)
{
  int beta = orc_beta_table[clip3(0, 51, qp + 2 * beta_off)];
  int tc = orc_tc_table[clip3(0, 53, qp + 2 * (bs - 1) + 2 * tc_off)];
#define P(i, l) pix[-(i + 1) * xs + (l) * ys]
#define Q(i, l) pix[(i) * xs + (l) * ys]
  int dp0 = abs(P(2, 0) - 2 * P(1, 0) + P(0, 0)), dp3 = abs(P(2, 3) - 2 * P(1, 3) + P(0, 3));
  int dq0 = abs(Q(2, 0) - 2 * Q(1, 0) + Q(0, 0)), dq3 = abs(Q(2, 3) - 2 * Q(1, 3) + Q(0, 3));
  int dpq0 = dp0 + dq0, dpq3 = dp3 + dq3, dp = dp0 + dp3, dq = dq0 + dq3;
  if (dpq0 + dpq3 >= beta) return;
  int strong0 = 2 * dpq0 < (beta >> 2) && abs(P(3, 0) - P(0, 0)) + abs(Q(0, 0) - Q(3, 0)) < (beta >> 3) &&
                abs(P(0, 0) - Q(0, 0)) < ((5 * tc + 1) >> 1);
  int strong3 = 2 * dpq3 < (beta >> 2) && abs(P(3, 3) - P(0, 3)) + abs(Q(0, 3) - Q(3, 3)) < (beta >> 3) &&
                abs(P(0, 3) - Q(0, 3)) < ((5 * tc + 1) >> 1);
  int dEp = dp < ((beta + (beta >> 1)) >> 3), dEq = dq < ((beta + (beta >> 1)) >> 3);
  for (int l = 0; l < 4; l++) {
    int p0 = P(0, l), p1 = P(1, l), p2 = P(2, l), p3 = P(3, l);
    int q0 = Q(0, l), q1 = Q(1, l), q2 = Q(2, l), q3 = Q(3, l);
    if (strong0 && strong3) {
      P(0, l) = (uint8_t)clip3(p0 - 2 * tc, p0 + 2 * tc, (p2 + 2 * p1 + 2 * p0 + 2 * q0 + q1 + 4) >> 3);
      P(1, l) = (uint8_t)clip3(p1 - 2 * tc, p1 + 2 * tc, (p2 + p1 + p0 + q0 + 2) >> 2);
      P(2, l) = (uint8_t)clip3(p2 - 2 * tc, p2 + 2 * tc, (2 * p3 + 3 * p2 + p1 + p0 + q0 + 4) >> 3);
      Q(0, l) = (uint8_t)clip3(q0 - 2 * tc, q0 + 2 * tc, (p1 + 2 * p0 + 2 * q0 + 2 * q1 + q2 + 4) >> 3);
      Q(1, l) = (uint8_t)clip3(q1 - 2 * tc, q1 + 2 * tc, (p0 + q0 + q1 + q2 + 2) >> 2);
      Q(2, l) = (uint8_t)clip3(q2 - 2 * tc, q2 + 2 * tc, (p0 + q0 + q1 + 3 * q2 + 2 * q3 + 4) >> 3);
    } else {
      int delta = (9 * (q0 - p0) - 3 * (q1 - p1) + 8) >> 4;
      if (abs(delta) < tc * 10) {
        delta = clip3(-tc, tc, delta);
        P(0, l) = (uint8_t)clip8(p0 + delta);
        Q(0, l) = (uint8_t)clip8(q0 - delta);
        if (dEp) P(1, l) = (uint8_t)clip8(p1 + clip3(-(tc >> 1), tc >> 1, (((p2 + p0 + 1) >> 1) - p1 + delta) >> 1));
        if (dEq) Q(1, l) = (uint8_t)clip8(q1 + clip3(-(tc >> 1), tc >> 1, (((q2 + q0 + 1) >> 1) - q1 - delta) >> 1));
      }
    }
  }
#undef P
#undef Q
}

/* chroma edges are filtered only for bS == 2; cQpPicOffset = 0 */
void orc_deblock_chroma_segment(uint8_t *pix, int xs, int ys, int qp_y, int lines)
{
  orc_deblock_chroma_segment2(pix, xs, ys, qp_y, lines, 0, 0);
}

/* c_off: cQpPicOffset = pps_cb_qp_offset / pps_cr_qp_offset; tc_off: slice_tc_offset_div2 (8.7.2.5.5) */
void orc_deblock_chroma_segment2(uint8_t *pix, int xs, int ys, int qp_y, int lines, int c_off, int tc_off)
{
  int qpc = orc_chroma_qp(qp_y + c_off);
  int tc = orc_tc_table[clip3(0, 53, qpc + 2 + 2 * tc_off)];
  for (int l = 0; l < lines; l++) {
    uint8_t *p = pix + l * ys;
    int p0 = p[-xs], p1 = p[-2 * xs], q0 = p[0], q1 = p[xs];
    int delta = clip3(-tc, tc, ((((q0 - p0) << 2) + p1 - q1 + 4) >> 3));
    p[-xs] = (uint8_t)clip8(p0 + delta);
    p[0] = (uint8_t)clip8(q0 - delta);
  }
}
