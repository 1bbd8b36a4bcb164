/* TEST INFRASTRUCTURE ONLY -- CPU restatement of the MJPG leg of libyuv::ConvertToI420
 * (reference call site /root/reference/src/media/processing/libyuvconverter.cpp:94,120-127).
 *
 * libyuv (eb6e7bb6, dependencies/libyuv.cmake:12 -- not vendored) decodes the frame with libjpeg in
 * raw-data mode (no upsampling, no colour conversion: the planes are the IDCT output), default DCT
 * method JDCT_ISLOW, then converts the subsampling to 4:2:0:
 *   4:2:0  planes copied                       4:2:2  chroma rows averaged in pairs, (a + b + 1) >> 1
 *   4:4:4  chroma 2x2 box, (a+b+c+d+2) >> 2    4:0:0  chroma = 128
 * Restated here: ITU-T T.81 baseline sequential Huffman decoding (F.2.2), dequantisation, the
 * "accurate integer" inverse DCT of libjpeg's jidctint.c (the Loeffler-Ligtenberg-Moschytz
 * factorisation with 13-bit constants and 2 extra bits after the column pass, restated from its
 * published description), level shift + clamp.  Pinned by FFmpeg's mjpeg decoder with idct=int
 * (tests/test_oracle_mjpg.py), which must reproduce every plane bit for bit.
 *
 * Scope: 8-bit baseline (SOF0), one interleaved scan, 1 or 3 components, sampling 1x1 / 2x1 / 2x2
 * luma with 1x1 chroma, restart intervals, missing DHT = the tables of T.81 Annex K (what UVC
 * cameras and AVI MJPG rely on). */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---- Annex K.3 tables for streams without DHT ------------------------------------------------ */
static const uint8_t k_dc_lum_bits[16] = {0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0};
static const uint8_t k_dc_chr_bits[16] = {0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0};
static const uint8_t k_dc_vals[12] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11};
static const uint8_t k_ac_lum_bits[16] = {0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7d};
static const uint8_t k_ac_lum_vals[162] = {
  0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07, 0x22, 0x71, 0x14, 0x32, 0x81, 0x91,
  0xa1, 0x08, 0x23, 0x42, 0xb1, 0xc1, 0x15, 0x52, 0xd1, 0xf0, 0x24, 0x33, 0x62, 0x72, 0x82, 0x09, 0x0a, 0x16, 0x17, 0x18, 0x19, 0x1a,
  0x25, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x34, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a, 0x53,
  0x54, 0x55, 0x56, 0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79,
  0x7a, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5,
  0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9,
  0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe1, 0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf1, 0xf2,
  0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa};
static const uint8_t k_ac_chr_bits[16] = {0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77};
static const uint8_t k_ac_chr_vals[162] = {
  0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71, 0x13, 0x22, 0x32, 0x81, 0x08, 0x14,
  0x42, 0x91, 0xa1, 0xb1, 0xc1, 0x09, 0x23, 0x33, 0x52, 0xf0, 0x15, 0x62, 0x72, 0xd1, 0x0a, 0x16, 0x24, 0x34, 0xe1, 0x25, 0xf1, 0x17,
  0x18, 0x19, 0x1a, 0x26, 0x27, 0x28, 0x29, 0x2a, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49, 0x4a,
  0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78,
  0x79, 0x7a, 0x82, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3,
  0xa4, 0xa5, 0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5, 0xc6, 0xc7,
  0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf2,
  0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8, 0xf9, 0xfa};

static const uint8_t k_zigzag[64] = {                 /* zig-zag index -> row-major position (T.81 figure A.6) */
  0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
  35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

/* ---- Huffman table as T.81 C.2 / F.2.2.3 builds it: codes of each length in order ------------ */
typedef struct {
  int present;
  uint8_t bits[16], vals[256];
  int mincode[17], maxcode[18], valptr[17];
} huff_t;

static void huff_build(huff_t *h, const uint8_t *bits, const uint8_t *vals, int nvals)
{
  memcpy(h->bits, bits, 16);
  memcpy(h->vals, vals, (size_t)nvals);
  int code = 0, k = 0;
  for (int len = 1; len <= 16; len++) {
    h->valptr[len] = k;
    h->mincode[len] = code;
    code += bits[len - 1];
    k += bits[len - 1];
    h->maxcode[len] = bits[len - 1] ? code - 1 : -1;
    code <<= 1;
  }
  h->maxcode[17] = 0x7fffffff;
  h->present = 1;
}

/* ---- entropy-coded segment reader: byte stuffing, markers ------------------------------------ */
typedef struct { const uint8_t *p, *end; uint32_t acc; int nbits; int marker; } bitrd_t;

static int rd_bit(bitrd_t *b)
{
  if (b->nbits == 0) {
    int c = 0;
    if (!b->marker && b->p < b->end) {
      c = *b->p++;
      if (c == 0xff) {
        int c2 = b->p < b->end ? *b->p : 0;
        if (c2 == 0) b->p++;                        /* stuffed zero */
        else { b->marker = c2; b->p--; c = 0; }     /* a marker: the segment ends, feed zeros */
      }
    }
    b->acc = (uint32_t)c; b->nbits = 8;
  }
  b->nbits--;
  return (int)((b->acc >> b->nbits) & 1);
}
static int rd_bits(bitrd_t *b, int n) { int v = 0; while (n--) v = (v << 1) | rd_bit(b); return v; }

static int huff_decode(bitrd_t *b, const huff_t *h)
{
  int code = 0;
  for (int len = 1; len <= 16; len++) {
    code = (code << 1) | rd_bit(b);
    if (h->maxcode[len] >= 0 && code <= h->maxcode[len] && code >= h->mincode[len])
      return h->vals[h->valptr[len] + code - h->mincode[len]];
  }
  return -1;
}
static int extend(int v, int t) { return t && v < (1 << (t - 1)) ? v - (1 << t) + 1 : v; }   /* F.2.2.1 */

/* ---- inverse DCT: jidctint.c ("islow") -- 13-bit constants, 2 extra bits after the column pass -- */
#define CB 13
#define P1 2
#define DESCALE(x, n) (((x) + (1 << ((n) - 1))) >> (n))
static void idct_islow(const int16_t *coef, const uint16_t *q, uint8_t *out, int stride)
{
  int ws[64];
  for (int c = 0; c < 8; c++) {                       /* columns */
    int in[8];
    for (int r = 0; r < 8; r++) in[r] = coef[r * 8 + c] * q[r * 8 + c];
    int z2 = in[2], z3 = in[6];
    int z1 = (z2 + z3) * 4433;
    int t2 = z1 + z3 * -15137, t3 = z1 + z2 * 6270;
    int t0 = (in[0] + in[4]) << CB, t1 = (in[0] - in[4]) << CB;
    int t10 = t0 + t3, t13 = t0 - t3, t11 = t1 + t2, t12 = t1 - t2;
    int o0 = in[7], o1 = in[5], o2 = in[3], o3 = in[1];
    int y1 = o0 + o3, y2 = o1 + o2, y3 = o0 + o2, y4 = o1 + o3, y5 = (y3 + y4) * 9633;
    o0 *= 2446; o1 *= 16819; o2 *= 25172; o3 *= 12299;
    y1 *= -7373; y2 *= -20995; y3 *= -16069; y4 *= -3196;
    y3 += y5; y4 += y5;
    o0 += y1 + y3; o1 += y2 + y4; o2 += y2 + y3; o3 += y1 + y4;
    ws[0 * 8 + c] = DESCALE(t10 + o3, CB - P1); ws[7 * 8 + c] = DESCALE(t10 - o3, CB - P1);
    ws[1 * 8 + c] = DESCALE(t11 + o2, CB - P1); ws[6 * 8 + c] = DESCALE(t11 - o2, CB - P1);
    ws[2 * 8 + c] = DESCALE(t12 + o1, CB - P1); ws[5 * 8 + c] = DESCALE(t12 - o1, CB - P1);
    ws[3 * 8 + c] = DESCALE(t13 + o0, CB - P1); ws[4 * 8 + c] = DESCALE(t13 - o0, CB - P1);
  }
  for (int r = 0; r < 8; r++) {                       /* rows */
    const int *in = ws + r * 8;
    int z2 = in[2], z3 = in[6];
    int z1 = (z2 + z3) * 4433;
    int t2 = z1 + z3 * -15137, t3 = z1 + z2 * 6270;
    int t0 = (in[0] + in[4]) << CB, t1 = (in[0] - in[4]) << CB;
    int t10 = t0 + t3, t13 = t0 - t3, t11 = t1 + t2, t12 = t1 - t2;
    int o0 = in[7], o1 = in[5], o2 = in[3], o3 = in[1];
    int y1 = o0 + o3, y2 = o1 + o2, y3 = o0 + o2, y4 = o1 + o3, y5 = (y3 + y4) * 9633;
    o0 *= 2446; o1 *= 16819; o2 *= 25172; o3 *= 12299;
    y1 *= -7373; y2 *= -20995; y3 *= -16069; y4 *= -3196;
    y3 += y5; y4 += y5;
    o0 += y1 + y3; o1 += y2 + y4; o2 += y2 + y3; o3 += y1 + y4;
    const int v[8] = {t10 + o3, t11 + o2, t12 + o1, t13 + o0, t13 - o0, t12 - o1, t11 - o2, t10 - o3};
    for (int c = 0; c < 8; c++) {
      int s = DESCALE(v[c], CB + P1 + 3) + 128;       /* level shift, then the range limit */
      out[r * stride + c] = (uint8_t)(s < 0 ? 0 : (s > 255 ? 255 : s));
    }
  }
}

typedef struct { int id, h, v, tq, td, ta, bw, bh; uint8_t *plane; int pred; } comp_t;

/* Decodes a baseline JPEG into its component planes (padded to whole MCUs).  Returns 0 on success.
 * comps[i].plane is malloc'ed: (bw * 8) x (bh * 8) samples. */
static int jpeg_decode(const uint8_t *d, size_t n, int *pw, int *ph, int *ncomp, comp_t comps[3])
{
  uint16_t qt[4][64];
  huff_t dc[4], ac[4];
  memset(dc, 0, sizeof(dc)); memset(ac, 0, sizeof(ac)); memset(qt, 0, sizeof(qt));
  int have_q[4] = {0, 0, 0, 0}, restart = 0, w = 0, h = 0, nc = 0, hmax = 1, vmax = 1;
  size_t p = 0;
  if (n < 4 || d[0] != 0xff || d[1] != 0xd8) return -1;
  p = 2;
  for (;;) {
    while (p < n && d[p] != 0xff) p++;
    while (p < n && d[p] == 0xff) p++;
    if (p >= n) return -1;
    const int m = d[p++];
    if (m == 0xd9) return -1;                         /* EOI before any scan */
    if (m >= 0xd0 && m <= 0xd7) continue;
    if (p + 2 > n) return -1;
    const size_t len = ((size_t)d[p] << 8) | d[p + 1];
    if (len < 2 || p + len > n) return -1;
    const uint8_t *s = d + p + 2;
    const size_t sl = len - 2;
    if (m == 0xdb) {                                  /* DQT */
      for (size_t o = 0; o < sl;) {
        const int pq = s[o] >> 4, tq = s[o] & 15;
        o++;
        if (tq > 3 || o + (pq ? 128 : 64) > sl) return -1;
        for (int i = 0; i < 64; i++) { qt[tq][k_zigzag[i]] = pq ? (uint16_t)((s[o] << 8) | s[o + 1]) : s[o]; o += pq ? 2 : 1; }
        have_q[tq] = 1;
      }
    } else if (m == 0xc4) {                           /* DHT */
      for (size_t o = 0; o + 17 <= sl;) {
        const int tc = s[o] >> 4, th = s[o] & 15;
        int cnt = 0;
        for (int i = 0; i < 16; i++) cnt += s[o + 1 + i];
        if (tc > 1 || th > 3 || cnt > 256 || o + 17 + (size_t)cnt > sl) return -1;
        huff_build(tc ? &ac[th] : &dc[th], s + o + 1, s + o + 17, cnt);
        o += 17 + (size_t)cnt;
      }
    } else if (m == 0xdd) {                           /* DRI */
      if (sl < 2) return -1;
      restart = (s[0] << 8) | s[1];
    } else if (m == 0xc0 || m == 0xc1) {              /* SOF0 (SOF1 with 8-bit precision decodes the same) */
      if (sl < 6 || s[0] != 8) return -1;
      h = (s[1] << 8) | s[2]; w = (s[3] << 8) | s[4]; nc = s[5];
      if ((nc != 1 && nc != 3) || sl < 6 + 3 * (size_t)nc || w <= 0 || h <= 0) return -1;
      for (int i = 0; i < nc; i++) {
        comps[i].id = s[6 + 3 * i]; comps[i].h = s[7 + 3 * i] >> 4; comps[i].v = s[7 + 3 * i] & 15; comps[i].tq = s[8 + 3 * i];
        if (comps[i].h < 1 || comps[i].h > 2 || comps[i].v < 1 || comps[i].v > 2 || comps[i].tq > 3) return -1;
        if (comps[i].h > hmax) hmax = comps[i].h;
        if (comps[i].v > vmax) vmax = comps[i].v;
      }
    } else if (m == 0xc2 || (m >= 0xc3 && m <= 0xcf && m != 0xc4 && m != 0xc8 && m != 0xcc)) {
      return -1;                                      /* progressive / lossless / arithmetic: not baseline */
    } else if (m == 0xda) {                           /* SOS */
      if (!nc || sl < 1 || s[0] != nc || sl < 1 + 2 * (size_t)nc + 3) return -1;
      for (int i = 0; i < nc; i++) {
        int ci = -1;
        for (int k = 0; k < nc; k++) if (comps[k].id == s[1 + 2 * i]) ci = k;
        if (ci != i) return -1;                       /* components in frame order, one interleaved scan */
        comps[i].td = s[2 + 2 * i] >> 4; comps[i].ta = s[2 + 2 * i] & 15;
        if (comps[i].td > 3 || comps[i].ta > 3) return -1;
      }
      p += len;
      break;
    }
    p += len;
  }
  if (!dc[0].present && !ac[0].present) {             /* no DHT at all: Annex K */
    huff_build(&dc[0], k_dc_lum_bits, k_dc_vals, 12); huff_build(&dc[1], k_dc_chr_bits, k_dc_vals, 12);
    huff_build(&ac[0], k_ac_lum_bits, k_ac_lum_vals, 162); huff_build(&ac[1], k_ac_chr_bits, k_ac_chr_vals, 162);
  }
  const int mcux = (w + 8 * hmax - 1) / (8 * hmax), mcuy = (h + 8 * vmax - 1) / (8 * vmax);
  for (int i = 0; i < nc; i++) {
    if (!have_q[comps[i].tq] || !dc[comps[i].td].present || !ac[comps[i].ta].present) return -1;
    if (nc == 1) { comps[i].h = comps[i].v = 1; }     /* a single component is never interleaved: 8x8 MCUs */
    comps[i].bw = (nc == 1 ? (w + 7) / 8 : mcux * comps[i].h);
    comps[i].bh = (nc == 1 ? (h + 7) / 8 : mcuy * comps[i].v);
    comps[i].plane = (uint8_t *)malloc((size_t)comps[i].bw * comps[i].bh * 64);
    comps[i].pred = 0;
  }
  const int nmx = nc == 1 ? comps[0].bw : mcux, nmy = nc == 1 ? comps[0].bh : mcuy;
  bitrd_t b = {d + p, d + n, 0, 0, 0};
  int count = 0;
  for (int my = 0; my < nmy; my++)
    for (int mx = 0; mx < nmx; mx++) {
      if (restart && count == restart) {              /* RSTn: byte-align, reset the predictors */
        b.nbits = 0;
        if (b.marker >= 0xd0 && b.marker <= 0xd7) { b.p += 2; b.marker = 0; }
        else if (b.p + 1 < b.end && b.p[0] == 0xff && b.p[1] >= 0xd0 && b.p[1] <= 0xd7) b.p += 2;
        for (int i = 0; i < nc; i++) comps[i].pred = 0;
        count = 0;
      }
      count++;
      for (int i = 0; i < nc; i++)
        for (int by = 0; by < comps[i].v; by++)
          for (int bx = 0; bx < comps[i].h; bx++) {
            int16_t blk[64];
            memset(blk, 0, sizeof(blk));
            int t = huff_decode(&b, &dc[comps[i].td]);
            if (t < 0 || t > 11) return -1;
            comps[i].pred += extend(rd_bits(&b, t), t);
            blk[0] = (int16_t)comps[i].pred;
            for (int k = 1; k < 64;) {
              const int rs = huff_decode(&b, &ac[comps[i].ta]);
              if (rs < 0) return -1;
              const int r = rs >> 4, sz = rs & 15;
              if (sz == 0) { if (r == 15) { k += 16; continue; } break; }
              k += r;
              if (k > 63) return -1;
              blk[k_zigzag[k]] = (int16_t)extend(rd_bits(&b, sz), sz);
              k++;
            }
            const int X = (mx * comps[i].h + bx) * 8, Y = (my * comps[i].v + by) * 8, st = comps[i].bw * 8;
            idct_islow(blk, qt[comps[i].tq], comps[i].plane + (size_t)Y * st + X, st);
          }
    }
  *pw = w; *ph = h; *ncomp = nc;
  return 0;
}

/* MJPG leg of ConvertToI420.  Returns 0 on success, -1 on a frame that cannot be decoded or whose size is
 * not w x h (libyuv::MJPGToI420 fails likewise); the output is then untouched. */
int oracle_mjpg_to_i420(const uint8_t *sample, size_t sample_size, uint8_t *y, int sy, uint8_t *u, int su, uint8_t *v, int sv, int w, int h)
{
  comp_t c[3];
  memset(c, 0, sizeof(c));
  int jw = 0, jh = 0, nc = 0, rc = -1;
  if (jpeg_decode(sample, sample_size, &jw, &jh, &nc, c) != 0 || jw != w || jh != h) goto done;
  const int hw = (w + 1) / 2, hh = (h + 1) / 2;
  int mode;                                            /* 0: 4:2:0, 1: 4:2:2, 2: 4:4:4, 3: 4:0:0 */
  if (nc == 1) mode = 3;
  else if (c[1].h != 1 || c[1].v != 1 || c[2].h != 1 || c[2].v != 1) goto done;
  else if (c[0].h == 2 && c[0].v == 2) mode = 0;
  else if (c[0].h == 2 && c[0].v == 1) mode = 1;
  else if (c[0].h == 1 && c[0].v == 1) mode = 2;
  else goto done;
  for (int j = 0; j < h; j++) memcpy(y + (size_t)j * sy, c[0].plane + (size_t)j * c[0].bw * 8, (size_t)w);
  for (int k = 1; k < 3; k++) {
    uint8_t *dst = k == 1 ? u : v;
    const int ds = k == 1 ? su : sv;
    if (mode == 3) { for (int j = 0; j < hh; j++) memset(dst + (size_t)j * ds, 128, (size_t)hw); continue; }
    const uint8_t *p = c[k].plane;
    const int st = c[k].bw * 8;
    for (int j = 0; j < hh; j++)
      for (int i = 0; i < hw; i++) {
        int val;
        if (mode == 0) val = p[(size_t)j * st + i];
        else if (mode == 1) {
          const int j1 = 2 * j + 1 < h ? 2 * j + 1 : 2 * j;
          val = (p[(size_t)(2 * j) * st + i] + p[(size_t)j1 * st + i] + 1) >> 1;
        } else {
          const int j1 = 2 * j + 1 < h ? 2 * j + 1 : 2 * j, i1 = 2 * i + 1 < w ? 2 * i + 1 : 2 * i;
          val = (p[(size_t)(2 * j) * st + 2 * i] + p[(size_t)(2 * j) * st + i1] + p[(size_t)j1 * st + 2 * i] + p[(size_t)j1 * st + i1] + 2) >> 2;
        }
        dst[(size_t)j * ds + i] = (uint8_t)val;
      }
  }
  rc = 0;
done:
  for (int i = 0; i < 3; i++) free(c[i].plane);
  return rc;
}

/* The raw component planes (tests: the pin against FFmpeg's decoder).  plane[i] must hold pw[i] * ph[i]
 * bytes with pw / ph the padded sizes returned by a first call with plane == NULL.  Returns the number
 * of components, -1 on error. */
int oracle_mjpg_planes(const uint8_t *sample, size_t sample_size, uint8_t **plane, int *pw, int *ph, int *w, int *h)
{
  comp_t c[3];
  memset(c, 0, sizeof(c));
  int nc = 0;
  if (jpeg_decode(sample, sample_size, w, h, &nc, c) != 0) { for (int i = 0; i < 3; i++) free(c[i].plane); return -1; }
  for (int i = 0; i < nc; i++) {
    pw[i] = c[i].bw * 8; ph[i] = c[i].bh * 8;
    if (plane && plane[i]) memcpy(plane[i], c[i].plane, (size_t)pw[i] * ph[i]);
    free(c[i].plane);
  }
  return nc;
}
