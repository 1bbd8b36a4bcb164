/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle for the colour-conversion stages.
 * Nothing under oracle/ is linked, imported or executed by the product path
 * (kvazzup_b200/); only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may use it, as the checker.
 *
 * Pinning status
 *   - I420 -> RGB32, half_rgb, flip_rgb: PINNED. Checked bit-exactly against
 *     the reference's own object code (oracle/_ref/libref_yuvconversions.so,
 *     built from /root/reference/src/media/processing/yuvconversions.cpp
 *     unmodified) in tests/test_oracle_conv.py, and against fixtures generated
 *     from it under tests/golden/.
 *   - camera format -> I420 (libyuv::ConvertToI420): PARITY UNPINNED. libyuv
 *     (pinned by the reference at git eb6e7bb63738e29efd82ea3cf2a115238a89fa51,
 *     dependencies/libyuv.cmake:12) is not vendored and not installed; this is
 *     a restatement of its published C row-function contract as invoked at
 *     /root/reference/src/media/processing/libyuvconverter.cpp:106-127.
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>

static inline int clamp255(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }

/* I420 -> RGB32, memory order B,G,R,0.
 * Follows the SIMD variants (canonical; the three SIMD variants are
 * bit-identical, the scalar _c fallback is not -- SURVEY.md 8a):
 * /root/reference/src/media/processing/yuvconversions.cpp:72-168, formulas :133-141.
 * All shifts are arithmetic shifts of signed 32-bit values. */
void oracle_i420_to_rgb32(const uint8_t *in, uint8_t *out, int w, int h)
{
  const uint8_t *py = in;
  const uint8_t *pu = in + (size_t)w * h;
  const uint8_t *pv = pu + (((size_t)w * h) >> 2);
  for (int y = 0; y < h; y++) {
    for (int x = 0; x < w; x++) {
      int Y = py[(size_t)y * w + x];
      int u = (int)pu[(size_t)(y >> 1) * (w >> 1) + (x >> 1)] - 128;
      int v = (int)pv[(size_t)(y >> 1) * (w >> 1) + (x >> 1)] - 128;
      int r = Y + (v + (v >> 2) + (v >> 3) + (v >> 5));
      int g = Y - (((u >> 2) + (u >> 4) + (u >> 5)) + ((v >> 1) + (v >> 3) + (v >> 4) + (v >> 5)));
      int b = Y + (u + (u >> 1) + (u >> 2) + (u >> 6));
      uint8_t *o = out + 4 * ((size_t)y * w + x);
      o[0] = (uint8_t)clamp255(b);
      o[1] = (uint8_t)clamp255(g);
      o[2] = (uint8_t)clamp255(r);
      o[3] = 0;
    }
  }
}

/* 2x point decimation, keeps the top-left pixel of every 2x2.
 * /root/reference/src/media/processing/yuvconversions.cpp:852-867 */
void oracle_half_rgb(const uint8_t *in, uint8_t *out, int w, int h)
{
  const uint32_t *src = (const uint32_t *)in;
  uint32_t *dst = (uint32_t *)out;
  for (int y = 0; y < h; y += 2)
    for (int x = 0; x < w; x += 2)
      dst[(size_t)(y / 2) * (w / 2) + x / 2] = src[(size_t)y * w + x];
}

/* Horizontal / vertical mirror; a no-op (output untouched) when both flags
 * are false. /root/reference/src/media/processing/yuvconversions.cpp:869-921 */
void oracle_flip_rgb(const uint8_t *in, uint8_t *out, int w, int h, int hor, int ver)
{
  if (!hor && !ver) return;
  const uint32_t *src = (const uint32_t *)in;
  uint32_t *dst = (uint32_t *)out;
  for (int y = 0; y < h; y++) {
    int sy = ver ? h - 1 - y : y;
    for (int x = 0; x < w; x++) {
      int sx = hor ? w - 1 - x : x;
      dst[(size_t)y * w + x] = src[(size_t)sy * w + sx];
    }
  }
}

/* ---------------------------------------------------------------------- */
/* libyuv::ConvertToI420 contract (restated, unpinned).                    */

#define ORC_FOURCC(a, b, c, d) \
  ((uint32_t)(a) | ((uint32_t)(b) << 8) | ((uint32_t)(c) << 16) | ((uint32_t)(d) << 24))

enum {
  ORC_FOURCC_I420 = ORC_FOURCC('I', '4', '2', '0'),
  ORC_FOURCC_I422 = ORC_FOURCC('I', '4', '2', '2'),
  ORC_FOURCC_NV12 = ORC_FOURCC('N', 'V', '1', '2'),
  ORC_FOURCC_NV21 = ORC_FOURCC('N', 'V', '2', '1'),
  ORC_FOURCC_YUY2 = ORC_FOURCC('Y', 'U', 'Y', '2'),
  ORC_FOURCC_YUYV = ORC_FOURCC('Y', 'U', 'Y', 'V'),
  ORC_FOURCC_UYVY = ORC_FOURCC('U', 'Y', 'V', 'Y'),
  ORC_FOURCC_ARGB = ORC_FOURCC('A', 'R', 'G', 'B'),
  ORC_FOURCC_BGRA = ORC_FOURCC('B', 'G', 'R', 'A'),
  ORC_FOURCC_ABGR = ORC_FOURCC('A', 'B', 'G', 'R'),
  ORC_FOURCC_RGBA = ORC_FOURCC('R', 'G', 'B', 'A'),
  ORC_FOURCC_24BG = ORC_FOURCC('2', '4', 'B', 'G'),
  ORC_FOURCC_RAW  = ORC_FOURCC('r', 'a', 'w', ' '),
  ORC_FOURCC_MJPG = ORC_FOURCC('M', 'J', 'P', 'G'),
};

/* BT.601 limited-range fixed point as documented for libyuv's C rows. */
static inline uint8_t rgb_to_y(int r, int g, int b) { return (uint8_t)((66 * r + 129 * g + 25 * b + 0x1080) >> 8); }
static inline uint8_t rgb_to_u(int r, int g, int b) { return (uint8_t)((112 * b - 74 * g - 38 * r + 0x8080) >> 8); }
static inline uint8_t rgb_to_v(int r, int g, int b) { return (uint8_t)((112 * r - 94 * g - 18 * b + 0x8080) >> 8); }

/* Generic packed-RGB -> I420: bpp bytes per pixel, channel byte offsets
 * (ro,go,bo). Chroma from the 2x2 box average with round-half-up. */
static void rgb_to_i420(const uint8_t *s, int bpp, int ro, int go, int bo,
                        uint8_t *y, int sy, uint8_t *u, int su, uint8_t *v, int sv,
                        int w, int h)
{
  size_t stride = (size_t)w * bpp;
  for (int j = 0; j < h; j++)
    for (int i = 0; i < w; i++) {
      const uint8_t *p = s + j * stride + (size_t)i * bpp;
      y[(size_t)j * sy + i] = rgb_to_y(p[ro], p[go], p[bo]);
    }
  for (int j = 0; j < h; j += 2) {
    int j1 = (j + 1 < h) ? j + 1 : j;
    for (int i = 0; i < w; i += 2) {
      int i1 = (i + 1 < w) ? i + 1 : i;
      const uint8_t *p00 = s + j * stride + (size_t)i * bpp;
      const uint8_t *p01 = s + j * stride + (size_t)i1 * bpp;
      const uint8_t *p10 = s + j1 * stride + (size_t)i * bpp;
      const uint8_t *p11 = s + j1 * stride + (size_t)i1 * bpp;
      int r = (p00[ro] + p01[ro] + p10[ro] + p11[ro] + 2) >> 2;
      int g = (p00[go] + p01[go] + p10[go] + p11[go] + 2) >> 2;
      int b = (p00[bo] + p01[bo] + p10[bo] + p11[bo] + 2) >> 2;
      u[(size_t)(j / 2) * su + i / 2] = rgb_to_u(r, g, b);
      v[(size_t)(j / 2) * sv + i / 2] = rgb_to_v(r, g, b);
    }
  }
}

/* Packed 4:2:2 (YUY2: Y0 U Y1 V ; UYVY: U Y0 V Y1) -> I420; chroma is the
 * round-half-up average of the two source rows. */
static void packed422_to_i420(const uint8_t *s, int yoff, int uoff, int voff,
                              uint8_t *y, int sy, uint8_t *u, int su, uint8_t *v, int sv,
                              int w, int h)
{
  size_t stride = (size_t)((w + 1) / 2) * 4;
  for (int j = 0; j < h; j++)
    for (int i = 0; i < w; i++)
      y[(size_t)j * sy + i] = s[j * stride + (size_t)(i >> 1) * 4 + yoff + 2 * (i & 1)];
  for (int j = 0; j < h; j += 2) {
    int j1 = (j + 1 < h) ? j + 1 : j;
    for (int i = 0; i < (w + 1) / 2; i++) {
      const uint8_t *a = s + j * stride + (size_t)i * 4;
      const uint8_t *b = s + j1 * stride + (size_t)i * 4;
      u[(size_t)(j / 2) * su + i] = (uint8_t)((a[uoff] + b[uoff] + 1) >> 1);
      v[(size_t)(j / 2) * sv + i] = (uint8_t)((a[voff] + b[voff] + 1) >> 1);
    }
  }
}

/* Same argument list as libyuv::ConvertToI420 with crop = whole frame and
 * rotation 0 (the only way the reference calls it,
 * /root/reference/src/media/processing/libyuvconverter.cpp:120-127).
 * Returns 0 on success, -1 for an unsupported fourcc (output untouched). */
int oracle_mjpg_to_i420(const uint8_t *sample, size_t sample_size, uint8_t *y, int sy, uint8_t *u, int su, uint8_t *v, int sv, int w, int h);

int oracle_convert_to_i420(const uint8_t *sample, size_t sample_size,
                           uint8_t *y, int sy, uint8_t *u, int su, uint8_t *v, int sv,
                           int w, int h, uint32_t fourcc)
{
  int hw = (w + 1) / 2, hh = (h + 1) / 2;
  if (!sample || !y || !u || !v || w <= 0 || h <= 0) return -1;
  switch (fourcc) {
  case ORC_FOURCC_I420: {
    for (int j = 0; j < h; j++) memcpy(y + (size_t)j * sy, sample + (size_t)j * w, w);
    const uint8_t *pu = sample + (size_t)w * h, *pv = pu + (size_t)hw * hh;
    for (int j = 0; j < hh; j++) {
      memcpy(u + (size_t)j * su, pu + (size_t)j * hw, hw);
      memcpy(v + (size_t)j * sv, pv + (size_t)j * hw, hw);
    }
    return 0;
  }
  case ORC_FOURCC_I422: {
    for (int j = 0; j < h; j++) memcpy(y + (size_t)j * sy, sample + (size_t)j * w, w);
    const uint8_t *pu = sample + (size_t)w * h, *pv = pu + (size_t)hw * h;
    for (int j = 0; j < hh; j++) {
      int j0 = 2 * j, j1 = (2 * j + 1 < h) ? 2 * j + 1 : 2 * j;
      for (int i = 0; i < hw; i++) {
        u[(size_t)j * su + i] = (uint8_t)((pu[(size_t)j0 * hw + i] + pu[(size_t)j1 * hw + i] + 1) >> 1);
        v[(size_t)j * sv + i] = (uint8_t)((pv[(size_t)j0 * hw + i] + pv[(size_t)j1 * hw + i] + 1) >> 1);
      }
    }
    return 0;
  }
  case ORC_FOURCC_NV12:
  case ORC_FOURCC_NV21: {
    for (int j = 0; j < h; j++) memcpy(y + (size_t)j * sy, sample + (size_t)j * w, w);
    const uint8_t *puv = sample + (size_t)w * h;
    int first_is_u = (fourcc == ORC_FOURCC_NV12);
    for (int j = 0; j < hh; j++)
      for (int i = 0; i < hw; i++) {
        uint8_t a = puv[(size_t)j * hw * 2 + 2 * i], b = puv[(size_t)j * hw * 2 + 2 * i + 1];
        u[(size_t)j * su + i] = first_is_u ? a : b;
        v[(size_t)j * sv + i] = first_is_u ? b : a;
      }
    return 0;
  }
  case ORC_FOURCC_YUY2:
  case ORC_FOURCC_YUYV:
    packed422_to_i420(sample, 0, 1, 3, y, sy, u, su, v, sv, w, h);
    return 0;
  case ORC_FOURCC_UYVY:
    packed422_to_i420(sample, 1, 0, 2, y, sy, u, su, v, sv, w, h);
    return 0;
  /* libyuv FOURCC names give the little-endian WORD order; memory order is reversed. */
  case ORC_FOURCC_ARGB: rgb_to_i420(sample, 4, 2, 1, 0, y, sy, u, su, v, sv, w, h); return 0; /* mem B,G,R,A */
  case ORC_FOURCC_BGRA: rgb_to_i420(sample, 4, 1, 2, 3, y, sy, u, su, v, sv, w, h); return 0; /* mem A,R,G,B */
  case ORC_FOURCC_ABGR: rgb_to_i420(sample, 4, 0, 1, 2, y, sy, u, su, v, sv, w, h); return 0; /* mem R,G,B,A */
  case ORC_FOURCC_RGBA: rgb_to_i420(sample, 4, 3, 2, 1, y, sy, u, su, v, sv, w, h); return 0; /* mem A,B,G,R */
  case ORC_FOURCC_24BG: rgb_to_i420(sample, 3, 2, 1, 0, y, sy, u, su, v, sv, w, h); return 0; /* mem B,G,R */
  case ORC_FOURCC_RAW:  rgb_to_i420(sample, 3, 0, 1, 2, y, sy, u, su, v, sv, w, h); return 0; /* mem R,G,B */
  case ORC_FOURCC_MJPG: return oracle_mjpg_to_i420(sample, sample_size, y, sy, u, su, v, sv, w, h);   /* mjpg_oracle.c */
  default:
    /* includes the literal `2` the reference passes for DT_RGB24VIDEO (libyuvconverter.cpp:84) */
    return -1;
  }
}
