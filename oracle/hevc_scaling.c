/* TEST INFRASTRUCTURE ONLY -- see hevc_scaling.h */
#include "hevc_scaling.h"

#include <stdlib.h>
#include <string.h>

#include "hevc_tables.h"

/* Table 7-6, in up-right diagonal order; every colour component shares them, every size >= 8x8 as well */
static const uint8_t default_intra[64] = {
    16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 17, 16, 17, 16, 17, 18, 17, 18, 18, 17, 18, 21, 19, 20, 21, 20, 19, 21, 24, 22, 22, 24,
    24, 22, 22, 24, 25, 25, 27, 30, 27, 25, 25, 29, 31, 35, 35, 31, 29, 36, 41, 44, 41, 36, 47, 54, 54, 47, 65, 70, 65, 88, 88, 115};
static const uint8_t default_inter[64] = {
    16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 17, 17, 17, 17, 17, 18, 18, 18, 18, 18, 18, 20, 20, 20, 20, 20, 20, 20, 24, 24, 24, 24,
    24, 24, 24, 24, 25, 25, 25, 25, 25, 25, 25, 28, 28, 28, 28, 28, 28, 33, 33, 33, 33, 33, 41, 41, 41, 41, 54, 54, 54, 71, 71, 91};

static void set_default(orc_scaling_t *s, int sid, int mid)
{
  if (sid == 0) memset(s->list[0][mid], 16, 64);                                  /* Table 7-5 */
  else memcpy(s->list[sid][mid], mid < 3 ? default_intra : default_inter, 64);
  s->dc[sid][mid] = 16;
}

static void expand(orc_scaling_t *s)
{
  for (int sid = 0; sid < 4; sid++)
    for (int mid = 0; mid < 6; mid++) {
      const int l2 = sid == 0 ? 2 : 3, cnt = 1 << (2 * l2);
      memset(s->m[sid][mid], 16, 64);
      for (int i = 0; i < cnt; i++) {
        int x, y;
        orc_scan_pos(0, l2, i, &x, &y);
        s->m[sid][mid][(y << l2) + x] = s->list[sid][mid][i];
      }
    }
}

void orc_scaling_default(orc_scaling_t *s)
{
  memset(s, 0, sizeof(*s));
  for (int sid = 0; sid < 4; sid++)
    for (int mid = 0; mid < 6; mid++) { set_default(s, sid, mid); s->how[sid][mid] = ORC_SL_DEFAULT; }
  expand(s);
}

void orc_scaling_test_lists(orc_scaling_t *s)
{
  orc_scaling_default(s);
  for (int sid = 0; sid < 4; sid++) {
    const int cnt = sid == 0 ? 16 : 64, step = sid == 3 ? 3 : 1;
    for (int mid = 0; mid < 6; mid += step) {
      const int k = mid / step;                       /* 0..5, or 0..1 for 32x32 */
      if (k % 3 == 0) {                               /* coded: rising with the scan position, one dip; DC apart */
        for (int i = 0; i < cnt; i++) {
          int v = 10 + sid * 2 + (mid ? 3 : 0) + (i * (5 + sid + mid)) / (sid == 0 ? 3 : 8);
          if (i == cnt / 2) v -= 9;
          s->list[sid][mid][i] = (uint8_t)(v < 1 ? 1 : (v > 255 ? 255 : v));
        }
        s->dc[sid][mid] = (uint8_t)(sid > 1 ? 12 + sid + mid : 16);
        s->how[sid][mid] = ORC_SL_CODED;
      } else if (k % 3 == 1) {                        /* copied from the coded list before it (and its DC) */
        memcpy(s->list[sid][mid], s->list[sid][mid - step], 64);
        s->dc[sid][mid] = s->dc[sid][mid - step];
        s->how[sid][mid] = ORC_SL_COPY; s->ref[sid][mid] = 1;
      } else if (sid == 1) {                          /* copied from two lists back */
        memcpy(s->list[sid][mid], s->list[sid][mid - 2], 64);
        s->dc[sid][mid] = s->dc[sid][mid - 2];
        s->how[sid][mid] = ORC_SL_COPY; s->ref[sid][mid] = 2;
      }                                               /* else: default */
    }
  }
  /* 32x32 chroma lists do not exist in 4:2:0 (matrix 1, 2, 4, 5 of sizeId 3 are never read) */
  expand(s);
}

int orc_scaling_factor(const orc_scaling_t *s, int log2n, int matrix, int x, int y)
{
  const int sid = log2n - 2;
  if (sid == 0) return s->m[0][matrix][y * 4 + x];
  if (sid >= 2 && x == 0 && y == 0) return s->dc[sid][matrix];
  return s->m[sid][matrix][((y >> (sid - 1)) << 3) + (x >> (sid - 1))];
}

void orc_scaling_table(int mode, uint8_t *out)
{
  orc_scaling_t s;
  if (mode >= 2) orc_scaling_test_lists(&s); else orc_scaling_default(&s);
  memcpy(out, s.m, sizeof(s.m));
  memcpy(out + sizeof(s.m), s.dc[2], 6);
  memcpy(out + sizeof(s.m) + 6, s.dc[3], 6);
}

void orc_scaling_write(orc_bits_t *b, const orc_scaling_t *s)
{
  for (int sid = 0; sid < 4; sid++)
    for (int mid = 0; mid < 6; mid += sid == 3 ? 3 : 1) {
      if (s->how[sid][mid] != ORC_SL_CODED) {
        orc_bits_put(b, 0, 1);                                                    /* scaling_list_pred_mode_flag */
        orc_bits_ue(b, s->how[sid][mid] == ORC_SL_COPY ? s->ref[sid][mid] : 0);   /* scaling_list_pred_matrix_id_delta */
        continue;
      }
      orc_bits_put(b, 1, 1);
      int next = 8;
      const int cnt = sid == 0 ? 16 : 64;
      if (sid > 1) { orc_bits_se(b, (int)s->dc[sid][mid] - 8); next = s->dc[sid][mid]; }
      for (int i = 0; i < cnt; i++) {
        int d = (int)s->list[sid][mid][i] - next;
        if (d > 127) d -= 256;
        if (d < -128) d += 256;
        orc_bits_se(b, d);                                                        /* scaling_list_delta_coef */
        next = s->list[sid][mid][i];
      }
    }
}

int orc_quant_sl(const int16_t *coeff, int16_t *level, int log2n, int qp, int intra_slice, const orc_scaling_t *s, int matrix)
{
  const int n = 1 << log2n;
  const int qbits = 14 + qp / 6 + (15 - 8 - log2n);
  const int64_t add = (int64_t)(intra_slice ? 171 : 85) << (qbits - 9);
  int nz = 0;
  for (int i = 0; i < n * n; i++) {
    const int scale = (orc_quant_scales[qp % 6] << 4) / orc_scaling_factor(s, log2n, matrix, i & (n - 1), i >> log2n);
    const int c = coeff[i];
    const int64_t a = ((int64_t)abs(c) * scale + add) >> qbits;
    int l = (int)(a > 32767 ? 32767 : a);
    if (c < 0) l = -l;
    level[i] = (int16_t)l;
    nz += l != 0;
  }
  return nz;
}

void orc_dequant_sl(const int16_t *level, int16_t *coeff, int log2n, int qp, const orc_scaling_t *s, int matrix)
{
  const int n = 1 << log2n, bd_shift = 8 + log2n - 5;
  for (int i = 0; i < n * n; i++) {
    const int m = orc_scaling_factor(s, log2n, matrix, i & (n - 1), i >> log2n);
    int64_t v = ((int64_t)level[i] * m * orc_level_scale[qp % 6]) << (qp / 6);
    v = (v + ((int64_t)1 << (bd_shift - 1))) >> bd_shift;
    coeff[i] = (int16_t)(v < -32768 ? -32768 : (v > 32767 ? 32767 : v));
  }
}
