/* TEST INFRASTRUCTURE ONLY -- see hevc_cabac.h. */
#include "hevc_cabac.h"

#include <stdlib.h>
#include <string.h>

/* ---- bit writer ------------------------------------------------------------------------- */

void orc_bits_init(orc_bits_t *b, uint8_t *buf, size_t cap)
{
  b->buf = buf; b->cap = cap; b->pos = 0; b->cur = 0; b->nbits = 0; b->overflow = 0;
}

static void put_byte(orc_bits_t *b, uint8_t v)
{
  if (b->pos < b->cap) b->buf[b->pos++] = v;
  else b->overflow = 1;
}

void orc_bits_put(orc_bits_t *b, uint32_t val, int n)
{
  for (int i = n - 1; i >= 0; i--) {
    b->cur = (b->cur << 1) | ((val >> i) & 1);
    if (++b->nbits == 8) { put_byte(b, (uint8_t)b->cur); b->cur = 0; b->nbits = 0; }
  }
}

void orc_bits_ue(orc_bits_t *b, uint32_t v)
{
  uint32_t x = v + 1;
  int len = 0;
  while ((x >> len) > 1) len++;
  orc_bits_put(b, 0, len);
  orc_bits_put(b, x, len + 1);
}

void orc_bits_se(orc_bits_t *b, int32_t v) { orc_bits_ue(b, v > 0 ? (uint32_t)(2 * v - 1) : (uint32_t)(-2 * v)); }

void orc_bits_trailing(orc_bits_t *b)
{
  orc_bits_put(b, 1, 1);
  while (b->nbits) orc_bits_put(b, 0, 1);
}

size_t orc_bits_bytes(const orc_bits_t *b) { return b->pos; }

size_t orc_nal_escape(const uint8_t *rbsp, size_t n, uint8_t *out, size_t cap)
{
  size_t o = 0;
  int zeros = 0;
  for (size_t i = 0; i < n; i++) {
    if (zeros >= 2 && rbsp[i] <= 3) {
      if (o < cap) out[o] = 3;
      o++;
      zeros = 0;
    }
    if (o < cap) out[o] = rbsp[i];
    o++;
    zeros = rbsp[i] == 0 ? zeros + 1 : 0;
  }
  return o;
}

/* ---- CABAC encoder (9.3.4.x; register layout of the HM TEncBinCABAC) --------------------- */

void orc_cabac_init_contexts(orc_cabac_t *c, int init_type, int slice_qp)
{
  int qp = slice_qp < 0 ? 0 : (slice_qp > 51 ? 51 : slice_qp);
  for (int i = 0; i < CTX_COUNT; i++) {          /* 9.3.2.2 */
    int iv = orc_ctx_init[init_type][i];
    int m = (iv >> 4) * 5 - 45;
    int n = ((iv & 15) << 3) - 16;
    int pre = ((m * qp) >> 4) + n;
    pre = pre < 1 ? 1 : (pre > 126 ? 126 : pre);
    int mps = pre <= 63 ? 0 : 1;
    int st = mps ? pre - 64 : 63 - pre;
    c->ctx[i] = (uint8_t)((st << 1) | mps);
  }
}

void orc_cabac_start(orc_cabac_t *c, orc_bits_t *out)
{
  c->low = 0; c->range = 510; c->bits_left = 23; c->num_buffered = 0; c->buffered_byte = 0xff;
  c->out = out;
}

static void cabac_write_out(orc_cabac_t *c)
{
  uint32_t lead = c->low >> (24 - c->bits_left);
  c->bits_left += 8;
  c->low &= 0xffffffffu >> c->bits_left;
  if (lead == 0xff) {
    c->num_buffered++;
  } else if (c->num_buffered > 0) {
    uint32_t carry = lead >> 8;
    uint32_t byte = (uint32_t)c->buffered_byte + carry;
    c->buffered_byte = (int)(lead & 0xff);
    orc_bits_put(c->out, byte & 0xff, 8);
    byte = (0xff + carry) & 0xff;
    while (c->num_buffered > 1) { orc_bits_put(c->out, byte, 8); c->num_buffered--; }
  } else {
    c->num_buffered = 1;
    c->buffered_byte = (int)lead;
  }
}

static inline void cabac_test_write(orc_cabac_t *c) { if (c->bits_left < 12) cabac_write_out(c); }

void orc_cabac_bin(orc_cabac_t *c, int ctx_idx, int bin)
{
  uint8_t *s = &c->ctx[ctx_idx];
  int st = *s >> 1, mps = *s & 1;
  uint32_t lps = orc_range_tab_lps[st][(c->range >> 6) & 3];
  c->bins++;
  c->range -= lps;
  if (bin != mps) {
    int nb = 0;
    { uint32_t t = lps; while (t < 256) { t <<= 1; nb++; } }   /* renorm shift for the LPS range */
    c->low = (c->low + c->range) << nb;
    c->range = lps << nb;
    if (st == 0) mps = 1 - mps;
    st = orc_trans_idx_lps[st];
    *s = (uint8_t)((st << 1) | mps);
    c->bits_left -= nb;
  } else {
    if (st < 62) st++;
    *s = (uint8_t)((st << 1) | mps);
    if (c->range >= 256) return;
    c->low <<= 1;
    c->range <<= 1;
    c->bits_left--;
  }
  cabac_test_write(c);
}

void orc_cabac_bypass(orc_cabac_t *c, int bin)
{
  c->bins++;
  c->low <<= 1;
  if (bin) c->low += c->range;
  c->bits_left--;
  cabac_test_write(c);
}

void orc_cabac_bypass_bits(orc_cabac_t *c, uint32_t bins, int n)
{
  for (int i = n - 1; i >= 0; i--) orc_cabac_bypass(c, (bins >> i) & 1);
}

void orc_cabac_terminate(orc_cabac_t *c, int bin)
{
  c->bins++;
  c->range -= 2;
  if (bin) {
    c->low += c->range;
    c->low <<= 7;
    c->range = 2 << 7;
    c->bits_left -= 7;
  } else if (c->range >= 256) {
    return;
  } else {
    c->low <<= 1;
    c->range <<= 1;
    c->bits_left--;
  }
  cabac_test_write(c);
}

void orc_cabac_finish(orc_cabac_t *c)
{
  if (c->low >> (32 - c->bits_left)) {
    orc_bits_put(c->out, (uint32_t)(c->buffered_byte + 1) & 0xff, 8);
    while (c->num_buffered > 1) { orc_bits_put(c->out, 0x00, 8); c->num_buffered--; }
    c->low -= 1u << (32 - c->bits_left);
  } else {
    if (c->num_buffered > 0) orc_bits_put(c->out, (uint32_t)c->buffered_byte, 8);
    while (c->num_buffered > 1) { orc_bits_put(c->out, 0xff, 8); c->num_buffered--; }
  }
  orc_bits_put(c->out, c->low >> 8, 24 - c->bits_left);
  orc_bits_trailing(c->out);      /* rbsp_stop_one_bit / alignment_bit_equal_to_one + zeros */
}

/* ---- residual_coding (7.3.8.11, 9.3.4.2.4-7) ------------------------------------------------ */

static void code_last_pos(orc_cabac_t *c, int pos, int log2n, int cidx, int ctx_base)
{
  /* group index of the position: prefix value (binarisation of 9.3.3.x / HM g_uiGroupIdx) */
  static const uint8_t group_idx[32] = {0, 1, 2, 3, 4, 4, 5, 5, 6, 6, 6, 6, 7, 7, 7, 7,
                                        8, 8, 8, 8, 8, 8, 8, 8, 9, 9, 9, 9, 9, 9, 9, 9};
  int offset, shift;
  if (cidx == 0) { offset = 3 * (log2n - 2) + ((log2n - 1) >> 2); shift = (log2n + 1) >> 2; }
  else { offset = 15; shift = log2n - 2; }
  int prefix = group_idx[pos];
  int cmax = (log2n << 1) - 1;
  for (int i = 0; i < prefix; i++) orc_cabac_bin(c, ctx_base + offset + (i >> shift), 1);
  if (prefix < cmax) orc_cabac_bin(c, ctx_base + offset + (prefix >> shift), 0);
}

static void code_last_suffix(orc_cabac_t *c, int pos)
{
  static const uint8_t group_idx[32] = {0, 1, 2, 3, 4, 4, 5, 5, 6, 6, 6, 6, 7, 7, 7, 7,
                                        8, 8, 8, 8, 8, 8, 8, 8, 9, 9, 9, 9, 9, 9, 9, 9};
  static const uint8_t min_in_group[10] = {0, 1, 2, 3, 4, 6, 8, 12, 16, 24};
  int g = group_idx[pos];
  if (g > 3) orc_cabac_bypass_bits(c, (uint32_t)(pos - min_in_group[g]), (g >> 1) - 1);
}

static void code_remaining(orc_cabac_t *c, int value, int rice)
{
  if (value < (3 << rice)) {
    int len = value >> rice;
    orc_cabac_bypass_bits(c, (1u << (len + 1)) - 2, len + 1);
    orc_cabac_bypass_bits(c, (uint32_t)value & ((1u << rice) - 1), rice);
  } else {
    int len = rice;
    value -= 3 << rice;
    while (value >= (1 << len)) { value -= 1 << len; len++; }
    int pre = 3 + len + 1 - rice;
    /* pre can exceed 32 only for levels beyond 16 bit; not reachable here */
    orc_cabac_bypass_bits(c, (uint32_t)((1ull << pre) - 2), pre);
    orc_cabac_bypass_bits(c, (uint32_t)value, len);
  }
}

void orc_code_residual(orc_cabac_t *c, const int16_t *lv, int stride, int log2n, int cidx, int scan_idx)
{
  orc_code_residual2(c, lv, stride, log2n, cidx, scan_idx, 0);
}

void orc_code_residual2(orc_cabac_t *c, const int16_t *lv, int stride, int log2n, int cidx, int scan_idx, int sign_hiding)
{
  const int n = 1 << log2n;
  const int sb_log2 = log2n - 2;           /* grid of 4x4 sub-blocks */
  const int nsb = 1 << (2 * sb_log2);
  /* find the last significant coefficient in scan order */
  int last_sb = -1, last_pos = -1, last_x = 0, last_y = 0;
  for (int i = nsb - 1; i >= 0 && last_sb < 0; i--) {
    int xs, ys;
    orc_scan_pos(scan_idx, sb_log2, i, &xs, &ys);
    for (int p = 15; p >= 0; p--) {
      int xp, yp;
      orc_scan_pos(scan_idx, 2, p, &xp, &yp);
      if (lv[(ys * 4 + yp) * stride + xs * 4 + xp]) {
        last_sb = i; last_pos = p; last_x = xs * 4 + xp; last_y = ys * 4 + yp;
        break;
      }
    }
  }
  if (last_sb < 0) return;                 /* caller must not code an all-zero block */
  {
    int px = last_x, py = last_y;
    if (scan_idx == 2) { int t = px; px = py; py = t; }     /* swapped for vertical scan */
    code_last_pos(c, px, log2n, cidx, CTX_LAST_X);
    code_last_pos(c, py, log2n, cidx, CTX_LAST_Y);
    code_last_suffix(c, px);
    code_last_suffix(c, py);
  }
  uint8_t csbf[8][8];
  memset(csbf, 0, sizeof(csbf));
  int c1 = 1;                              /* greater1 context state carried across sub-blocks */
  (void)n;
  for (int i = last_sb; i >= 0; i--) {
    int xs, ys;
    orc_scan_pos(scan_idx, sb_log2, i, &xs, &ys);
    int right = xs + 1 < (1 << sb_log2) ? csbf[ys][xs + 1] : 0;
    int below = ys + 1 < (1 << sb_log2) ? csbf[ys + 1][xs] : 0;
    int prev_csbf = right | (below << 1);
    /* gather the sub-block in scan order */
    int abs_lv[16], sign[16], sig[16], any = 0;
    for (int p = 0; p < 16; p++) {
      int xp, yp;
      orc_scan_pos(scan_idx, 2, p, &xp, &yp);
      int v = lv[(ys * 4 + yp) * stride + xs * 4 + xp];
      if (i == last_sb && p > last_pos) v = 0;
      abs_lv[p] = abs(v); sign[p] = v < 0; sig[p] = v != 0; any |= sig[p];
    }
    int infer_dc = 0;
    if (i < last_sb && i > 0) {
      orc_cabac_bin(c, CTX_CSBF + (prev_csbf ? 1 : 0) + (cidx ? 2 : 0), any);
      csbf[ys][xs] = (uint8_t)any;
      infer_dc = 1;
    } else {
      csbf[ys][xs] = 1;
      any = 1;
    }
    if (!any) continue;
    /* sig_coeff_flag */
    int start = (i == last_sb) ? last_pos - 1 : 15;
    for (int p = start; p >= 0; p--) {
      if (p == 0 && infer_dc) break;       /* inferred 1 when nothing else in the sub-block was set */
      int xp, yp;
      orc_scan_pos(scan_idx, 2, p, &xp, &yp);
      int xc = xs * 4 + xp, yc = ys * 4 + yp, sctx;
      if (log2n == 2) sctx = orc_sig_ctx_map_4x4[(yc << 2) + xc];
      else if (xc + yc == 0) sctx = 0;
      else {
        if (prev_csbf == 0) sctx = (xp + yp == 0) ? 2 : (xp + yp < 3) ? 1 : 0;
        else if (prev_csbf == 1) sctx = yp == 0 ? 2 : (yp == 1 ? 1 : 0);
        else if (prev_csbf == 2) sctx = xp == 0 ? 2 : (xp == 1 ? 1 : 0);
        else sctx = 2;
        if (cidx == 0) {
          if (xs || ys) sctx += 3;
          sctx += log2n == 3 ? (scan_idx == 0 ? 9 : 15) : 21;
        } else {
          sctx += log2n == 3 ? 9 : 12;
        }
      }
      orc_cabac_bin(c, CTX_SIG + (cidx == 0 ? sctx : 27 + sctx), sig[p]);
      if (sig[p]) infer_dc = 0;
    }
    /* greater1 / greater2 flags */
    int ctx_set = (i > 0 && cidx == 0) ? 2 : 0;
    if (c1 == 0) ctx_set++;
    c1 = 1;
    int num_g1 = 0, first_g1_pos = -1;
    int g1[16] = {0}, g2flag = 0;
    for (int p = 15; p >= 0; p--) {
      if (!sig[p]) continue;
      if (num_g1 < 8) {
        g1[p] = abs_lv[p] > 1;
        orc_cabac_bin(c, CTX_GT1 + (cidx ? 16 : 0) + 4 * ctx_set + c1, g1[p]);
        if (g1[p]) { c1 = 0; if (first_g1_pos < 0) first_g1_pos = p; }
        else if (c1 < 3 && c1 > 0) c1++;
        num_g1++;
      }
    }
    if (first_g1_pos >= 0) {
      g2flag = abs_lv[first_g1_pos] > 2;
      orc_cabac_bin(c, CTX_GT2 + (cidx ? 4 : 0) + ctx_set, g2flag);
    }
    /* signs; with sign data hiding the sign of the first coefficient in scan order is not coded when
     * the significant coefficients of the group span more than three scan positions (7.3.8.11) */
    int first_sig = -1, last_sig = -1;
    for (int p = 0; p < 16; p++) if (sig[p]) { if (first_sig < 0) first_sig = p; last_sig = p; }
    const int hidden = sign_hiding && last_sig - first_sig > 3;
    for (int p = 15; p >= 0; p--)
      if (sig[p] && !(hidden && p == first_sig)) orc_cabac_bypass(c, sign[p]);
    /* remaining levels */
    int num_sig = 0, rice = 0;
    for (int p = 15; p >= 0; p--) {
      if (!sig[p]) continue;
      int base = 1 + (num_sig < 8 ? g1[p] : 0) + (p == first_g1_pos ? g2flag : 0);
      int thresh = num_sig < 8 ? (p == first_g1_pos ? 3 : 2) : 1;
      if (base == thresh) {
        int rem = abs_lv[p] - base;
        code_remaining(c, rem, rice);
        if (abs_lv[p] > (3 << rice)) rice = rice < 4 ? rice + 1 : 4;
      }
      num_sig++;
    }
  }
}
