/* TEST INFRASTRUCTURE ONLY -- see hevc_enc.h.
 *
 * Encoder structure (the algorithm the CUDA encoder mirrors):
 *   I pictures  CTUs in raster order, CUs of 16x16 (8x8 where 16 does not fit) in z-order;
 *               35-mode SAD + lambda*bits search on reconstructed neighbours; chroma = DM.
 *   P pictures  1 reference (previous reconstruction), 2Nx2N inter CUs of 32/16/8:
 *               exhaustive full-sample search (+-range) at 8x8 granularity with SADs summed
 *               to 16x16 / 32x32, bottom-up partition decision, half- then quarter-sample
 *               refinement, residual DCT/quant/recon with TU = CU, then deblocking.
 *               merge / skip / AMVP are resolved at entropy-coding time from the final
 *               motion field, so no reconstruction step depends on a neighbouring CU.
 *   Entropy     CABAC, one substream per CTU row (WPP, entropy_coding_sync), one slice.
 *
 * Reference call sites this stands in for: kvz_api encoder_encode as used at
 * /root/reference/src/media/processing/kvazaarfilter.cpp:435-449 (one AU per call, in order).
 */
#include "hevc_enc.h"
#include "hevc_cabac.h"
#include "hevc_prims.h"
#include "hevc_scaling.h"
#include "hevc_tables.h"

#include <limits.h>
#include <stdlib.h>
#include <string.h>

#define CTB_LOG2 6
#define CTB 64
#define PAD 160                /* luma padding of the reference planes: covers 4 * me_coarse + search_range + interpolation taps */
#define CU_OVERHEAD_BITS 3
#define MAX_MERGE 5
#define MAX_REFS 4
/* intra CUs in P pictures (cfg.intra_in_p): a 16x16 block whose best inter cost (SAD + lambda * bits)
 * exceeds INTRA_TRY_COST gets the 35-mode source-based intra search; intra wins when 1.5 x its cost
 * (the search predicts from source neighbours; the real prediction, from reconstructed ones, is worse
 * and its residual dearer) plus INTRA_OVERHEAD_BITS of signalling is smaller.  Tuned on the synthetic
 * sequences: never worse than +0.8 % bits on `camera`, -0.7 ... -1.6 % on `sports` (scene cut); the
 * same CUs end up intra with INTRA_TRY_COST anywhere in 1024 ... 4096, at a fifth of the searches. */
#define INTRA_TRY_COST 4096
#define INTRA_OVERHEAD_BITS 24

static const uint16_t lambda_q4_tab[52] = {   /* round(16*sqrt(0.57*2^((qp-12)/3))) */
  3, 3, 4, 4, 5, 5, 6, 7, 8, 9, 10, 11, 12, 14, 15, 17, 19, 22, 24, 27, 30, 34, 38, 43, 48, 54, 61, 68,
  77, 86, 97, 108, 122, 137, 153, 172, 193, 217, 244, 273, 307, 344, 387, 434, 487, 547, 614, 689, 773,
  868, 974, 1093};

/* motion of one 8x8 unit of a reference picture, for temporal motion vector prediction */
struct orc_mvf { int16_t mvx, mvy; int16_t ref_poc; uint8_t inter; };

/* SAO parameters of one CTU (7.4.9.3): [0] luma, [1] chroma (type and class shared by Cb / Cr) */
struct orc_sao {
  uint8_t type[2];               /* 0 off, 1 band offset, 2 edge offset */
  uint8_t eo_class[2];
  uint8_t band_pos[3];
  int8_t offset[3][4];
};

struct orc_encoder {
  orc_enc_cfg_t cfg;
  int w, h, cw, ch, w8, h8, ctb_cols, ctb_rows;
  int frame_idx, poc, is_idr;
  int8_t *ctu_dqp;               /* per-CTU QP offset (ROI), zeros by default */
  orc_scaling_t sl;              /* cfg.scaling_list: the lists in force */
  int8_t *vaq_dqp;               /* per-CTU QP offset of variance adaptive quantisation (cfg.vaq), per picture */
  int8_t *ctu_delta;             /* CuQpDeltaVal coded in each CTU (0 if none) */
  uint8_t *ctu_first;            /* z-index (8x8 units) of the first CU with a coded residual, 64 = none */
  struct orc_sao *sao;           /* per CTU, when cfg.sao */
  uint8_t *dbk;                  /* copy of the deblocked picture: SAO reads it, writes rec */
  const uint8_t *src;
  uint8_t *rec, *rec_pre;        /* packed I420 */
  /* decoded picture buffer: dpb[0] is the previous picture (reference index 0), dpb[1] the one before ... */
  struct orc_ref {
    uint8_t *pad[3];             /* padded reconstruction */
    uint8_t *q;                  /* me_coarse: quarter-resolution luma */
    struct orc_mvf *mvf;         /* tmvp: motion field, one entry per 8x8 unit */
    int poc;
  } dpb[MAX_REFS];
  int n_dpb;                     /* pictures in the buffer (reset by an IDR) */
  int n_refs;                    /* of the picture being coded: min(cfg.refs, n_dpb) */
  int delta_coded;               /* entropy stage: the current CTU's cu_qp_delta has been coded */
  int16_t *pen_ctr;              /* per 8x8 unit (origin unit of a CU): centre of its mv penalty, quarter samples */
  uint8_t *src_q;                /* me_coarse: quarter-resolution luma of the source */
  int refstride[3];
  orc_cu_t *cu;
  int16_t *levels;               /* I420-shaped */
  uint8_t *sub;                  /* substream scratch */
  size_t sub_cap;
  unsigned long long bins;
};

static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }
static inline int clip3i(int lo, int hi, int v) { return v < lo ? lo : (v > hi ? hi : v); }

static uint8_t *plane(uint8_t *base, int w, int h, int c)
{
  return c == 0 ? base : base + (size_t)w * h + (c == 2 ? (size_t)(w / 2) * (h / 2) : 0);
}
static int16_t *lplane(int16_t *base, int w, int h, int c)
{
  return c == 0 ? base : base + (size_t)w * h + (c == 2 ? (size_t)(w / 2) * (h / 2) : 0);
}

orc_encoder_t *orc_enc_open(const orc_enc_cfg_t *cfg)
{
  if (!cfg || cfg->width <= 0 || cfg->height <= 0 || (cfg->width & 7) || (cfg->height & 7)) return NULL;
  if (cfg->qp < 0 || cfg->qp > 51 || cfg->search_range < 1 || cfg->search_range > 32) return NULL;
  if (cfg->refs < 0 || cfg->refs > MAX_REFS) return NULL;
  if (cfg->tr_depth < 0 || cfg->tr_depth > 3) return NULL;
  if (cfg->cb_qp_offset < -12 || cfg->cb_qp_offset > 12 || cfg->cr_qp_offset < -12 || cfg->cr_qp_offset > 12) return NULL;
  if (cfg->beta_offset_div2 < -6 || cfg->beta_offset_div2 > 6 || cfg->tc_offset_div2 < -6 || cfg->tc_offset_div2 > 6) return NULL;
  if (cfg->me_coarse < 0 || cfg->me_coarse > 32 || (cfg->me_coarse > 0 && cfg->search_range > 16)) return NULL;
  if (cfg->vaq < 0 || cfg->vaq > 20 || (cfg->vaq && !cfg->qp_delta)) return NULL;
  if (cfg->scaling_list < 0 || cfg->scaling_list > 3) return NULL;
  if (cfg->conf_right < 0 || cfg->conf_right > 6 || (cfg->conf_right & 1) || cfg->conf_bottom < 0 || cfg->conf_bottom > 6 || (cfg->conf_bottom & 1)) return NULL;
  orc_encoder_t *e = (orc_encoder_t *)calloc(1, sizeof(*e));
  if (!e) return NULL;
  e->cfg = *cfg;
  if (cfg->scaling_list >= 2) orc_scaling_test_lists(&e->sl); else orc_scaling_default(&e->sl);
  e->w = cfg->width; e->h = cfg->height; e->cw = e->w / 2; e->ch = e->h / 2;
  e->w8 = e->w / 8; e->h8 = e->h / 8;
  e->ctb_cols = (e->w + CTB - 1) / CTB; e->ctb_rows = (e->h + CTB - 1) / CTB;
  e->ctu_dqp = (int8_t *)calloc((size_t)e->ctb_cols * e->ctb_rows, 1);
  e->vaq_dqp = (int8_t *)calloc((size_t)e->ctb_cols * e->ctb_rows, 1);
  e->ctu_delta = (int8_t *)calloc((size_t)e->ctb_cols * e->ctb_rows, 1);
  e->ctu_first = (uint8_t *)calloc((size_t)e->ctb_cols * e->ctb_rows, 1);
  if (cfg->sao) {
    e->sao = (struct orc_sao *)calloc((size_t)e->ctb_cols * e->ctb_rows, sizeof(struct orc_sao));
    e->dbk = (uint8_t *)malloc((size_t)e->w * e->h * 3 / 2);
  }
  size_t fsz = (size_t)e->w * e->h * 3 / 2;
  e->rec = (uint8_t *)calloc(fsz, 1);
  e->rec_pre = (uint8_t *)calloc(fsz, 1);
  if (e->cfg.refs < 1) e->cfg.refs = 1;
  for (int r = 0; r < e->cfg.refs; r++) {
    for (int c = 0; c < 3; c++) {
      int pw = c ? e->cw : e->w, ph = c ? e->ch : e->h, pad = c ? PAD / 2 : PAD;
      e->refstride[c] = pw + 2 * pad;
      e->dpb[r].pad[c] = (uint8_t *)calloc((size_t)e->refstride[c] * (ph + 2 * pad), 1);
    }
    e->dpb[r].q = (uint8_t *)calloc((size_t)(e->w / 4) * (e->h / 4), 1);
    e->dpb[r].mvf = (struct orc_mvf *)calloc((size_t)e->w8 * e->h8, sizeof(struct orc_mvf));
  }
  e->pen_ctr = (int16_t *)calloc((size_t)e->w8 * e->h8 * 2, sizeof(int16_t));
  e->src_q = (uint8_t *)calloc((size_t)(e->w / 4) * (e->h / 4), 1);
  e->cu = (orc_cu_t *)calloc((size_t)e->w8 * e->h8, sizeof(orc_cu_t));
  e->levels = (int16_t *)calloc(fsz, sizeof(int16_t));
  e->sub_cap = fsz * 2 + 65536;
  e->sub = (uint8_t *)malloc(e->sub_cap);
  return e;
}

void orc_enc_close(orc_encoder_t *e)
{
  if (e) { free(e->ctu_dqp); free(e->vaq_dqp); free(e->ctu_delta); free(e->ctu_first); free(e->sao); free(e->dbk); }
  if (!e) return;
  free(e->rec); free(e->rec_pre);
  for (int r = 0; r < MAX_REFS; r++) {
    for (int c = 0; c < 3; c++) free(e->dpb[r].pad[c]);
    free(e->dpb[r].q); free(e->dpb[r].mvf);
  }
  free(e->cu); free(e->levels); free(e->sub); free(e->src_q); free(e->pen_ctr);
  free(e);
}

#ifdef _OPENMP
#include <omp.h>
#endif
/* Threads used by the parallel loops of this library (cpu_baseline / reference arm of bench.py). */
void orc_set_threads(int n)
{
#ifdef _OPENMP
  omp_set_num_threads(n < 1 ? 1 : n);
#else
  (void)n;
#endif
}
int orc_max_threads(void)
{
#ifdef _OPENMP
  return omp_get_num_procs();
#else
  return 1;
#endif
}

const uint8_t *orc_enc_recon(const orc_encoder_t *e) { return e->rec; }
const uint8_t *orc_enc_recon_predeblock(const orc_encoder_t *e) { return e->rec_pre; }
const orc_cu_t *orc_enc_cu_map(const orc_encoder_t *e) { return e->cu; }
const int16_t *orc_enc_levels(const orc_encoder_t *e) { return e->levels; }
int orc_enc_last_was_idr(const orc_encoder_t *e) { return e->is_idr; }
unsigned long long orc_enc_bins(const orc_encoder_t *e) { return e->bins; }
/* QP of the CTU that holds luma sample (x, y): the slice QP plus the ROI offset plus the offset of
 * variance adaptive quantisation */
static int ctu_qp(const orc_encoder_t *e, int x, int y)
{
  if (!e->cfg.qp_delta) return e->cfg.qp;
  const int i = (y / CTB) * e->ctb_cols + x / CTB;
  return clip3i(0, 51, e->cfg.qp + e->ctu_dqp[i] + e->vaq_dqp[i]);
}

/* ---- variance adaptive quantisation ("vaq", Kvazaar's --vaq <strength>, set by the reference at
 * kvazaarfilter.cpp:280-284 when the user setting is 1..20) ----
 * Kvazaar (encoderstate.c, the "Variance adaptive quantization" block of encoder_state_init_new_frame;
 * third-party, not under /root/reference) gives every CTU the offset
 *     dqp = strength * 0.1 * (ln(max(var_ctu, 4)) - ln(var_picture))
 * (here the picture's variance has the same floor of 4, so that a flat picture moves nothing),
 * var = luma variance + the two chroma variances of the CTU / the picture, and quantises the CTU at
 * the picture QP + round(dqp).  Restated here in integer arithmetic so that the CPU and the GPU agree
 * bit for bit: variances as exact fractions (n * sum x^2 - (sum x)^2) / n^2 over a common
 * denominator, log2 in Q8 fixed point by repeated squaring of a 32-bit mantissa (truncating), and
 *     dqp = clip(round_half_away(strength * 71 * (L_ctu - L_picture) / 2^18), -12, 12)
 * (0.1 * ln 2 / 256 = 71 / 2^18 to four digits; the clip bounds the swing of a CTU that is flat
 * inside a busy picture). */
typedef unsigned __int128 u128;

/* floor(256 * log2(v)), v > 0 */
static int log2_q8(u128 v)
{
  int msb = 127;
  while (!(v >> msb)) msb--;
  uint64_t m = msb >= 31 ? (uint64_t)(v >> (msb - 31)) : (uint64_t)(v << (31 - msb));    /* [2^31, 2^32) */
  int frac = 0;
  for (int i = 0; i < 8; i++) {
    m = (m * m) >> 31;                               /* [2^31, 2^33) */
    frac <<= 1;
    if (m >> 32) { frac |= 1; m >>= 1; }
  }
  return msb * 256 + frac;
}

/* n^2 * (var Y + var U + var V) of a region of n luma and n / 4 samples per chroma plane */
static u128 var_numer(uint64_t n, const uint64_t s[3], const uint64_t ss[3])
{
  const uint64_t nc = n / 4;
  u128 y = (u128)n * ss[0] - (u128)s[0] * s[0];
  u128 u = (u128)nc * ss[1] - (u128)s[1] * s[1];
  u128 v = (u128)nc * ss[2] - (u128)s[2] * s[2];
  return y + 16 * (u + v);
}

static void region_sums(const uint8_t *i420, int w, int h, int x0, int y0, int bw, int bh, uint64_t s[3], uint64_t ss[3])
{
  for (int c = 0; c < 3; c++) {
    const uint8_t *pl = c == 0 ? i420 : i420 + (size_t)w * h + (size_t)(c - 1) * (w / 2) * (h / 2);
    const int st = c ? w / 2 : w, sh = c ? 1 : 0;
    uint64_t a = 0, b = 0;
    for (int y = y0 >> sh; y < (y0 + bh) >> sh; y++)
      for (int x = x0 >> sh; x < (x0 + bw) >> sh; x++) { const unsigned v = pl[(size_t)y * st + x]; a += v; b += v * v; }
    s[c] = a; ss[c] = b;
  }
}

/* per-CTU QP offsets (raster, (w+63)/64 x (h+63)/64) of one I420 picture; w and h even */
void orc_vaq_offsets(const uint8_t *i420, int w, int h, int strength, int8_t *out)
{
  const int cols = (w + CTB - 1) / CTB, rows = (h + CTB - 1) / CTB;
  uint64_t fs[3] = {0, 0, 0}, fss[3] = {0, 0, 0};
  uint64_t *cs = (uint64_t *)malloc(sizeof(uint64_t) * 6 * (size_t)cols * rows);
  for (int r = 0; r < rows; r++)
    for (int c = 0; c < cols; c++) {
      uint64_t *q = cs + 6 * ((size_t)r * cols + c);
      region_sums(i420, w, h, c * CTB, r * CTB, imin(CTB, w - c * CTB), imin(CTB, h - r * CTB), q, q + 3);
      for (int k = 0; k < 3; k++) { fs[k] += q[k]; fss[k] += q[3 + k]; }
    }
  const uint64_t fn = (uint64_t)w * h;
  u128 fnum = var_numer(fn, fs, fss);
  if (fnum < (u128)4 * fn * fn) fnum = (u128)4 * fn * fn;      /* the floor of the CTUs: a flat picture moves nothing */
  const int lf = log2_q8(fnum) - 2 * log2_q8(fn);
  for (int r = 0; r < rows; r++)
    for (int c = 0; c < cols; c++) {
      const uint64_t *q = cs + 6 * ((size_t)r * cols + c);
      const uint64_t n = (uint64_t)imin(CTB, w - c * CTB) * imin(CTB, h - r * CTB);
      u128 num = var_numer(n, q, q + 3);
      if (num < (u128)4 * n * n) num = (u128)4 * n * n;
      const int lc = log2_q8(num) - 2 * log2_q8(n);
      const int t = strength * 71 * (lc - lf);
      const int d = t >= 0 ? (t + (1 << 17)) >> 18 : -((-t + (1 << 17)) >> 18);
      out[r * cols + c] = (int8_t)clip3i(-12, 12, d);
    }
  free(cs);
}
static int lambda_at(const orc_encoder_t *e, int x, int y) { return lambda_q4_tab[ctu_qp(e, x, y)]; }

/* slice QP of the following pictures (what a rate controller or a GOP structure with QP offsets sets per picture) */
int orc_enc_set_qp(orc_encoder_t *e, int qp)
{
  if (!e || qp < 0 || qp > 51) return -1;
  e->cfg.qp = qp;
  return 0;
}

int orc_enc_set_ctu_dqp(orc_encoder_t *e, const int8_t *dqp)
{
  if (!e || !e->cfg.qp_delta) return -1;
  if (dqp) memcpy(e->ctu_dqp, dqp, (size_t)e->ctb_cols * e->ctb_rows);
  else memset(e->ctu_dqp, 0, (size_t)e->ctb_cols * e->ctb_rows);
  return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* residual path shared by intra and inter: src - pred -> DCT -> Q -> (IQ -> IDCT) -> recon     */

/* QP of plane c at luma position (x, y): Table 8-10 with the PPS chroma offsets */
static int plane_qp(const orc_encoder_t *e, int c, int x, int y)
{
  const int qpy = ctu_qp(e, x, y);
  return c == 0 ? qpy : orc_chroma_qp(qpy + (c == 1 ? e->cfg.cb_qp_offset : e->cfg.cr_qp_offset));
}

/* Sign data hiding, encoder side: in every 4x4 group whose significant levels span more than three
 * scan positions the decoder takes the sign of the first one (lowest scan position) from the parity of
 * the sum of magnitudes (odd = negative).  Where the parity is wrong one magnitude between the first
 * and the last significant position changes by one -- the change with the smallest increase of the
 * quantisation error; the first and the last position never become zero, so the span is unchanged. */
static int sign_hide(const int16_t *coef, int16_t *level, int log2n, int qp, int scan_idx, const orc_scaling_t *sl, int matrix)
{
  const int n = 1 << log2n, nsb = 1 << (log2n - 2);
  const int qbits = 14 + qp / 6 + (15 - 8 - log2n);
  int nz = 0;
  for (int ys = 0; ys < nsb; ys++)
    for (int xs = 0; xs < nsb; xs++) {
      int idx[16], first = -1, last = -1, sum = 0;
      for (int p = 0; p < 16; p++) {
        int xp, yp;
        orc_scan_pos(scan_idx, 2, p, &xp, &yp);
        idx[p] = (ys * 4 + yp) * n + xs * 4 + xp;
        if (level[idx[p]]) { if (first < 0) first = p; last = p; sum += abs(level[idx[p]]); }
      }
      if (first >= 0 && last - first > 3 && (sum & 1) != (level[idx[first]] < 0)) {
        int64_t best = INT64_MAX;
        int bp = last, bd = 1;
        for (int p = last; p >= first; p--) {
          const int l = abs(level[idx[p]]);
          const int64_t scale = sl ? (orc_quant_scales[qp % 6] << 4) / orc_scaling_factor(sl, log2n, matrix, idx[p] & (n - 1), idx[p] >> log2n)
                                   : orc_quant_scales[qp % 6];
          const int64_t t = (int64_t)abs(coef[idx[p]]) * scale, d0 = llabs(t - ((int64_t)l << qbits));
          for (int d = 1; d >= -1; d -= 2) {
            if (l + d < 0 || l + d > 32767) continue;
            if (l + d == 0 && (p == first || p == last)) continue;
            const int64_t cost = llabs(t - ((int64_t)(l + d) << qbits)) - d0;
            if (cost < best) { best = cost; bp = p; bd = d; }
          }
        }
        const int l = abs(level[idx[bp]]) + bd;
        const int neg = level[idx[bp]] ? level[idx[bp]] < 0 : coef[idx[bp]] < 0;
        level[idx[bp]] = (int16_t)(neg ? -l : l);
      }
    }
  for (int i = 0; i < n * n; i++) nz += level[i] != 0;
  return nz;
}

/* `pred`: prediction samples with row pitch `ps`.  rd (may be NULL): adds the squared error of the
 * reconstruction and a bit estimate of the levels (what the transform-tree decision compares).
 * dst: DST-VII (4x4 intra luma) instead of the DCT.  scan_idx: coefficient scan of the block (sign hiding). */
typedef struct { long long sse; int bits; } tb_rd_t;

static int recon_tb_x(orc_encoder_t *e, int c, int x0, int y0, int log2n, const uint8_t *pred, int ps, tb_rd_t *rd, int dst, int scan_idx, int intra)
{
  const orc_scaling_t *sl = e->cfg.scaling_list ? &e->sl : NULL;
  const int matrix = (intra ? 0 : 3) + c;
  const int n = 1 << log2n;
  const int pw = c ? e->cw : e->w;
  const uint8_t *src = plane((uint8_t *)e->src, e->w, e->h, c);
  uint8_t *rec = plane(e->rec, e->w, e->h, c);
  int16_t *lv = lplane(e->levels, e->w, e->h, c);
  int16_t resid[32 * 32], coef[32 * 32], level[32 * 32];
  const int qp = c ? plane_qp(e, c, x0 * 2, y0 * 2) : plane_qp(e, 0, x0, y0);
  for (int y = 0; y < n; y++)
    for (int x = 0; x < n; x++)
      resid[y * n + x] = (int16_t)((int)src[(size_t)(y0 + y) * pw + x0 + x] - (int)pred[y * ps + x]);
  if (dst) orc_fdst4(resid, coef); else orc_fdct(resid, coef, log2n);
  int nz = sl ? orc_quant_sl(coef, level, log2n, qp, e->is_idr, sl, matrix) : orc_quant(coef, level, log2n, qp, e->is_idr);
  if (nz && e->cfg.sign_hiding) nz = sign_hide(coef, level, log2n, qp, scan_idx, sl, matrix);
  for (int y = 0; y < n; y++) memcpy(lv + (size_t)(y0 + y) * pw + x0, level + y * n, n * sizeof(int16_t));
  if (nz) {
    if (sl) orc_dequant_sl(level, coef, log2n, qp, sl, matrix); else orc_dequant(level, coef, log2n, qp);
    if (dst) orc_idst4(coef, resid); else orc_idct(coef, resid, log2n);
    for (int y = 0; y < n; y++)
      for (int x = 0; x < n; x++)
        rec[(size_t)(y0 + y) * pw + x0 + x] = (uint8_t)clip3i(0, 255, pred[y * ps + x] + resid[y * n + x]);
  } else {
    for (int y = 0; y < n; y++) memcpy(rec + (size_t)(y0 + y) * pw + x0, pred + y * ps, n);
  }
  if (rd) {
    for (int y = 0; y < n; y++)
      for (int x = 0; x < n; x++) {
        const int d = (int)src[(size_t)(y0 + y) * pw + x0 + x] - (int)rec[(size_t)(y0 + y) * pw + x0 + x];
        rd->sse += d * d;
      }
    rd->bits += 1;                                               /* the coded block flag */
    for (int i = 0; i < n * n; i++)
      if (level[i]) { int a = abs(level[i]), l = 0; while (a >> (l + 1)) l++; rd->bits += 2 * l + 3; }
  }
  return nz != 0;
}

/* ---- transform tree (7.3.8.8) ------------------------------------------------------------------
 * cfg.tr_depth = max_transform_hierarchy_depth_inter = _intra.  A transform unit of a CU may be
 * split into four (down to 8x8 luma, with cfg.tu4 down to four 4x4 luma blocks + one 4x4 block per
 * chroma plane); the decision compares squared error + lambda * estimated bits of both alternatives,
 * bottom-up.  The tree is kept in the cu map: every 8x8 unit records the size of the transform unit
 * that covers it and that unit's coded block flags (tu_log2 = 2: bits 4..7 = the four luma blocks). */

/* what a transform unit predicts with.  pred[0] != NULL: inter, the CU's prediction with origin (x0, y0)
 * luma.  Else intra: luma mode(s) (four when the CU is NxN, else all equal) and the chroma mode. */
typedef struct { const uint8_t *pred[3]; int ps[3]; int x0, y0; int mode[4], chroma_mode; } cu_pred_t;

static void gather_refs(const orc_encoder_t *e, int c, int x0, int y0, int n, uint8_t *refs);
static int scan_idx_for(int pred_mode, int intra_mode, int log2n, int cidx);

/* one transform block of plane c at plane position (px, py) */
static int tu_block(orc_encoder_t *e, int c, int px, int py, int l2, const cu_pred_t *ip, int mode, tb_rd_t *rd)
{
  const int sh = c ? 1 : 0, n = 1 << l2;
  if (ip->pred[0])
    return recon_tb_x(e, c, px, py, l2, ip->pred[c] + (size_t)(py - (ip->y0 >> sh)) * ip->ps[c] + (px - (ip->x0 >> sh)), ip->ps[c], rd, 0, 0, 0);
  uint8_t refs[4 * 32 + 1], pred[32 * 32];
  gather_refs(e, c, px, py, n, refs);
  orc_intra_predict2(refs, l2, mode, c, e->cfg.strong_intra, pred, n);
  return recon_tb_x(e, c, px, py, l2, pred, n, rd, c == 0 && l2 == 2, scan_idx_for(1, mode, l2, c), 1);
}

/* one transform unit, not split: luma block of 1 << log2tu, chroma blocks of half that.  Returns the cbf bits. */
static int tu_leaf(orc_encoder_t *e, int x0, int y0, int log2tu, const cu_pred_t *ip, tb_rd_t *rd)
{
  int cbf = 0;
  for (int c = 0; c < 3; c++) {
    const int sh = c ? 1 : 0;
    cbf |= tu_block(e, c, x0 >> sh, y0 >> sh, log2tu - sh, ip, c ? ip->chroma_mode : ip->mode[0], rd) << c;
  }
  const int n8 = 1 << (log2tu - 3);
  for (int j = 0; j < n8; j++)
    for (int i = 0; i < n8; i++) {
      orc_cu_t *u = &e->cu[(size_t)(y0 / 8 + j) * e->w8 + x0 / 8 + i];
      u->tu_log2 = (uint8_t)log2tu; u->cbf = (uint8_t)cbf;
    }
  return cbf;
}

/* an 8x8 transform unit split into four 4x4 luma blocks (in z order, each predicted from the
 * reconstruction of the ones before it) and, after the fourth, one 4x4 block per chroma plane */
static int tu_leaf4(orc_encoder_t *e, int x0, int y0, const cu_pred_t *ip, tb_rd_t *rd)
{
  int cbf = 0;
  for (int b = 0; b < 4; b++)
    if (tu_block(e, 0, x0 + 4 * (b & 1), y0 + 4 * (b >> 1), 2, ip, ip->mode[b], rd)) cbf |= 1 | (16 << b);
  for (int c = 1; c < 3; c++) cbf |= tu_block(e, c, x0 >> 1, y0 >> 1, 2, ip, ip->chroma_mode, rd) << c;
  orc_cu_t *u = &e->cu[(size_t)(y0 / 8) * e->w8 + x0 / 8];
  u->tu_log2 = 2; u->cbf = (uint8_t)cbf;
  return cbf;
}

static long long tu_node(orc_encoder_t *e, int x0, int y0, int log2tu, int depth, const cu_pred_t *ip)
{
  const int lq = lambda_at(e, x0, y0);
  const long long lam = (lq * lq + 128) >> 8;                             /* lambda, SSE domain */
  tb_rd_t rd = {0, 0};
  tu_leaf(e, x0, y0, log2tu, ip, &rd);
  long long cost0 = rd.sse + lam * (rd.bits + 1);
  if (depth >= e->cfg.tr_depth || log2tu < 3 + !e->cfg.tu4) return cost0;
  /* keep the unsplit result, try the split, take the cheaper */
  const int n = 1 << log2tu, n8 = n / 8;
  uint8_t save_rec[32 * 32 * 3 / 2];
  int16_t save_lv[32 * 32 * 3 / 2];
  orc_cu_t save_cu[16];
  size_t o = 0;
  for (int c = 0; c < 3; c++) {
    const int sh = c ? 1 : 0, pw = c ? e->cw : e->w, m = n >> sh;
    uint8_t *rec = plane(e->rec, e->w, e->h, c) + (size_t)(y0 >> sh) * pw + (x0 >> sh);
    int16_t *lv = lplane(e->levels, e->w, e->h, c) + (size_t)(y0 >> sh) * pw + (x0 >> sh);
    for (int y = 0; y < m; y++) { memcpy(save_rec + o, rec + (size_t)y * pw, m); memcpy(save_lv + o, lv + (size_t)y * pw, m * sizeof(int16_t)); o += m; }
  }
  for (int j = 0; j < n8; j++) for (int i = 0; i < n8; i++) save_cu[j * n8 + i] = e->cu[(size_t)(y0 / 8 + j) * e->w8 + x0 / 8 + i];
  long long cost1 = lam;
  if (log2tu == 3) {
    tb_rd_t rd4 = {0, 0};
    tu_leaf4(e, x0, y0, ip, &rd4);
    cost1 += rd4.sse + lam * (rd4.bits + 1);
  } else {
    for (int q = 0; q < 4; q++) cost1 += tu_node(e, x0 + (q & 1) * n / 2, y0 + (q >> 1) * n / 2, log2tu - 1, depth + 1, ip);
  }
  if (cost1 < cost0) return cost1;
  o = 0;
  for (int c = 0; c < 3; c++) {
    const int sh = c ? 1 : 0, pw = c ? e->cw : e->w, m = n >> sh;
    uint8_t *rec = plane(e->rec, e->w, e->h, c) + (size_t)(y0 >> sh) * pw + (x0 >> sh);
    int16_t *lv = lplane(e->levels, e->w, e->h, c) + (size_t)(y0 >> sh) * pw + (x0 >> sh);
    for (int y = 0; y < m; y++) { memcpy(rec + (size_t)y * pw, save_rec + o, m); memcpy(lv + (size_t)y * pw, save_lv + o, m * sizeof(int16_t)); o += m; }
  }
  for (int j = 0; j < n8; j++) for (int i = 0; i < n8; i++) e->cu[(size_t)(y0 / 8 + j) * e->w8 + x0 / 8 + i] = save_cu[j * n8 + i];
  return cost0;
}

/* after the tree of a CU: bit 1 of `flags` on all its units = the CU has a coded residual (rqt_root_cbf) */
static int cu_root_cbf(orc_encoder_t *e, int x0, int y0, int log2)
{
  const int n8 = 1 << (log2 - 3);
  int any = 0;
  for (int j = 0; j < n8; j++) for (int i = 0; i < n8; i++) any |= e->cu[(size_t)(y0 / 8 + j) * e->w8 + x0 / 8 + i].cbf;
  for (int j = 0; j < n8; j++)
    for (int i = 0; i < n8; i++) {
      orc_cu_t *u = &e->cu[(size_t)(y0 / 8 + j) * e->w8 + x0 / 8 + i];
      u->flags = (uint8_t)((u->flags & ~2) | (any ? 2 : 0));
    }
  return any;
}

static void set_cu(orc_encoder_t *e, int x0, int y0, int log2, const orc_cu_t *v)
{
  int n8 = 1 << (log2 - 3);
  for (int j = 0; j < n8; j++)
    for (int i = 0; i < n8; i++) e->cu[(size_t)(y0 / 8 + j) * e->w8 + x0 / 8 + i] = *v;
}

/* ------------------------------------------------------------------------------------------ */
/* intra                                                                                          */

static unsigned zorder8(int x8, int y8)      /* z-scan index of an 8x8 unit inside its CTB */
{
  unsigned z = 0;
  for (int b = 0; b < 3; b++) z |= (unsigned)((x8 >> b) & 1) << (2 * b) | (unsigned)((y8 >> b) & 1) << (2 * b + 1);
  return z;
}
static unsigned coding_order(const orc_encoder_t *e, int x, int y)
{
  return (unsigned)((y >> CTB_LOG2) * e->ctb_cols + (x >> CTB_LOG2)) * 64 + zorder8((x >> 3) & 7, (y >> 3) & 7);
}
/* 6.4.1 availability of luma location (x, y) for the block at (xc, yc): inside the picture and
 * earlier in coding order.  (In a P picture the inter CUs are all reconstructed before the intra
 * CUs are visited, but a later CU is still unavailable -- the decoder has not seen it yet.) */
static unsigned coding_order4(const orc_encoder_t *e, int x, int y)      /* ... at 4x4 granularity (blocks of an NxN CU / a split 8x8 unit) */
{
  return coding_order(e, x, y) * 4 + (unsigned)(((y >> 2) & 1) << 1 | ((x >> 2) & 1));
}
static int avail_luma(const orc_encoder_t *e, int xc, int yc, int x, int y)
{
  if (x < 0 || y < 0 || x >= e->w || y >= e->h) return 0;
  return coding_order4(e, x, y) < coding_order4(e, xc, yc);
}

/* 8.4.4.2.2: gather the 4N+1 neighbours (layout of orc_intra_predict) with substitution */
static void gather_refs_from(const orc_encoder_t *e, const uint8_t *frame, int c, int x0, int y0, int n, uint8_t *refs)
{
  const int pw = c ? e->cw : e->w;
  const int sh = c ? 1 : 0;
  const uint8_t *rec = plane((uint8_t *)frame, e->w, e->h, c);
  uint8_t av[4 * 32 + 1];
  int any = 0;
  for (int k = 0; k < 2 * n; k++) {
    int x = x0 - 1, y = y0 + 2 * n - 1 - k;
    av[k] = (uint8_t)avail_luma(e, x0 << sh, y0 << sh, x << sh, y << sh);
    if (av[k]) refs[k] = rec[(size_t)y * pw + x];
  }
  av[2 * n] = (uint8_t)avail_luma(e, x0 << sh, y0 << sh, (x0 - 1) << sh, (y0 - 1) << sh);
  if (av[2 * n]) refs[2 * n] = rec[(size_t)(y0 - 1) * pw + x0 - 1];
  for (int k = 0; k < 2 * n; k++) {
    int x = x0 + k, y = y0 - 1;
    av[2 * n + 1 + k] = (uint8_t)avail_luma(e, x0 << sh, y0 << sh, x << sh, y << sh);
    if (av[2 * n + 1 + k]) refs[2 * n + 1 + k] = rec[(size_t)y * pw + x];
  }
  for (int k = 0; k <= 4 * n; k++) any |= av[k];
  if (!any) { memset(refs, 128, 4 * n + 1); return; }
  if (!av[0]) {
    int k = 1;
    while (!av[k]) k++;
    refs[0] = refs[k];
  }
  for (int k = 1; k <= 4 * n; k++)
    if (!av[k]) refs[k] = refs[k - 1];
}

static void gather_refs(const orc_encoder_t *e, int c, int x0, int y0, int n, uint8_t *refs)
{
  gather_refs_from(e, e->rec, c, x0, y0, n, refs);
}

/* Mode decision on SOURCE neighbours (same availability and substitution rules as the real
 * prediction): it does not depend on any reconstruction, so the GPU takes it for every CU of the
 * picture in one parallel pass and the reconstruction wavefront only predicts the chosen mode.
 * Cost = SAD + lambda * bits with a fixed prior (planar / DC / vertical cheap) because the
 * neighbours' modes -- hence the MPM list -- are not known in a parallel pass. */
static uint32_t intra_mode_search_x(const orc_encoder_t *e, int x0, int y0, int log2, int *mode_out, int satd)
{
  const int n = 1 << log2;
  uint8_t refs[4 * 32 + 1], pred[32 * 32];
  gather_refs_from(e, e->src, 0, x0, y0, n, refs);
  uint32_t best_cost = UINT_MAX;
  int best_mode = 0;
  for (int mode = 0; mode < 35; mode++) {
    orc_intra_predict2(refs, log2, mode, 0, e->cfg.strong_intra, pred, n);
    /* cfg.intra_satd (I pictures, blocks of 8x8 and larger): Hadamard SATD of the residual, the measure
     * Kvazaar / HM use in the rough mode search, instead of the SAD */
    uint32_t sad = satd ? orc_satd(e->src + (size_t)y0 * e->w + x0, e->w, pred, n, n, n)
                        : orc_sad(e->src + (size_t)y0 * e->w + x0, e->w, pred, n, n, n);
    int bits = (mode == 0 || mode == 1 || mode == 26) ? 2 : 6;
    uint32_t cost = sad + (uint32_t)((lambda_at(e, x0, y0) * bits) >> 4);
    if (cost < best_cost) { best_cost = cost; best_mode = mode; }
  }
  *mode_out = best_mode;
  return best_cost;
}

static uint32_t intra_mode_search(const orc_encoder_t *e, int x0, int y0, int log2, int *mode_out)
{
  return intra_mode_search_x(e, x0, y0, log2, mode_out, 0);
}

/* cfg.chroma_modes: intra_chroma_pred_mode of a CU by SAD on source neighbours of both chroma planes --
 * planar, vertical, horizontal, DC (each replaced by mode 34 where it equals the luma mode, 8.4.3) or
 * the luma mode itself (derived, one bin instead of three) */
static int intra_chroma_search(const orc_encoder_t *e, int x0, int y0, int log2, int luma_mode)
{
  static const uint8_t base[4] = {0, 26, 10, 1};
  const int l2 = log2 - 1, n = 1 << l2;
  uint8_t refs[2][4 * 32 + 1], pred[16 * 16];
  for (int c = 1; c < 3; c++) gather_refs_from(e, e->src, c, x0 / 2, y0 / 2, n, refs[c - 1]);
  uint32_t best = UINT_MAX;
  int best_mode = luma_mode;
  for (int k = 4; k >= 0; k--) {                      /* derived first: it wins ties */
    const int mode = k == 4 ? luma_mode : (base[k] == luma_mode ? 34 : base[k]);
    uint32_t cost = (uint32_t)((lambda_at(e, x0, y0) * (k == 4 ? 1 : 3)) >> 4);
    for (int c = 1; c < 3; c++) {
      orc_intra_predict2(refs[c - 1], l2, mode, c, 0, pred, n);
      cost += orc_sad(plane((uint8_t *)e->src, e->w, e->h, c) + (size_t)(y0 / 2) * e->cw + x0 / 2, e->cw, pred, n, n, n);
    }
    if (cost < best) { best = cost; best_mode = mode; }
  }
  return best_mode;
}

/* prediction of the given mode(s) from RECONSTRUCTED neighbours, residual, reconstruction, cu map;
 * prediction and reconstruction go transform block by transform block (8.4.4.1).  nxn: four 4x4
 * prediction blocks with modes[0..3] (8x8 CUs only), else modes[0] for the whole CU. */
static void intra_cu_recon_x(orc_encoder_t *e, int x0, int y0, int log2, const int *modes, int nxn)
{
  orc_cu_t cu;
  memset(&cu, 0, sizeof(cu));
  cu.log2_size = (uint8_t)log2; cu.pred_mode = 1; cu.intra_mode = (uint8_t)modes[0];
  cu.merge_idx = 0xff; cu.tu_log2 = (uint8_t)log2;
  cu_pred_t ip;
  memset(&ip, 0, sizeof(ip));
  for (int b = 0; b < 4; b++) ip.mode[b] = nxn ? modes[b] : modes[0];
  ip.chroma_mode = e->cfg.chroma_modes ? intra_chroma_search(e, x0, y0, log2, modes[0]) : modes[0];
  cu.chroma_mode = (uint8_t)ip.chroma_mode;
  if (nxn) {
    cu.flags = 1;
    cu.mvx = (int16_t)(modes[1] | (modes[2] << 8)); cu.mvy = (int16_t)modes[3];
  }
  set_cu(e, x0, y0, log2, &cu);
  if (nxn) tu_leaf4(e, x0, y0, &ip, NULL);                /* IntraSplitFlag: the split is inferred */
  else tu_node(e, x0, y0, log2, 0, &ip);
  cu_root_cbf(e, x0, y0, log2);
}

static void intra_cu_recon(orc_encoder_t *e, int x0, int y0, int log2, int best_mode)
{
  intra_cu_recon_x(e, x0, y0, log2, &best_mode, 0);
}

/* I pictures.  Default: 16x16 CUs (8x8 where 16 does not fit the picture).  cfg.intra_sizes widens the
 * choice bottom-up by the search costs (source neighbours): NxN vs 2Nx2N in 8x8 CUs, four 8x8 CUs vs
 * one 16x16, four 16x16 vs one 32x32; every CU pays CU_OVERHEAD_BITS.  The tree is decided first, then
 * coded in z order. */
typedef struct { uint32_t cost; int8_t log2, nxn; uint8_t mode[4]; } intra_plan_t;

static uint32_t intra_plan(orc_encoder_t *e, int x0, int y0, int log2, intra_plan_t *plan /* per 8x8 unit of the CTU */)
{
  if (x0 >= e->w || y0 >= e->h) return 0;
  const int n = 1 << log2, fits = x0 + n <= e->w && y0 + n <= e->h;
  const uint32_t ovh = (uint32_t)((lambda_at(e, x0, y0) * CU_OVERHEAD_BITS) >> 4);
  intra_plan_t *me = &plan[((y0 & (CTB - 1)) >> 3) * 8 + ((x0 & (CTB - 1)) >> 3)];
  if (log2 == 3) {
    int mode;
    me->cost = intra_mode_search_x(e, x0, y0, 3, &mode, e->cfg.intra_satd) + ovh;
    me->log2 = 3; me->nxn = 0; me->mode[0] = (uint8_t)mode;
    if (e->cfg.intra_sizes & 4) {
      uint32_t c4 = ovh + (uint32_t)((lambda_at(e, x0, y0) * 2) >> 4);
      int m4[4];
      for (int b = 0; b < 4; b++) c4 += intra_mode_search(e, x0 + 4 * (b & 1), y0 + 4 * (b >> 1), 2, &m4[b]);
      if (c4 < me->cost) { me->cost = c4; me->nxn = 1; for (int b = 0; b < 4; b++) me->mode[b] = (uint8_t)m4[b]; }
    }
    return me->cost;
  }
  const int can_split = log2 > 4 || (e->cfg.intra_sizes & 1) || !fits;
  const int can_whole = fits && (log2 == 4 || (log2 == 5 && (e->cfg.intra_sizes & 2)));
  uint32_t split_cost = UINT_MAX;
  if (can_split) {
    split_cost = 0;
    for (int q = 0; q < 4; q++) split_cost += intra_plan(e, x0 + (q & 1) * n / 2, y0 + (q >> 1) * n / 2, log2 - 1, plan);
  }
  if (can_whole) {
    int mode;
    const uint32_t whole = intra_mode_search_x(e, x0, y0, log2, &mode, e->cfg.intra_satd) + ovh;
    if (whole <= split_cost) {
      me->cost = whole; me->log2 = (int8_t)log2; me->nxn = 0; me->mode[0] = (uint8_t)mode;
      return whole;
    }
  }
  return split_cost;                                     /* split: the entry keeps the plan of the first sub-block */
}

static void intra_emit(orc_encoder_t *e, int x0, int y0, int log2, const intra_plan_t *plan)
{
  if (x0 >= e->w || y0 >= e->h) return;
  const intra_plan_t *me = &plan[((y0 & (CTB - 1)) >> 3) * 8 + ((x0 & (CTB - 1)) >> 3)];
  if (me->log2 == log2) {
    int modes[4] = {me->mode[0], me->mode[1], me->mode[2], me->mode[3]};
    intra_cu_recon_x(e, x0, y0, log2, modes, me->nxn);
    return;
  }
  const int hn = 1 << (log2 - 1);
  for (int q = 0; q < 4; q++) intra_emit(e, x0 + (q & 1) * hn, y0 + (q >> 1) * hn, log2 - 1, plan);
}

static void intra_quadtree(orc_encoder_t *e, int x0, int y0, int log2)
{
  intra_plan_t plan[64];
  memset(plan, 0, sizeof(plan));
  intra_plan(e, x0, y0, log2, plan);
  intra_emit(e, x0, y0, log2, plan);
}

/* ------------------------------------------------------------------------------------------ */
/* inter                                                                                          */

/* HM TComRdCost::xGetComponentBits: exp-Golomb-like length of one mvd component */
static int mv_comp_bits(int v)
{
  int len = 1;
  unsigned t = v <= 0 ? ((unsigned)(-v) << 1) + 1 : (unsigned)v << 1;
  while (t != 1) { t >>= 1; len += 2; }
  return len;
}
static inline uint32_t mv_penalty(int lambda_q4, int mvx, int mvy)
{
  return (uint32_t)((lambda_q4 * (mv_comp_bits(mvx) + mv_comp_bits(mvy))) >> 4);
}

/* Tile-column mode (cfg.mv_edges): may a block at x, n wide, use horizontal motion mvx?  With a
 * fractional luma or chroma position ((mvx & 7) != 0) the interpolation taps reach up to 4 luma
 * samples further on either side. */
static inline int mv_allowed(const orc_encoder_t *e, int x, int n, int mvx)
{
  if (!e->cfg.mv_edges) return 1;
  const int ix = mvx >> 2, m = (mvx & 7) ? 4 : 0;
  if ((e->cfg.mv_edges & 1) && x + ix - m < 0) return 0;
  if ((e->cfg.mv_edges & 2) && x + n + ix + m > e->w) return 0;
  return 1;
}
/* ... and vertical motion mvy for a block at y, n high (tile rows: mv_edges bit 2 / 3 = top / bottom edge) */
static inline int mv_allowed_v(const orc_encoder_t *e, int y, int n, int mvy)
{
  if (!e->cfg.mv_edges) return 1;
  const int iy = mvy >> 2, m = (mvy & 7) ? 4 : 0;
  if ((e->cfg.mv_edges & 4) && y + iy - m < 0) return 0;
  if ((e->cfg.mv_edges & 8) && y + n + iy + m > e->h) return 0;
  return 1;
}

/* The finished picture enters the decoded picture buffer as reference index 0; the oldest one leaves. */
static void down4(const uint8_t *p, int w, int h, uint8_t *out);
static void dpb_insert(orc_encoder_t *e)
{
  struct orc_ref last = e->dpb[e->cfg.refs - 1];
  for (int r = e->cfg.refs - 1; r > 0; r--) e->dpb[r] = e->dpb[r - 1];
  e->dpb[0] = last;
  struct orc_ref *d = &e->dpb[0];
  for (int c = 0; c < 3; c++) {
    int pw = c ? e->cw : e->w, ph = c ? e->ch : e->h, pad = c ? PAD / 2 : PAD, st = e->refstride[c];
    const uint8_t *s = plane(e->rec, e->w, e->h, c);
    for (int y = -pad; y < ph + pad; y++) {
      const uint8_t *row = s + (size_t)clip3i(0, ph - 1, y) * pw;
      uint8_t *o = d->pad[c] + (size_t)(y + pad) * st;
      memset(o, row[0], pad);
      memcpy(o + pad, row, pw);
      memset(o + pad + pw, row[pw - 1], pad);
    }
  }
  if (e->cfg.me_coarse > 0) down4(e->rec, e->w, e->h, d->q);
  d->poc = e->poc;
  if (e->cfg.tmvp)
    for (size_t i = 0; i < (size_t)e->w8 * e->h8; i++) {
      const orc_cu_t *cu = &e->cu[i];
      struct orc_mvf *m = &d->mvf[i];
      m->inter = cu->pred_mode == 0 && cu->log2_size >= 3;
      m->mvx = cu->mvx; m->mvy = cu->mvy;
      m->ref_poc = (int16_t)(e->poc - 1 - cu->ref_idx);            /* list 0 of a picture: POC - 1, POC - 2, ... */
    }
  if (e->n_dpb < e->cfg.refs) e->n_dpb++;
}

static void mc_luma(const orc_encoder_t *e, int ref, int x0, int y0, int n, int mvx, int mvy, uint8_t *dst)
{
  orc_mc_luma(e->dpb[ref].pad[0], e->refstride[0], e->w + 2 * PAD, e->h + 2 * PAD, x0 + PAD, y0 + PAD, n, n, mvx, mvy, dst, n);
}
static void mc_chroma(const orc_encoder_t *e, int ref, int c, int x0, int y0, int n, int mvx, int mvy, uint8_t *dst)
{
  orc_mc_chroma(e->dpb[ref].pad[c], e->refstride[c], e->cw + PAD, e->ch + PAD, x0 + PAD / 2, y0 + PAD / 2, n, n, mvx, mvy, dst, n);
}

typedef struct { uint32_t cost; int dx, dy, cx, cy, ref; } me_best_t;     /* vector, the centre it was found around (full samples), reference index */

/* quarter-resolution picture: every sample the rounded mean of a 4x4 block of the luma plane */
static void down4(const uint8_t *p, int w, int h, uint8_t *out)
{
  const int wq = w / 4, hq = h / 4;
  for (int y = 0; y < hq; y++)
    for (int x = 0; x < wq; x++) {
      int s = 8;
      for (int j = 0; j < 4; j++)
        for (int i = 0; i < 4; i++) s += p[(size_t)(4 * y + j) * w + 4 * x + i];
      out[(size_t)y * wq + x] = (uint8_t)(s >> 4);
    }
}

/* Coarse level of the motion search: best displacement of the 32x32 block at (qx, qy) on the
 * quarter-resolution pictures (an 8x8 block of coarse samples, cut at the picture edge), within
 * +-me_coarse coarse samples, reference coordinates clamped to the picture.  Cost = 16 * SAD + lambda *
 * bits of the vector; raster order, the first strictly smaller cost wins; kept only when below 3/4 of
 * the cost of the zero displacement.  out = full samples. */
static void coarse_search(const orc_encoder_t *e, int ref, int qx, int qy, int lam, int out[2])
{
  out[0] = out[1] = 0;
  if (qx >= e->w || qy >= e->h) return;
  const int wq = e->w / 4, hq = e->h / 4, Rc = e->cfg.me_coarse;
  const int x0 = qx / 4, y0 = qy / 4, bw = imin(8, wq - x0), bh = imin(8, hq - y0);
  uint32_t best = UINT_MAX, zero = UINT_MAX;
  for (int dy = -Rc; dy <= Rc; dy++)
    for (int dx = -Rc; dx <= Rc; dx++) {
      if (!mv_allowed(e, qx, imin(32, e->w - qx), dx * 16) || !mv_allowed_v(e, qy, imin(32, e->h - qy), dy * 16)) continue;
      uint32_t sad = 0;
      for (int y = 0; y < bh; y++)
        for (int x = 0; x < bw; x++)
          sad += (uint32_t)abs((int)e->src_q[(size_t)(y0 + y) * wq + x0 + x] -
                               (int)e->dpb[ref].q[(size_t)clip3i(0, hq - 1, y0 + y + dy) * wq + clip3i(0, wq - 1, x0 + x + dx)]);
      const uint32_t cost = 16 * sad + mv_penalty(lam, dx * 16, dy * 16);
      if (dx == 0 && dy == 0) zero = cost;
      if (cost < best) { best = cost; out[0] = 4 * dx; out[1] = 4 * dy; }
    }
  /* a second centre only where it clearly beats staying put: smooth content matches equally well at
   * many coarse displacements (aperture problem) and would otherwise buy a second window for nothing */
  if (best >= zero - (zero >> 2)) out[0] = out[1] = 0;
  /* ... and only where the zero-centred window does not cover it anyway (the vector behind a coarse
   * displacement lies within half a coarse step, 2 samples, of it) */
  if (imax(abs(out[0]), abs(out[1])) + 2 <= e->cfg.search_range) out[0] = out[1] = 0;
}

/* bins of ref_idx_l0 = r (TR, cMax = n - 1) */
static int ref_idx_bits(int r, int n) { return n <= 1 ? 0 : (r < n - 1 ? r + 1 : r); }

/* the centre the mv penalty of a CU counts from (quarter samples), kept for the fractional refinement */
static void set_pen_centre(orc_encoder_t *e, int x0, int y0, const me_best_t *b)
{
  int16_t *p = e->pen_ctr + 2 * ((size_t)(y0 / 8) * e->w8 + x0 / 8);
  p[0] = (int16_t)(4 * b->cx); p[1] = (int16_t)(4 * b->cy);
}

/* full-sample exhaustive search of one CTU at 8x8 granularity, partition decision, cu-map fill */
static void me_ctu(orc_encoder_t *e, int cx, int cy)
{
  const int R = e->cfg.search_range;
  const int lam = lambda_at(e, cx, cy);
  const uint8_t *src = e->src;
  const int rs = e->refstride[0];
  me_best_t b8[8][8], b16[4][4], b32[2][2];
  for (int j = 0; j < 8; j++) for (int i = 0; i < 8; i++) b8[j][i].cost = UINT_MAX;
  for (int j = 0; j < 4; j++) for (int i = 0; i < 4; i++) b16[j][i].cost = UINT_MAX;
  for (int j = 0; j < 2; j++) for (int i = 0; i < 2; i++) b32[j][i].cost = UINT_MAX;
  /* Search centres.  Set 0: the zero vector.  Set 1 (cfg.me_coarse): per 32x32 quadrant, the vector
   * the coarse level found (quarter-resolution pictures, +-me_coarse coarse samples); skipped where it
   * is the zero vector again.  Around every centre the same +-R full-sample window is searched; the
   * mv penalty counts from the centre (the predictor of the real coder is expected near it).
   * Candidates are ordered set 0 (raster), then set 1 (raster); the first strictly smaller cost wins. */
  /* With several reference pictures (cfg.refs) everything is repeated per reference; a candidate of
   * reference index r pays lambda * bits(ref_idx = r) on top.  Order: reference 0 first. */
  for (int r = 0; r < e->n_refs; r++) {
  const uint8_t *ref = e->dpb[r].pad[0];
  const uint32_t refpen = (uint32_t)((lam * ref_idx_bits(r, e->n_refs)) >> 4);
  int ctr[2][4][2], nsets = 1;
  memset(ctr, 0, sizeof(ctr));
  if (e->cfg.me_coarse > 0) {
    nsets = 2;
    for (int q = 0; q < 4; q++) coarse_search(e, r, cx + 32 * (q & 1), cy + 32 * (q >> 1), lam, ctr[1][q]);
  }
  for (int set = 0; set < nsets; set++)
    for (int q = 0; q < 4; q++) {
      const int qx = cx + 32 * (q & 1), qy = cy + 32 * (q >> 1);
      if (qx >= e->w || qy >= e->h) continue;
      const int mx0 = ctr[set][q][0], my0 = ctr[set][q][1];
      if (set == 1 && mx0 == 0 && my0 == 0) continue;
      for (int dy = -R; dy <= R; dy++)
        for (int dx = -R; dx <= R; dx++) {
          const uint32_t pen = mv_penalty(lam, dx * 4, dy * 4) + refpen;
          const int mx = mx0 + dx, my = my0 + dy;                    /* full-sample vector of this candidate */
          uint32_t s8[4][4], s16[2][2];
          for (int j = 0; j < 4; j++)
            for (int i = 0; i < 4; i++) {
              const int x = qx + 8 * i, y = qy + 8 * j, jj = 4 * (q >> 1) + j, ii = 4 * (q & 1) + i;
              if (x >= e->w || y >= e->h) { s8[j][i] = 0; continue; }
              s8[j][i] = orc_sad(src + (size_t)y * e->w + x, e->w, ref + (size_t)(y + PAD + my) * rs + x + PAD + mx, rs, 8, 8);
              uint32_t cost = s8[j][i] + pen;
              if (cost < b8[jj][ii].cost && mv_allowed(e, x, 8, mx * 4) && mv_allowed_v(e, y, 8, my * 4)) { b8[jj][ii].cost = cost; b8[jj][ii].dx = mx; b8[jj][ii].dy = my; b8[jj][ii].cx = mx0; b8[jj][ii].cy = my0; b8[jj][ii].ref = r; }
            }
          for (int j = 0; j < 2; j++)
            for (int i = 0; i < 2; i++) {
              const int x = qx + 16 * i, y = qy + 16 * j, jj = 2 * (q >> 1) + j, ii = 2 * (q & 1) + i;
              s16[j][i] = s8[2 * j][2 * i] + s8[2 * j][2 * i + 1] + s8[2 * j + 1][2 * i] + s8[2 * j + 1][2 * i + 1];
              if (x + 16 > e->w || y + 16 > e->h) continue;
              uint32_t cost = s16[j][i] + pen;
              if (cost < b16[jj][ii].cost && mv_allowed(e, x, 16, mx * 4) && mv_allowed_v(e, y, 16, my * 4)) { b16[jj][ii].cost = cost; b16[jj][ii].dx = mx; b16[jj][ii].dy = my; b16[jj][ii].cx = mx0; b16[jj][ii].cy = my0; b16[jj][ii].ref = r; }
            }
          if (qx + 32 <= e->w && qy + 32 <= e->h) {
            uint32_t cost = s16[0][0] + s16[0][1] + s16[1][0] + s16[1][1] + pen;
            me_best_t *b = &b32[q >> 1][q & 1];
            if (cost < b->cost && mv_allowed(e, qx, 32, mx * 4) && mv_allowed_v(e, qy, 32, my * 4)) { b->cost = cost; b->dx = mx; b->dy = my; b->cx = mx0; b->cy = my0; b->ref = r; }
          }
        }
    }
  }
  /* bottom-up partition decision */
  const uint32_t ovh = (uint32_t)((lam * CU_OVERHEAD_BITS) >> 4);
  uint32_t eff16[4][4];
  uint8_t use16[4][4];
  int intra16[4][4];
  for (int j = 0; j < 4; j++)
    for (int i = 0; i < 4; i++) {
      uint32_t sum8 = 0;
      for (int q = 0; q < 4; q++) {
        int jj = 2 * j + (q >> 1), ii = 2 * i + (q & 1);
        if (b8[jj][ii].cost != UINT_MAX) sum8 += b8[jj][ii].cost + ovh;
      }
      use16[j][i] = b16[j][i].cost != UINT_MAX && b16[j][i].cost + ovh <= sum8;
      eff16[j][i] = use16[j][i] ? b16[j][i].cost + ovh : sum8;
      intra16[j][i] = -1;
      /* intra candidate: 16x16 blocks that lie wholly inside the picture and predict badly */
      if (e->cfg.intra_in_p && b16[j][i].cost != UINT_MAX && eff16[j][i] > INTRA_TRY_COST) {
        int mode;
        uint32_t ic = intra_mode_search(e, cx + 16 * i, cy + 16 * j, 4, &mode);
        ic += ic / 2 + (uint32_t)((lam * INTRA_OVERHEAD_BITS) >> 4) + ovh;
        if (ic < eff16[j][i]) { eff16[j][i] = ic; intra16[j][i] = mode; }
      }
    }
  for (int j = 0; j < 2; j++)
    for (int i = 0; i < 2; i++) {
      uint32_t sum16 = 0;
      for (int q = 0; q < 4; q++) sum16 += eff16[2 * j + (q >> 1)][2 * i + (q & 1)];
      int use32 = b32[j][i].cost != UINT_MAX && b32[j][i].cost + ovh <= sum16;
      orc_cu_t cu;
      memset(&cu, 0, sizeof(cu));
      cu.merge_idx = 0xff;
      if (use32) {
        cu.log2_size = 5; cu.mvx = (int16_t)(b32[j][i].dx * 4); cu.mvy = (int16_t)(b32[j][i].dy * 4); cu.ref_idx = (uint8_t)b32[j][i].ref;
        set_cu(e, cx + 32 * i, cy + 32 * j, 5, &cu);
        set_pen_centre(e, cx + 32 * i, cy + 32 * j, &b32[j][i]);
        continue;
      }
      for (int q = 0; q < 4; q++) {
        int jj = 2 * j + (q >> 1), ii = 2 * i + (q & 1);
        if (intra16[jj][ii] >= 0) {                     /* reconstructed after all inter CUs, in coding order */
          orc_cu_t ic;
          memset(&ic, 0, sizeof(ic));
          ic.merge_idx = 0xff; ic.log2_size = 4; ic.pred_mode = 1; ic.intra_mode = (uint8_t)intra16[jj][ii];
          set_cu(e, cx + 16 * ii, cy + 16 * jj, 4, &ic);
          continue;
        }
        if (use16[jj][ii]) {
          cu.log2_size = 4; cu.mvx = (int16_t)(b16[jj][ii].dx * 4); cu.mvy = (int16_t)(b16[jj][ii].dy * 4); cu.ref_idx = (uint8_t)b16[jj][ii].ref;
          set_cu(e, cx + 16 * ii, cy + 16 * jj, 4, &cu);
          set_pen_centre(e, cx + 16 * ii, cy + 16 * jj, &b16[jj][ii]);
          continue;
        }
        for (int r = 0; r < 4; r++) {
          int j8 = 2 * jj + (r >> 1), i8 = 2 * ii + (r & 1);
          if (b8[j8][i8].cost == UINT_MAX) continue;
          cu.log2_size = 3; cu.mvx = (int16_t)(b8[j8][i8].dx * 4); cu.mvy = (int16_t)(b8[j8][i8].dy * 4); cu.ref_idx = (uint8_t)b8[j8][i8].ref;
          set_cu(e, cx + 8 * i8, cy + 8 * j8, 3, &cu);
          set_pen_centre(e, cx + 8 * i8, cy + 8 * j8, &b8[j8][i8]);
        }
      }
    }
}

/* half- then quarter-sample refinement of one CU, then prediction + residual + reconstruction */
static void inter_cu(orc_encoder_t *e, int x0, int y0, int log2)
{
  static const int8_t off[8][2] = {{-1, -1}, {0, -1}, {1, -1}, {-1, 0}, {1, 0}, {-1, 1}, {0, 1}, {1, 1}};
  const int n = 1 << log2;
  orc_cu_t cu = e->cu[(size_t)(y0 / 8) * e->w8 + x0 / 8];
  const uint8_t *src = e->src + (size_t)y0 * e->w + x0;
  uint8_t pred[32 * 32], best_pred[32 * 32];
  int bx = cu.mvx, by = cu.mvy;
  const int rf = cu.ref_idx;
  mc_luma(e, rf, x0, y0, n, bx, by, best_pred);
  const int lam = lambda_at(e, x0, y0);
  const int satd = e->cfg.subme_satd;
  const int16_t *pc = e->pen_ctr + 2 * ((size_t)(y0 / 8) * e->w8 + x0 / 8);     /* the mv penalty counts from the search centre */
  uint32_t best = (satd ? orc_satd(src, e->w, best_pred, n, n, n) : orc_sad(src, e->w, best_pred, n, n, n)) + mv_penalty(lam, bx - pc[0], by - pc[1]);
  for (int step = 2; step >= 1; step--) {
    int cxm = bx, cym = by;
    for (int k = 0; k < 8; k++) {
      int mx = cxm + off[k][0] * step, my = cym + off[k][1] * step;
      if (!mv_allowed(e, x0, n, mx) || !mv_allowed_v(e, y0, n, my)) continue;
      mc_luma(e, rf, x0, y0, n, mx, my, pred);
      uint32_t cost = (satd ? orc_satd(src, e->w, pred, n, n, n) : orc_sad(src, e->w, pred, n, n, n)) + mv_penalty(lam, mx - pc[0], my - pc[1]);
      if (cost < best) { best = cost; bx = mx; by = my; memcpy(best_pred, pred, (size_t)n * n); }
    }
  }
  cu.mvx = (int16_t)bx; cu.mvy = (int16_t)by; cu.tu_log2 = (uint8_t)log2;
  set_cu(e, x0, y0, log2, &cu);
  uint8_t pred_c[2][16 * 16];
  for (int c = 1; c < 3; c++) mc_chroma(e, rf, c, x0 / 2, y0 / 2, n / 2, bx, by, pred_c[c - 1]);
  cu_pred_t ip = {{best_pred, pred_c[0], pred_c[1]}, {n, n / 2, n / 2}, x0, y0, {0, 0, 0, 0}, 0};
  tu_node(e, x0, y0, log2, 0, &ip);
  cu_root_cbf(e, x0, y0, log2);
}

/* P pictures have no intra-picture dependency before the entropy stage, so both loops are
 * plain parallel-for over CTUs / CUs (threads set by orc_set_threads; 1 by default). */
static void inter_frame(orc_encoder_t *e)
{
  const int nctb = e->ctb_cols * e->ctb_rows;
  if (e->cfg.me_coarse > 0) down4(e->src, e->w, e->h, e->src_q);
#pragma omp parallel for schedule(dynamic, 1)
  for (int i = 0; i < nctb; i++) me_ctu(e, (i % e->ctb_cols) * CTB, (i / e->ctb_cols) * CTB);
#pragma omp parallel for schedule(dynamic, 4)
  for (int y8 = 0; y8 < e->h8; y8++)
    for (int x8 = 0; x8 < e->w8; x8++) {
      const orc_cu_t *cu = &e->cu[(size_t)y8 * e->w8 + x8];
      int n8 = 1 << (cu->log2_size - 3);
      if (cu->pred_mode == 0 && (x8 & (n8 - 1)) == 0 && (y8 & (n8 - 1)) == 0) inter_cu(e, x8 * 8, y8 * 8, cu->log2_size);
    }
  if (!e->cfg.intra_in_p) return;
  /* intra CUs: they predict from reconstructed neighbours (inter CUs included, constrained_intra_pred
   * is off), so they follow all inter CUs and go in coding order among themselves */
  for (int ctu = 0; ctu < nctb; ctu++)
    for (int z = 0; z < 64; z++) {
      int x8 = (ctu % e->ctb_cols) * 8, y8 = (ctu / e->ctb_cols) * 8;
      for (int b = 0; b < 3; b++) { x8 += ((z >> (2 * b)) & 1) << b; y8 += ((z >> (2 * b + 1)) & 1) << b; }
      if (x8 >= e->w8 || y8 >= e->h8) continue;
      const orc_cu_t *cu = &e->cu[(size_t)y8 * e->w8 + x8];
      int n8 = 1 << (cu->log2_size - 3);
      if (cu->pred_mode == 1 && (x8 & (n8 - 1)) == 0 && (y8 & (n8 - 1)) == 0) intra_cu_recon(e, x8 * 8, y8 * 8, cu->log2_size, cu->intra_mode);
    }
}

/* ------------------------------------------------------------------------------------------ */
/* luma QP derivation with one quantisation group per CTU (8.6.1, diff_cu_qp_delta_depth = 0).
 * The left and above quantisation groups lie in other CTBs, so qPY_PRED = qPY_PREV: the slice QP
 * at the start of every CTU row (WPP), else the QP of the last CU of the previous CTU.  Inside a
 * CTU, CUs before the first one with a coded residual keep the predicted QP (CuQpDeltaVal is still
 * 0 there); that CU codes the delta and it and all later CUs have the CTU's target QP. */
static void derive_cu_qps(orc_encoder_t *e)
{
  for (int r = 0; r < e->ctb_rows; r++) {
    int pred = e->cfg.qp;
    for (int cidx = 0; cidx < e->ctb_cols; cidx++) {
      const int ctu = r * e->ctb_cols + cidx, target = ctu_qp(e, cidx * CTB, r * CTB);
      int first = 64;
      for (int z = 0; z < 64 && first == 64; z++) {
        int x8 = cidx * 8, y8 = r * 8;
        for (int b = 0; b < 3; b++) { x8 += ((z >> (2 * b)) & 1) << b; y8 += ((z >> (2 * b + 1)) & 1) << b; }
        if (x8 < e->w8 && y8 < e->h8 && (e->cu[(size_t)y8 * e->w8 + x8].flags & 2)) first = z;
      }
      for (int z = 0; z < 64; z++) {
        int x8 = cidx * 8, y8 = r * 8;
        for (int b = 0; b < 3; b++) { x8 += ((z >> (2 * b)) & 1) << b; y8 += ((z >> (2 * b + 1)) & 1) << b; }
        if (x8 < e->w8 && y8 < e->h8) e->cu[(size_t)y8 * e->w8 + x8].qp = (uint8_t)(z < first ? pred : target);
      }
      e->ctu_first[ctu] = (uint8_t)first;
      e->ctu_delta[ctu] = (int8_t)(first < 64 ? ((target - pred + 26 + 52) % 52) - 26 : 0);
      if (first < 64) pred = target;
    }
  }
}

/* ------------------------------------------------------------------------------------------ */
/* deblocking (8.7.2): all vertical edges of the picture, then all horizontal edges             */

/* luma coded block flag of the transform block of unit u that touches edge segment `seg` (0 / 1: first /
 * second four samples along the edge); side 0: u lies before the edge (P), 1: after it (Q) */
static int edge_cbf(const orc_cu_t *u, int dir, int side, int seg)
{
  if (u->tu_log2 != 2) return u->cbf & 1;
  const int b = dir == 0 ? (side ? 0 : 1) + 2 * seg : (side ? 0 : 2) + seg;
  return (u->cbf >> (4 + b)) & 1;
}

static int edge_bs(const orc_cu_t *p, const orc_cu_t *q, int dir, int seg)
{
  if (p->pred_mode == 1 || q->pred_mode == 1) return 2;
  if (edge_cbf(p, dir, 0, seg) || edge_cbf(q, dir, 1, seg)) return 1;      /* either transform block has coefficients */
  if (p->ref_idx != q->ref_idx) return 1;              /* different reference pictures (list 0 has no duplicates here) */
  return abs(p->mvx - q->mvx) >= 4 || abs(p->mvy - q->mvy) >= 4;
}

static void deblock_frame(orc_encoder_t *e)
{
  uint8_t *Y = e->rec, *U = plane(e->rec, e->w, e->h, 1), *V = plane(e->rec, e->w, e->h, 2);
  for (int dir = 0; dir < 2; dir++)                    /* 0: vertical edges, 1: horizontal edges */
#pragma omp parallel for schedule(static)
    for (int y8 = 0; y8 < e->h8; y8++)
      for (int x8 = 0; x8 < e->w8; x8++) {
        const orc_cu_t *q = &e->cu[(size_t)y8 * e->w8 + x8];
        int n8 = q->tu_log2 > 3 ? 1 << (q->tu_log2 - 3) : 1;     /* transform unit edges (they include the CU edges) */
        if (dir == 0 ? (x8 == 0 || (x8 & (n8 - 1))) : (y8 == 0 || (y8 & (n8 - 1)))) continue;
        const orc_cu_t *p = dir == 0 ? q - 1 : q - e->w8;
        int x = x8 * 8, y = y8 * 8, bs = 0;
        const int qp = e->cfg.qp_delta ? (p->qp + q->qp + 1) >> 1 : e->cfg.qp;     /* QpL (8.7.2.5.3) */
        const int bo = e->cfg.beta_offset_div2, to = e->cfg.tc_offset_div2;
        for (int seg = 0; seg < 2; seg++) {
          bs = edge_bs(p, q, dir, seg);
          if (!bs) continue;
          if (dir == 0) orc_deblock_luma_segment2(Y + (size_t)(y + 4 * seg) * e->w + x, 1, e->w, bs, qp, bo, to);
          else          orc_deblock_luma_segment2(Y + (size_t)y * e->w + x + 4 * seg, e->w, 1, bs, qp, bo, to);
        }
        if (bs == 2 && (dir == 0 ? (x & 15) == 0 : (y & 15) == 0)) {
          int xc = x / 2, yc = y / 2;
          if (dir == 0) {
            orc_deblock_chroma_segment2(U + (size_t)yc * e->cw + xc, 1, e->cw, qp, 4, e->cfg.cb_qp_offset, to);
            orc_deblock_chroma_segment2(V + (size_t)yc * e->cw + xc, 1, e->cw, qp, 4, e->cfg.cr_qp_offset, to);
          } else {
            orc_deblock_chroma_segment2(U + (size_t)yc * e->cw + xc, e->cw, 1, qp, 4, e->cfg.cb_qp_offset, to);
            orc_deblock_chroma_segment2(V + (size_t)yc * e->cw + xc, e->cw, 1, qp, 4, e->cfg.cr_qp_offset, to);
          }
        }
      }
}


/* ------------------------------------------------------------------------------------------ */
/* sample adaptive offset (8.7.3): decision per CTU on source vs deblocked samples, then applied   */
/* from the deblocked copy into rec.  Never merged with a neighbour (the flags are coded as 0).    */

static inline int sgn(int v) { return (v > 0) - (v < 0); }

static const int8_t sao_dir[4][2][2] = {          /* class: {a, b} as (dx, dy) */
  {{-1, 0}, {1, 0}}, {{0, -1}, {0, 1}}, {{-1, -1}, {1, 1}}, {{1, -1}, {-1, 1}}};

/* category 1..4 of a sample for an edge class, 0 = none (also when a neighbour is outside the picture) */
static int sao_category(const uint8_t *p, int stride, int x, int y, int pw, int ph, int cls)
{
  const int ax = x + sao_dir[cls][0][0], ay = y + sao_dir[cls][0][1];
  const int bx = x + sao_dir[cls][1][0], by = y + sao_dir[cls][1][1];
  if (ax < 0 || ay < 0 || bx < 0 || by < 0 || ax >= pw || ay >= ph || bx >= pw || by >= ph) return 0;
  const int c = p[(size_t)y * stride + x];
  const int e = 2 + sgn(c - p[(size_t)ay * stride + ax]) + sgn(c - p[(size_t)by * stride + bx]);
  return e == 2 ? 0 : (e < 2 ? e + 1 : e);          /* 0 -> 1 (local minimum), 1 -> 2, 3 -> 3, 4 -> 4 (local maximum) */
}

typedef struct { long long gain_cost; int type, cls, band; int8_t off[4]; } sao_choice_t;

/* best offset of a category given count and sum of (source - deblocked); returns the change in SSE */
static long long sao_best_offset(long long count, long long sum, int lo, int hi, int *off)
{
  if (count == 0) { *off = 0; return 0; }
  int o = (int)((sum >= 0 ? sum + count / 2 : sum - count / 2) / count);
  o = clip3i(lo, hi, o);
  long long best = 0;
  int best_o = 0;
  for (int k = o; k != 0; k += (k > 0 ? -1 : 1)) {          /* smaller magnitudes cost fewer bits: check them all */
    long long d = count * k * k - 2 * k * sum;
    if (d < best) { best = d; best_o = k; }
  }
  *off = best_o;
  return best;
}

/* statistics and decision of one component group of one CTU: comps = {0} or {1, 2} */
static void sao_decide_group(orc_encoder_t *e, int cx, int cy, int group, struct orc_sao *out)
{
  const int lq = lambda_at(e, cx, cy);
  const int lam = (lq * lq + 128) >> 8;                                /* lambda, SSE domain */
  const int c_first = group ? 1 : 0, c_last = group ? 2 : 0;
  long long best_cost = 0;                                             /* "off": no change, one bin */
  int best_type = 0, best_cls = 0, best_band[3] = {0, 0, 0};
  int8_t best_off[3][4];
  memset(best_off, 0, sizeof(best_off));
  /* edge offset, classes 0..3 */
  for (int cls = 0; cls < 4; cls++) {
    long long cost = (long long)lam * 4;                               /* type + class */
    int8_t off[3][4];
    for (int c = c_first; c <= c_last; c++) {
      const int pw = c ? e->cw : e->w, ph = c ? e->ch : e->h, sh = c ? 1 : 0;
      const uint8_t *src = plane((uint8_t *)e->src, e->w, e->h, c), *dbk = plane(e->dbk, e->w, e->h, c);
      const int x0 = cx >> sh, y0 = cy >> sh, x1 = imin(pw, (cx + CTB) >> sh), y1 = imin(ph, (cy + CTB) >> sh);
      long long cnt[5] = {0}, sum[5] = {0};
      for (int y = y0; y < y1; y++)
        for (int x = x0; x < x1; x++) {
          int k = sao_category(dbk, pw, x, y, pw, ph, cls);
          cnt[k]++; sum[k] += (int)src[(size_t)y * pw + x] - (int)dbk[(size_t)y * pw + x];
        }
      for (int k = 1; k <= 4; k++) {
        int o;
        cost += sao_best_offset(cnt[k], sum[k], k <= 2 ? 0 : -7, k <= 2 ? 7 : 0, &o);
        cost += (long long)lam * (abs(o) + 1);
        off[c][k - 1] = (int8_t)o;
      }
    }
    if (cost < best_cost) {
      best_cost = cost; best_type = 2; best_cls = cls;
      for (int c = c_first; c <= c_last; c++) memcpy(best_off[c], off[c], 4);
    }
  }
  /* band offset: four consecutive bands of 8 levels */
  {
    long long cost = (long long)lam * 2;
    int8_t off[3][4];
    int band[3] = {0, 0, 0};
    for (int c = c_first; c <= c_last; c++) {
      const int pw = c ? e->cw : e->w, ph = c ? e->ch : e->h, sh = c ? 1 : 0;
      const uint8_t *src = plane((uint8_t *)e->src, e->w, e->h, c), *dbk = plane(e->dbk, e->w, e->h, c);
      const int x0 = cx >> sh, y0 = cy >> sh, x1 = imin(pw, (cx + CTB) >> sh), y1 = imin(ph, (cy + CTB) >> sh);
      long long cnt[32] = {0}, sum[32] = {0}, gain[32];
      int bo[32];
      for (int y = y0; y < y1; y++)
        for (int x = x0; x < x1; x++) {
          int k = dbk[(size_t)y * pw + x] >> 3;
          cnt[k]++; sum[k] += (int)src[(size_t)y * pw + x] - (int)dbk[(size_t)y * pw + x];
        }
      for (int k = 0; k < 32; k++) {
        gain[k] = sao_best_offset(cnt[k], sum[k], -7, 7, &bo[k]);
        gain[k] += (long long)lam * (abs(bo[k]) + 1 + (bo[k] != 0));
      }
      long long bg = 0;
      int bs = 0;
      for (int st = 0; st <= 28; st++) {
        long long g = gain[st] + gain[st + 1] + gain[st + 2] + gain[st + 3];
        if (st == 0 || g < bg) { bg = g; bs = st; }
      }
      cost += bg + (long long)lam * 5;
      band[c] = bs;
      for (int k = 0; k < 4; k++) off[c][k] = (int8_t)bo[bs + k];
    }
    if (cost < best_cost) {
      best_cost = cost; best_type = 1;
      for (int c = c_first; c <= c_last; c++) { memcpy(best_off[c], off[c], 4); best_band[c] = band[c]; }
    }
  }
  out->type[group] = (uint8_t)best_type;
  out->eo_class[group] = (uint8_t)(best_type == 2 ? best_cls : 0);
  for (int c = c_first; c <= c_last; c++) {
    out->band_pos[c] = (uint8_t)best_band[c];
    memcpy(out->offset[c], best_off[c], 4);
  }
}

static void sao_frame(orc_encoder_t *e)
{
  const size_t fsz = (size_t)e->w * e->h * 3 / 2;
  memcpy(e->dbk, e->rec, fsz);
  const int nctb = e->ctb_cols * e->ctb_rows;
#pragma omp parallel for schedule(dynamic, 1)
  for (int i = 0; i < nctb; i++) {
    const int cx = (i % e->ctb_cols) * CTB, cy = (i / e->ctb_cols) * CTB;
    struct orc_sao *p = &e->sao[i];
    memset(p, 0, sizeof(*p));
    sao_decide_group(e, cx, cy, 0, p);
    sao_decide_group(e, cx, cy, 1, p);
    for (int c = 0; c < 3; c++) {
      const int g = c ? 1 : 0;
      if (!p->type[g]) continue;
      const int pw = c ? e->cw : e->w, ph = c ? e->ch : e->h, sh = c ? 1 : 0;
      const uint8_t *dbk = plane(e->dbk, e->w, e->h, c);
      uint8_t *rec = plane(e->rec, e->w, e->h, c);
      const int x0 = cx >> sh, y0 = cy >> sh, x1 = imin(pw, (cx + CTB) >> sh), y1 = imin(ph, (cy + CTB) >> sh);
      for (int y = y0; y < y1; y++)
        for (int x = x0; x < x1; x++) {
          const int v = dbk[(size_t)y * pw + x];
          int o = 0;
          if (p->type[g] == 2) {
            int k = sao_category(dbk, pw, x, y, pw, ph, p->eo_class[g]);
            if (k) o = p->offset[c][k - 1];
          } else {
            int k = (v >> 3) - p->band_pos[c];
            if (k >= 0 && k < 4) o = p->offset[c][k];
          }
          rec[(size_t)y * pw + x] = (uint8_t)clip3i(0, 255, v + o);
        }
    }
  }
}

/* sao() syntax of one CTU (7.3.8.3).  cfg.sao == 1: merge candidates are never used; cfg.sao == 2: a
 * CTU whose parameters equal those of the CTU to its left (else of the one above) says so with the
 * merge flag instead of repeating them (the parameters themselves are decided independently, so this
 * changes the rate only). */
static void code_sao(const orc_encoder_t *e, orc_cabac_t *c, int rx, int ry)
{
  const struct orc_sao *p = &e->sao[ry * e->ctb_cols + rx];
  const int try_merge = e->cfg.sao == 2;
  if (rx > 0) {
    const int m = try_merge && memcmp(p, p - 1, sizeof(*p)) == 0;
    orc_cabac_bin(c, CTX_SAO_MERGE, m);                           /* sao_merge_left_flag */
    if (m) return;
  }
  if (ry > 0) {
    const int m = try_merge && memcmp(p, p - e->ctb_cols, sizeof(*p)) == 0;
    orc_cabac_bin(c, CTX_SAO_MERGE, m);                           /* sao_merge_up_flag */
    if (m) return;
  }
  for (int comp = 0; comp < 3; comp++) {
    const int g = comp ? 1 : 0;
    if (comp < 2) {                                               /* sao_type_idx_luma / _chroma: TR cMax 2, first bin coded */
      orc_cabac_bin(c, CTX_SAO_TYPE, p->type[g] != 0);
      if (p->type[g]) orc_cabac_bypass(c, p->type[g] == 2);
    }
    if (!p->type[g]) continue;
    for (int k = 0; k < 4; k++) {                                 /* sao_offset_abs: TR cMax 7, bypass */
      const int a = abs(p->offset[comp][k]);
      for (int i = 0; i < a; i++) orc_cabac_bypass(c, 1);
      if (a < 7) orc_cabac_bypass(c, 0);
    }
    if (p->type[g] == 1) {
      for (int k = 0; k < 4; k++)
        if (p->offset[comp][k]) orc_cabac_bypass(c, p->offset[comp][k] < 0);     /* sao_offset_sign */
      orc_cabac_bypass_bits(c, p->band_pos[comp], 5);             /* sao_band_position */
    } else if (comp < 2) {
      orc_cabac_bypass_bits(c, p->eo_class[g], 2);                /* sao_eo_class_luma / _chroma */
    }
  }
}

/* ------------------------------------------------------------------------------------------ */
/* entropy coding                                                                                 */

typedef struct { int16_t x, y; int8_t ref; } mv_t;

/* 6.4.2: neighbouring prediction block available for inter candidates */
static const orc_cu_t *inter_nb(const orc_encoder_t *e, int xc, int yc, int xn, int yn)
{
  if (xn < 0 || yn < 0 || xn >= e->w || yn >= e->h) return NULL;
  if (coding_order(e, xn, yn) >= coding_order(e, xc, yc)) return NULL;
  const orc_cu_t *n = &e->cu[(size_t)(yn >> 3) * e->w8 + (xn >> 3)];
  return n->pred_mode == 0 ? n : NULL;
}
static inline int same_mv(const orc_cu_t *a, const orc_cu_t *b) { return a->mvx == b->mvx && a->mvy == b->mvy && a->ref_idx == b->ref_idx; }

/* 8.5.3.2.7 / 8.5.3.2.8: a vector that spans td pictures rescaled to span tb pictures */
static int scale_mv(int mv, int td, int tb)
{
  td = clip3i(-128, 127, td); tb = clip3i(-128, 127, tb);
  const int tx = (16384 + (abs(td) >> 1)) / td;
  const int dsf = clip3i(-4096, 4095, (tb * tx + 32) >> 6);
  const int p = dsf * mv;
  return clip3i(-32768, 32767, (p < 0 ? -1 : 1) * ((abs(p) + 127) >> 8));
}

/* 8.5.3.2.8 temporal luma motion vector prediction for a block whose target is reference index
 * `ref`: the collocated picture is reference index 0 (collocated_ref_idx = 0), bottom-right
 * candidate first (same CTB row, inside the picture), then the centre; motion is read from the
 * 16x16-compressed motion field. */
static int temporal_mv(const orc_encoder_t *e, int x0, int y0, int n, int ref, mv_t *out)
{
  if (!e->cfg.tmvp || e->n_refs < 1) return 0;
  const struct orc_ref *col = &e->dpb[0];
  const int cand[2][2] = {{x0 + n, y0 + n}, {x0 + n / 2, y0 + n / 2}};
  for (int k = 0; k < 2; k++) {
    const int x = cand[k][0], y = cand[k][1];
    if (k == 0 && ((y0 >> CTB_LOG2) != (y >> CTB_LOG2) || y >= e->h || x >= e->w)) continue;
    const struct orc_mvf *m = &col->mvf[(size_t)(((y >> 4) << 4) >> 3) * e->w8 + ((((x >> 4) << 4)) >> 3)];
    if (!m->inter) continue;
    const int col_diff = col->poc - m->ref_poc, cur_diff = 1 + ref;      /* POC distances: the collocated block's, ours */
    out->ref = (int8_t)ref;
    if (col_diff == cur_diff) { out->x = m->mvx; out->y = m->mvy; }
    else { out->x = (int16_t)scale_mv(m->mvx, col_diff, cur_diff); out->y = (int16_t)scale_mv(m->mvy, col_diff, cur_diff); }
    return 1;
  }
  return 0;
}

/* 8.5.3.2.2-8.5.3.2.5 merge candidate list of a P slice: spatial, temporal, zero candidates */
static int merge_candidates(const orc_encoder_t *e, int x0, int y0, int n, mv_t out[MAX_MERGE])
{
  const orc_cu_t *a1 = inter_nb(e, x0, y0, x0 - 1, y0 + n - 1);
  const orc_cu_t *b1 = inter_nb(e, x0, y0, x0 + n - 1, y0 - 1);
  const orc_cu_t *b0 = inter_nb(e, x0, y0, x0 + n, y0 - 1);
  const orc_cu_t *a0 = inter_nb(e, x0, y0, x0 - 1, y0 + n);
  const orc_cu_t *b2 = inter_nb(e, x0, y0, x0 - 1, y0 - 1);
  int cnt = 0;
  if (b1 && a1 && same_mv(b1, a1)) b1 = NULL;
  const orc_cu_t *b1o = inter_nb(e, x0, y0, x0 + n - 1, y0 - 1);   /* comparisons use the unpruned B1 */
  if (b0 && b1o && same_mv(b0, b1o)) b0 = NULL;
  if (a0 && a1 && same_mv(a0, a1)) a0 = NULL;
  if (b2 && ((a1 && same_mv(b2, a1)) || (b1o && same_mv(b2, b1o)))) b2 = NULL;
#define PUSH(c) do { out[cnt].x = (c)->mvx; out[cnt].y = (c)->mvy; out[cnt++].ref = (int8_t)(c)->ref_idx; } while (0)
  if (a1) PUSH(a1);
  if (b1) PUSH(b1);
  if (b0) PUSH(b0);
  if (a0) PUSH(a0);
  if (b2 && cnt < 4) PUSH(b2);
#undef PUSH
  if (cnt < MAX_MERGE && temporal_mv(e, x0, y0, n, 0, &out[cnt])) cnt++;
  for (int zero_idx = 0; cnt < MAX_MERGE; zero_idx++) {              /* zero candidates walk the reference indices */
    out[cnt].x = 0; out[cnt].y = 0; out[cnt++].ref = (int8_t)(zero_idx < e->n_refs ? zero_idx : 0);
  }
  return cnt;
}

/* 8.5.3.2.6-8.5.3.2.8 AMVP list for reference index `ref`: spatial candidates A (A0, A1) and B (B0, B1,
 * B2) -- first the neighbours that point to the same reference picture, then any, rescaled by the
 * ratio of POC distances -- then the temporal candidate, then zero vectors */
static void amvp_candidates(const orc_encoder_t *e, int x0, int y0, int n, int ref, mv_t out[2])
{
  const orc_cu_t *ak[2] = {inter_nb(e, x0, y0, x0 - 1, y0 + n), inter_nb(e, x0, y0, x0 - 1, y0 + n - 1)};
  const orc_cu_t *bk[3] = {inter_nb(e, x0, y0, x0 + n, y0 - 1), inter_nb(e, x0, y0, x0 + n - 1, y0 - 1), inter_nb(e, x0, y0, x0 - 1, y0 - 1)};
  const int is_scaled = ak[0] || ak[1];
  int fa = 0, fb = 0;
  mv_t a = {0, 0, 0}, b = {0, 0, 0};
  for (int k = 0; k < 2 && !fa; k++)
    if (ak[k] && ak[k]->ref_idx == ref) { a.x = ak[k]->mvx; a.y = ak[k]->mvy; fa = 1; }
  for (int k = 0; k < 2 && !fa; k++)
    if (ak[k]) {
      a.x = (int16_t)scale_mv(ak[k]->mvx, 1 + ak[k]->ref_idx, 1 + ref); a.y = (int16_t)scale_mv(ak[k]->mvy, 1 + ak[k]->ref_idx, 1 + ref);
      fa = 1;
    }
  for (int k = 0; k < 3 && !fb; k++)
    if (bk[k] && bk[k]->ref_idx == ref) { b.x = bk[k]->mvx; b.y = bk[k]->mvy; fb = 1; }
  if (!is_scaled && fb) { a = b; fa = 1; }
  if (!is_scaled) {
    fb = 0;
    for (int k = 0; k < 3 && !fb; k++)
      if (bk[k]) {
        if (bk[k]->ref_idx == ref) { b.x = bk[k]->mvx; b.y = bk[k]->mvy; }
        else { b.x = (int16_t)scale_mv(bk[k]->mvx, 1 + bk[k]->ref_idx, 1 + ref); b.y = (int16_t)scale_mv(bk[k]->mvy, 1 + bk[k]->ref_idx, 1 + ref); }
        fb = 1;
      }
  }
  int cnt = 0;
  if (fa) out[cnt++] = a;
  if (fb && !(fa && a.x == b.x && a.y == b.y)) out[cnt++] = b;
  if (cnt < 2 && temporal_mv(e, x0, y0, n, ref, &out[cnt])) cnt++;
  while (cnt < 2) { out[cnt].x = 0; out[cnt++].y = 0; }
  out[0].ref = out[1].ref = (int8_t)ref;
}

static void code_mvd(orc_cabac_t *c, int dx, int dy)
{
  int ax = abs(dx), ay = abs(dy);
  orc_cabac_bin(c, CTX_MVD_GT0, ax > 0);
  orc_cabac_bin(c, CTX_MVD_GT0, ay > 0);
  if (ax) orc_cabac_bin(c, CTX_MVD_GT1, ax > 1);
  if (ay) orc_cabac_bin(c, CTX_MVD_GT1, ay > 1);
  for (int k = 0; k < 2; k++) {
    int a = k ? ay : ax, d = k ? dy : dx;
    if (!a) continue;
    if (a > 1) {                         /* abs_mvd_minus2: EG1 */
      int v = a - 2, kk = 1;
      while (v >= (1 << kk)) { orc_cabac_bypass(c, 1); v -= 1 << kk; kk++; }
      orc_cabac_bypass(c, 0);
      orc_cabac_bypass_bits(c, (uint32_t)v, kk);
    }
    orc_cabac_bypass(c, d < 0);
  }
}

static int scan_idx_for(int pred_mode, int intra_mode, int log2n, int cidx)
{
  if (pred_mode != 1) return 0;
  if (!(log2n == 2 || (log2n == 3 && cidx == 0))) return 0;
  if (intra_mode >= 6 && intra_mode <= 14) return 2;
  if (intra_mode >= 22 && intra_mode <= 30) return 1;
  return 0;
}

/* cu_qp_delta_abs / sign (7.3.8.10, 9.3.3.10): prefix TR cMax 5 (ctx 0, then ctx 1), suffix EG0 bypass */
static void code_cu_qp_delta(orc_cabac_t *c, int d)
{
  const int a = abs(d), pre = a < 5 ? a : 5;
  for (int i = 0; i < pre; i++) orc_cabac_bin(c, CTX_CU_QP_DELTA + (i ? 1 : 0), 1);
  if (pre < 5) orc_cabac_bin(c, CTX_CU_QP_DELTA + (pre ? 1 : 0), 0);
  else {
    int v = a - 5, k = 0;
    while (v >= (1 << k)) { orc_cabac_bypass(c, 1); v -= 1 << k; k++; }
    orc_cabac_bypass(c, 0);
    if (k) orc_cabac_bypass_bits(c, (uint32_t)v, k);
  }
  if (a) orc_cabac_bypass(c, d < 0);
}

/* luma intra prediction mode of the 4x4 block that covers luma sample (x, y) of an intra CU */
static int intra_mode_at(const orc_cu_t *u, int x, int y)
{
  if (!(u->flags & 1)) return u->intra_mode;
  const int part = (((y >> 2) & 1) << 1) | ((x >> 2) & 1);
  return part == 0 ? u->intra_mode : part == 1 ? (u->mvx & 0xff) : part == 2 ? ((u->mvx >> 8) & 0xff) : (u->mvy & 0xff);
}

static void code_tu_residuals(orc_encoder_t *e, orc_cabac_t *c, int x0, int y0, int log2, int cidx, const orc_cu_t *cu)
{
  const int sh = cidx ? 1 : 0, pw = cidx ? e->cw : e->w;
  const int mode = cidx ? cu->chroma_mode : intra_mode_at(cu, x0, y0);
  orc_code_residual2(c, lplane(e->levels, e->w, e->h, cidx) + (size_t)(y0 >> sh) * pw + (x0 >> sh), pw, log2, cidx,
                     scan_idx_for(cu->pred_mode, mode, log2, cidx), e->cfg.sign_hiding);
}

static void code_delta_qp_once(orc_encoder_t *e, orc_cabac_t *c, int x0, int y0)
{
  if (!e->cfg.qp_delta || e->delta_coded) return;
  /* once per quantisation group (= CTU), in its first transform unit with a coded block flag */
  code_cu_qp_delta(c, e->ctu_delta[(y0 / CTB) * e->ctb_cols + x0 / CTB]);
  e->delta_coded = 1;
}

/* transform_tree (7.3.8.8) of the node (x0, y0, log2) at trafoDepth `depth`; the tree itself is read
 * from the cu map (size of the transform unit covering each 8x8 unit).  par_cb / par_cr: the parent's
 * chroma flags (1 at depth 0). */
static void code_transform_tree(orc_encoder_t *e, orc_cabac_t *c, int x0, int y0, int log2, int depth, int par_cb, int par_cr,
                                const orc_cu_t *cu)
{
  const int n8 = 1 << (log2 - 3);
  const orc_cu_t *u0 = &e->cu[(size_t)(y0 >> 3) * e->w8 + (x0 >> 3)];
  const int split = u0->tu_log2 < log2;
  const int nxn = cu->pred_mode == 1 && (cu->flags & 1);             /* IntraSplitFlag */
  if (log2 <= 5 && log2 > 2 && depth < e->cfg.tr_depth + nxn && !(nxn && depth == 0))
    orc_cabac_bin(c, CTX_SPLIT_TRANSFORM + 5 - log2, split);
  int cb = 0, cr = 0;                                    /* of the node: set when any transform unit below has them */
  for (int j = 0; j < n8; j++)
    for (int i = 0; i < n8; i++) {
      const orc_cu_t *u = &e->cu[(size_t)((y0 >> 3) + j) * e->w8 + (x0 >> 3) + i];
      cb |= (u->cbf >> 1) & 1; cr |= (u->cbf >> 2) & 1;
    }
  if (par_cb) orc_cabac_bin(c, CTX_CBF_CHROMA + depth, cb);
  if (par_cr) orc_cabac_bin(c, CTX_CBF_CHROMA + depth, cr);
  if (split && log2 == 3) {
    /* four 4x4 luma transform units; the chroma blocks of the 8x8 node follow the fourth */
    for (int b = 0; b < 4; b++) {
      const int lu = (u0->cbf >> (4 + b)) & 1, xb = x0 + 4 * (b & 1), yb = y0 + 4 * (b >> 1);
      orc_cabac_bin(c, CTX_CBF_LUMA, lu);                /* trafoDepth > 0 */
      if (lu || cb || cr) code_delta_qp_once(e, c, x0, y0);
      if (lu) code_tu_residuals(e, c, xb, yb, 2, 0, cu);
    }
    if (cb) code_tu_residuals(e, c, x0, y0, 2, 1, cu);
    if (cr) code_tu_residuals(e, c, x0, y0, 2, 2, cu);
    return;
  }
  if (split) {
    const int h = 1 << (log2 - 1);
    for (int q = 0; q < 4; q++) code_transform_tree(e, c, x0 + (q & 1) * h, y0 + (q >> 1) * h, log2 - 1, depth + 1, cb, cr, cu);
    return;
  }
  const int lu = u0->cbf & 1;
  if (cu->pred_mode == 1 || depth != 0 || cb || cr) orc_cabac_bin(c, CTX_CBF_LUMA + (depth == 0 ? 1 : 0), lu);
  if (lu || cb || cr) code_delta_qp_once(e, c, x0, y0);
  if (lu) code_tu_residuals(e, c, x0, y0, log2, 0, cu);
  if (cb) code_tu_residuals(e, c, x0, y0, log2 - 1, 1, cu);
  if (cr) code_tu_residuals(e, c, x0, y0, log2 - 1, 2, cu);
}

static void code_transform_unit(orc_encoder_t *e, orc_cabac_t *c, int x0, int y0, int log2, const orc_cu_t *cu)
{
  code_transform_tree(e, c, x0, y0, log2, 0, 1, 1, cu);
}

/* the three most probable modes (8.4.2) of the luma prediction block at (x, y) */
static void intra_mpm(const orc_encoder_t *e, int x, int y, int cand[3])
{
  /* a neighbour that is not intra-coded, or lies in the CTB row above, counts as DC */
  int a = 1, b = 1;
  if (x > 0) {
    const orc_cu_t *l = &e->cu[(size_t)(y >> 3) * e->w8 + ((x - 1) >> 3)];
    if (l->pred_mode == 1) a = intra_mode_at(l, x - 1, y);
  }
  if (y > 0 && (y & (CTB - 1))) {
    const orc_cu_t *u = &e->cu[(size_t)((y - 1) >> 3) * e->w8 + (x >> 3)];
    if (u->pred_mode == 1) b = intra_mode_at(u, x, y - 1);
  }
  if (a == b) {
    if (a < 2) { cand[0] = 0; cand[1] = 1; cand[2] = 26; }
    else { cand[0] = a; cand[1] = 2 + ((a + 29) % 32); cand[2] = 2 + ((a - 2 + 1) % 32); }
  } else {
    cand[0] = a; cand[1] = b;
    cand[2] = (a != 0 && b != 0) ? 0 : ((a != 1 && b != 1) ? 1 : 26);
  }
}

/* part_mode, prev_intra_luma_pred_flag / mpm_idx / rem_intra_luma_pred_mode (8.4.2) of every prediction
 * block (the flags of all blocks first, 7.3.8.5) and intra_chroma_pred_mode */
static void code_intra_modes(orc_encoder_t *e, orc_cabac_t *c, int x0, int y0, int log2, const orc_cu_t *cu)
{
  const int nxn = cu->flags & 1, parts = nxn ? 4 : 1;
  if (log2 == 3) orc_cabac_bin(c, CTX_PART_MODE, !nxn);          /* PART_2Nx2N / PART_NxN */
  int mpm[4], cand[4][3], mode[4];
  for (int b = 0; b < parts; b++) {
    const int xb = x0 + 4 * (b & 1), yb = y0 + 4 * (b >> 1);
    intra_mpm(e, xb, yb, cand[b]);
    mode[b] = intra_mode_at(cu, xb, yb);
    mpm[b] = -1;
    for (int i = 0; i < 3; i++) if (cand[b][i] == mode[b]) mpm[b] = i;
    orc_cabac_bin(c, CTX_PREV_INTRA_LUMA, mpm[b] >= 0);
  }
  for (int b = 0; b < parts; b++) {
    if (mpm[b] >= 0) {
      orc_cabac_bypass(c, mpm[b] > 0);
      if (mpm[b] > 0) orc_cabac_bypass(c, mpm[b] > 1);
    } else {
      int *cd = cand[b];
      if (cd[0] > cd[1]) { int t = cd[0]; cd[0] = cd[1]; cd[1] = t; }
      if (cd[0] > cd[2]) { int t = cd[0]; cd[0] = cd[2]; cd[2] = t; }
      if (cd[1] > cd[2]) { int t = cd[1]; cd[1] = cd[2]; cd[2] = t; }
      int rem = mode[b];
      for (int i = 2; i >= 0; i--) if (rem > cd[i]) rem--;
      orc_cabac_bypass_bits(c, (uint32_t)rem, 5);
    }
  }
  /* intra_chroma_pred_mode: 4 = derived from the luma mode (of the first block); 0..3 = planar, vertical,
   * horizontal, DC, where the one that equals the luma mode stands for mode 34 */
  static const uint8_t base[4] = {0, 26, 10, 1};
  int cidx = 4;
  if (cu->chroma_mode != cu->intra_mode)
    for (int k = 0; k < 4; k++)
      if (cu->chroma_mode == (base[k] == cu->intra_mode ? 34 : base[k])) cidx = k;
  orc_cabac_bin(c, CTX_INTRA_CHROMA, cidx != 4);
  if (cidx != 4) orc_cabac_bypass_bits(c, (uint32_t)cidx, 2);
}

static void code_cu(orc_encoder_t *e, orc_cabac_t *c, int x0, int y0, int log2)
{
  orc_cu_t *cu = &e->cu[(size_t)(y0 >> 3) * e->w8 + (x0 >> 3)];
  const int n = 1 << log2;
  orc_cu_t upd = *cu;
  if (!e->is_idr) {
    int ctx = 0;
    if (x0 > 0) ctx += e->cu[(size_t)(y0 >> 3) * e->w8 + ((x0 - 1) >> 3)].skip;
    if (y0 > 0) ctx += e->cu[(size_t)((y0 - 1) >> 3) * e->w8 + (x0 >> 3)].skip;
    if (cu->pred_mode == 1) {                                     /* intra CU in a P slice */
      orc_cabac_bin(c, CTX_SKIP + ctx, 0);
      orc_cabac_bin(c, CTX_PRED_MODE, 1);
      code_intra_modes(e, c, x0, y0, log2, cu);
      code_transform_unit(e, c, x0, y0, log2, cu);
      return;
    }
    mv_t mc[MAX_MERGE], ac[2];
    merge_candidates(e, x0, y0, n, mc);
    int midx = -1;
    for (int i = 0; i < MAX_MERGE && midx < 0; i++)
      if (mc[i].x == cu->mvx && mc[i].y == cu->mvy && mc[i].ref == cu->ref_idx) midx = i;
    const int root_cbf = (cu->flags & 2) != 0;
    int skip = midx >= 0 && !root_cbf;
    orc_cabac_bin(c, CTX_SKIP + ctx, skip);
    upd.skip = (uint8_t)skip;
    upd.merge_idx = (uint8_t)(midx >= 0 ? midx : 0xff);
    if (midx >= 0) {
      if (!skip) {
        orc_cabac_bin(c, CTX_PRED_MODE, 0);
        orc_cabac_bin(c, CTX_PART_MODE, 1);
        orc_cabac_bin(c, CTX_MERGE_FLAG, 1);
      }
      orc_cabac_bin(c, CTX_MERGE_IDX, midx > 0);                 /* TR, cMax 4: first bin coded */
      for (int i = 1; i < MAX_MERGE - 1 && midx >= i; i++) orc_cabac_bypass(c, midx > i);
      if (!skip) code_transform_unit(e, c, x0, y0, log2, cu);    /* rqt_root_cbf inferred 1 */
    } else {
      orc_cabac_bin(c, CTX_PRED_MODE, 0);
      orc_cabac_bin(c, CTX_PART_MODE, 1);
      orc_cabac_bin(c, CTX_MERGE_FLAG, 0);
      if (e->n_refs > 1) {                                       /* ref_idx_l0: TR cMax n_refs - 1, two context-coded bins, then bypass */
        const int r = cu->ref_idx, cmax = e->n_refs - 1;
        for (int i = 0; i < r; i++) { if (i < 2) orc_cabac_bin(c, CTX_REF_IDX + i, 1); else orc_cabac_bypass(c, 1); }
        if (r < cmax) { if (r < 2) orc_cabac_bin(c, CTX_REF_IDX + r, 0); else orc_cabac_bypass(c, 0); }
      }
      amvp_candidates(e, x0, y0, n, cu->ref_idx, ac);
      int b0 = mv_comp_bits(cu->mvx - ac[0].x) + mv_comp_bits(cu->mvy - ac[0].y);
      int b1 = mv_comp_bits(cu->mvx - ac[1].x) + mv_comp_bits(cu->mvy - ac[1].y);
      int pi = b1 < b0;
      upd.mvp_idx = (uint8_t)pi;
      code_mvd(c, cu->mvx - ac[pi].x, cu->mvy - ac[pi].y);
      orc_cabac_bin(c, CTX_MVP_IDX, pi);
      orc_cabac_bin(c, CTX_RQT_ROOT_CBF, root_cbf);
      if (root_cbf) code_transform_unit(e, c, x0, y0, log2, cu);
    }
  } else {
    code_intra_modes(e, c, x0, y0, log2, cu);
    code_transform_unit(e, c, x0, y0, log2, cu);
  }
  set_cu(e, x0, y0, log2, &upd);
}

static void code_quadtree(orc_encoder_t *e, orc_cabac_t *c, int x0, int y0, int log2, int depth)
{
  if (x0 >= e->w || y0 >= e->h) return;
  const int n = 1 << log2;
  const orc_cu_t *cu = &e->cu[(size_t)(y0 >> 3) * e->w8 + (x0 >> 3)];
  int split;
  if (x0 + n <= e->w && y0 + n <= e->h && log2 > 3) {
    split = cu->log2_size < log2;
    int ctx = 0;
    if (x0 > 0) ctx += (CTB_LOG2 - e->cu[(size_t)(y0 >> 3) * e->w8 + ((x0 - 1) >> 3)].log2_size) > depth;
    if (y0 > 0) ctx += (CTB_LOG2 - e->cu[(size_t)((y0 - 1) >> 3) * e->w8 + (x0 >> 3)].log2_size) > depth;
    orc_cabac_bin(c, CTX_SPLIT_CU + ctx, split);
  } else {
    split = log2 > 3;
  }
  if (split) {
    int hn = n / 2;
    code_quadtree(e, c, x0, y0, log2 - 1, depth + 1);
    code_quadtree(e, c, x0 + hn, y0, log2 - 1, depth + 1);
    code_quadtree(e, c, x0, y0 + hn, log2 - 1, depth + 1);
    code_quadtree(e, c, x0 + hn, y0 + hn, log2 - 1, depth + 1);
  } else {
    code_cu(e, c, x0, y0, log2);
  }
}

/* ---- parameter sets and slice header ------------------------------------------------------ */

static void put_ptl(orc_bits_t *b, int level_idc)
{
  orc_bits_put(b, 0, 2);            /* general_profile_space */
  orc_bits_put(b, 0, 1);            /* general_tier_flag */
  orc_bits_put(b, 1, 5);            /* general_profile_idc = Main */
  orc_bits_put(b, 0x60000000u, 32); /* compatibility flags: Main (1) and Main 10 (2) */
  orc_bits_put(b, 1, 1);            /* progressive_source */
  orc_bits_put(b, 0, 1);            /* interlaced_source */
  orc_bits_put(b, 0, 1);            /* non_packed_constraint */
  orc_bits_put(b, 1, 1);            /* frame_only_constraint */
  orc_bits_put(b, 0, 32); orc_bits_put(b, 0, 11);   /* 43 reserved zero bits */
  orc_bits_put(b, 0, 1);            /* general_inbld_flag / reserved */
  orc_bits_put(b, (uint32_t)level_idc, 8);
}

static int level_for(int w, int h)
{
  long px = (long)w * h;
  return px <= 552960 ? 93 : px <= 983040 ? 120 : px <= 2228224 ? 123 : px <= 8912896 ? 153 : 183;
}

static size_t write_nal(uint8_t *out, size_t cap, int type, const uint8_t *rbsp, size_t n)
{
  if (cap < 6) return n + 6 + n / 2;
  out[0] = 0; out[1] = 0; out[2] = 0; out[3] = 1;
  out[4] = (uint8_t)(type << 1); out[5] = 1;
  return 6 + orc_nal_escape(rbsp, n, out + 6, cap - 6);
}

static size_t write_parameter_sets(const orc_encoder_t *e, uint8_t *out, size_t cap)
{
  uint8_t tmp[2048];                  /* scaling list data may run to a few hundred bytes */
  orc_bits_t b;
  size_t o = 0;
  int level = level_for(e->w, e->h);
  /* VPS (7.3.2.1) */
  orc_bits_init(&b, tmp, sizeof(tmp));
  orc_bits_put(&b, 0, 4); orc_bits_put(&b, 1, 1); orc_bits_put(&b, 1, 1);
  orc_bits_put(&b, 0, 6); orc_bits_put(&b, 0, 3); orc_bits_put(&b, 1, 1);
  orc_bits_put(&b, 0xffff, 16);
  put_ptl(&b, level);
  orc_bits_put(&b, 0, 1);             /* vps_sub_layer_ordering_info_present_flag */
  orc_bits_ue(&b, (uint32_t)imax(1, e->cfg.refs)); orc_bits_ue(&b, 0); orc_bits_ue(&b, 0);   /* max_dec_pic_buffering_minus1, ... */
  orc_bits_put(&b, 0, 6);             /* vps_max_layer_id */
  orc_bits_ue(&b, 0);                 /* vps_num_layer_sets_minus1 */
  orc_bits_put(&b, 0, 1);             /* vps_timing_info_present_flag */
  orc_bits_put(&b, 0, 1);             /* vps_extension_flag */
  orc_bits_trailing(&b);
  o += write_nal(out + o, cap - o, 32, tmp, orc_bits_bytes(&b));
  /* SPS (7.3.2.2) */
  orc_bits_init(&b, tmp, sizeof(tmp));
  orc_bits_put(&b, 0, 4); orc_bits_put(&b, 0, 3); orc_bits_put(&b, 1, 1);
  put_ptl(&b, level);
  orc_bits_ue(&b, 0);                 /* sps_seq_parameter_set_id */
  orc_bits_ue(&b, 1);                 /* chroma_format_idc 4:2:0 */
  orc_bits_ue(&b, (uint32_t)e->w); orc_bits_ue(&b, (uint32_t)e->h);
  if (e->cfg.conf_right || e->cfg.conf_bottom) {
    orc_bits_put(&b, 1, 1);           /* conformance_window_flag: offsets in chroma samples (SubWidthC = SubHeightC = 2) */
    orc_bits_ue(&b, 0); orc_bits_ue(&b, (uint32_t)e->cfg.conf_right / 2);
    orc_bits_ue(&b, 0); orc_bits_ue(&b, (uint32_t)e->cfg.conf_bottom / 2);
  } else {
    orc_bits_put(&b, 0, 1);
  }
  orc_bits_ue(&b, 0); orc_bits_ue(&b, 0);       /* bit depths - 8 */
  orc_bits_ue(&b, 4);                 /* log2_max_pic_order_cnt_lsb_minus4 -> 8 bits */
  orc_bits_put(&b, 0, 1);             /* sps_sub_layer_ordering_info_present_flag */
  orc_bits_ue(&b, (uint32_t)imax(1, e->cfg.refs)); orc_bits_ue(&b, 0); orc_bits_ue(&b, 0);
  orc_bits_ue(&b, 0);                 /* log2_min_luma_coding_block_size_minus3 -> 8 */
  orc_bits_ue(&b, 3);                 /* log2_diff_max_min -> 64 */
  orc_bits_ue(&b, 0);                 /* log2_min_luma_transform_block_size_minus2 -> 4 */
  orc_bits_ue(&b, 3);                 /* log2_diff_max_min transform -> 32 */
  orc_bits_ue(&b, (uint32_t)e->cfg.tr_depth); orc_bits_ue(&b, (uint32_t)e->cfg.tr_depth);   /* max_transform_hierarchy_depth_inter / intra */
  orc_bits_put(&b, e->cfg.scaling_list ? 1 : 0, 1);       /* scaling_list_enabled_flag */
  if (e->cfg.scaling_list) {
    orc_bits_put(&b, e->cfg.scaling_list == 2 ? 1 : 0, 1);  /* sps_scaling_list_data_present_flag (else the default lists) */
    if (e->cfg.scaling_list == 2) orc_scaling_write(&b, &e->sl);
  }
  orc_bits_put(&b, 0, 1);             /* amp_enabled_flag */
  orc_bits_put(&b, e->cfg.sao ? 1 : 0, 1);          /* sample_adaptive_offset_enabled_flag */
  orc_bits_put(&b, 0, 1);             /* pcm_enabled_flag */
  orc_bits_ue(&b, 1);                 /* num_short_term_ref_pic_sets */
  orc_bits_ue(&b, (uint32_t)imax(1, e->cfg.refs)); orc_bits_ue(&b, 0);      /* num_negative_pics, num_positive_pics = 0 */
  for (int r = 0; r < imax(1, e->cfg.refs); r++) { orc_bits_ue(&b, 0); orc_bits_put(&b, 1, 1); }   /* delta_poc_s0_minus1 = 0, used_by_curr_pic_s0 */
  orc_bits_put(&b, 0, 1);             /* long_term_ref_pics_present_flag */
  orc_bits_put(&b, e->cfg.tmvp ? 1 : 0, 1);     /* sps_temporal_mvp_enabled_flag */
  orc_bits_put(&b, e->cfg.strong_intra ? 1 : 0, 1);   /* strong_intra_smoothing_enabled_flag */
  if (e->cfg.fps_num > 0 && e->cfg.fps_den > 0) {
    /* VUI (E.2.1) carrying only the timing: the reference copies the decoder's frame rate into
     * vInfo (openhevcfilter.cpp:232-233) and DisplayFilter divides by it (displayfilter.cpp:153) */
    orc_bits_put(&b, 1, 1);           /* vui_parameters_present_flag */
    orc_bits_put(&b, 0, 1);           /* aspect_ratio_info_present_flag */
    orc_bits_put(&b, 0, 1);           /* overscan_info_present_flag */
    orc_bits_put(&b, 0, 1);           /* video_signal_type_present_flag */
    orc_bits_put(&b, 0, 1);           /* chroma_loc_info_present_flag */
    orc_bits_put(&b, 0, 3);           /* neutral_chroma_indication, field_seq, frame_field_info_present */
    orc_bits_put(&b, 0, 1);           /* default_display_window_flag */
    orc_bits_put(&b, 1, 1);           /* vui_timing_info_present_flag */
    orc_bits_put(&b, (uint32_t)e->cfg.fps_den, 32);   /* vui_num_units_in_tick */
    orc_bits_put(&b, (uint32_t)e->cfg.fps_num, 32);   /* vui_time_scale */
    orc_bits_put(&b, 0, 1);           /* vui_poc_proportional_to_timing_flag */
    orc_bits_put(&b, 0, 1);           /* vui_hrd_parameters_present_flag */
    orc_bits_put(&b, 0, 1);           /* bitstream_restriction_flag */
  } else {
    orc_bits_put(&b, 0, 1);           /* vui_parameters_present_flag */
  }
  orc_bits_put(&b, 0, 1);             /* sps_extension_present_flag */
  orc_bits_trailing(&b);
  o += write_nal(out + o, cap - o, 33, tmp, orc_bits_bytes(&b));
  /* PPS (7.3.2.3) */
  orc_bits_init(&b, tmp, sizeof(tmp));
  orc_bits_ue(&b, 0); orc_bits_ue(&b, 0);
  orc_bits_put(&b, 0, 1);             /* dependent_slice_segments_enabled_flag */
  orc_bits_put(&b, 0, 1);             /* output_flag_present_flag */
  orc_bits_put(&b, 0, 3);             /* num_extra_slice_header_bits */
  orc_bits_put(&b, e->cfg.sign_hiding ? 1 : 0, 1);    /* sign_data_hiding_enabled_flag */
  orc_bits_put(&b, e->cfg.cabac_init ? 1 : 0, 1);   /* cabac_init_present_flag */
  orc_bits_ue(&b, 0); orc_bits_ue(&b, 0);       /* num_ref_idx_l0/l1_default_active_minus1 */
  orc_bits_se(&b, 0);                 /* init_qp_minus26 */
  orc_bits_put(&b, 0, 1);             /* constrained_intra_pred_flag */
  orc_bits_put(&b, 0, 1);             /* transform_skip_enabled_flag */
  orc_bits_put(&b, e->cfg.qp_delta ? 1 : 0, 1);     /* cu_qp_delta_enabled_flag */
  if (e->cfg.qp_delta) orc_bits_ue(&b, 0);          /* diff_cu_qp_delta_depth: one quantisation group per CTB */
  orc_bits_se(&b, e->cfg.cb_qp_offset); orc_bits_se(&b, e->cfg.cr_qp_offset);       /* pps_cb_qp_offset, pps_cr_qp_offset */
  orc_bits_put(&b, 0, 1);             /* pps_slice_chroma_qp_offsets_present_flag */
  orc_bits_put(&b, 0, 1);             /* weighted_pred_flag */
  orc_bits_put(&b, 0, 1);             /* weighted_bipred_flag */
  orc_bits_put(&b, 0, 1);             /* transquant_bypass_enabled_flag */
  orc_bits_put(&b, e->cfg.tile_cols > 1 || e->cfg.tile_rows > 1, 1);        /* tiles_enabled_flag */
  orc_bits_put(&b, e->cfg.no_wpp ? 0 : 1, 1);       /* entropy_coding_sync_enabled_flag (WPP) */
  if (e->cfg.tile_cols > 1 || e->cfg.tile_rows > 1) {
    orc_bits_ue(&b, (uint32_t)(imax(e->cfg.tile_cols, 1) - 1));   /* num_tile_columns_minus1 */
    orc_bits_ue(&b, (uint32_t)(imax(e->cfg.tile_rows, 1) - 1));   /* num_tile_rows_minus1 */
    orc_bits_put(&b, 1, 1);                              /* uniform_spacing_flag */
    orc_bits_put(&b, 0, 1);                              /* loop_filter_across_tiles_enabled_flag */
  }
  orc_bits_put(&b, 1, 1);             /* pps_loop_filter_across_slices_enabled_flag */
  if (e->cfg.deblock && !e->cfg.beta_offset_div2 && !e->cfg.tc_offset_div2) {
    orc_bits_put(&b, 0, 1);           /* deblocking_filter_control_present_flag */
  } else {
    orc_bits_put(&b, 1, 1);
    orc_bits_put(&b, 0, 1);           /* deblocking_filter_override_enabled_flag */
    orc_bits_put(&b, e->cfg.deblock ? 0 : 1, 1);      /* pps_deblocking_filter_disabled_flag */
    if (e->cfg.deblock) { orc_bits_se(&b, e->cfg.beta_offset_div2); orc_bits_se(&b, e->cfg.tc_offset_div2); }
  }
  orc_bits_put(&b, e->cfg.scaling_list == 3 ? 1 : 0, 1);  /* pps_scaling_list_data_present_flag */
  if (e->cfg.scaling_list == 3) orc_scaling_write(&b, &e->sl);
  orc_bits_put(&b, 0, 1);             /* lists_modification_present_flag */
  orc_bits_ue(&b, 0);                 /* log2_parallel_merge_level_minus2 */
  orc_bits_put(&b, 0, 1);             /* slice_segment_header_extension_present_flag */
  orc_bits_put(&b, 0, 1);             /* pps_extension_present_flag */
  orc_bits_trailing(&b);
  o += write_nal(out + o, cap - o, 34, tmp, orc_bits_bytes(&b));
  return o;
}

/* ---- minimal MD5 for the decoded picture hash SEI (test use) ------------------------------ */

typedef struct { uint32_t s[4]; uint64_t len; uint8_t buf[64]; int fill; } md5_t;
static void md5_block(uint32_t s[4], const uint8_t *p)
{
  static const uint8_t r[64] = {7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 5, 9, 14, 20, 5, 9, 14, 20,
    5, 9, 14, 20, 5, 9, 14, 20, 4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23, 6, 10, 15, 21, 6, 10, 15, 21,
    6, 10, 15, 21, 6, 10, 15, 21};
  static const uint32_t k[64] = {
    0xd76aa478, 0xe8c7b756, 0x242070db, 0xc1bdceee, 0xf57c0faf, 0x4787c62a, 0xa8304613, 0xfd469501,
    0x698098d8, 0x8b44f7af, 0xffff5bb1, 0x895cd7be, 0x6b901122, 0xfd987193, 0xa679438e, 0x49b40821,
    0xf61e2562, 0xc040b340, 0x265e5a51, 0xe9b6c7aa, 0xd62f105d, 0x02441453, 0xd8a1e681, 0xe7d3fbc8,
    0x21e1cde6, 0xc33707d6, 0xf4d50d87, 0x455a14ed, 0xa9e3e905, 0xfcefa3f8, 0x676f02d9, 0x8d2a4c8a,
    0xfffa3942, 0x8771f681, 0x6d9d6122, 0xfde5380c, 0xa4beea44, 0x4bdecfa9, 0xf6bb4b60, 0xbebfbc70,
    0x289b7ec6, 0xeaa127fa, 0xd4ef3085, 0x04881d05, 0xd9d4d039, 0xe6db99e5, 0x1fa27cf8, 0xc4ac5665,
    0xf4292244, 0x432aff97, 0xab9423a7, 0xfc93a039, 0x655b59c3, 0x8f0ccc92, 0xffeff47d, 0x85845dd1,
    0x6fa87e4f, 0xfe2ce6e0, 0xa3014314, 0x4e0811a1, 0xf7537e82, 0xbd3af235, 0x2ad7d2bb, 0xeb86d391};
  uint32_t w[16], a = s[0], b = s[1], c = s[2], d = s[3];
  for (int i = 0; i < 16; i++) w[i] = p[4 * i] | (p[4 * i + 1] << 8) | (p[4 * i + 2] << 16) | ((uint32_t)p[4 * i + 3] << 24);
  for (int i = 0; i < 64; i++) {
    uint32_t f; int g;
    if (i < 16) { f = (b & c) | (~b & d); g = i; }
    else if (i < 32) { f = (d & b) | (~d & c); g = (5 * i + 1) & 15; }
    else if (i < 48) { f = b ^ c ^ d; g = (3 * i + 5) & 15; }
    else { f = c ^ (b | ~d); g = (7 * i) & 15; }
    uint32_t t = d; d = c; c = b;
    uint32_t x = a + f + k[i] + w[g];
    b = b + ((x << r[i]) | (x >> (32 - r[i])));
    a = t;
  }
  s[0] += a; s[1] += b; s[2] += c; s[3] += d;
}
static void md5_init(md5_t *m) { m->s[0] = 0x67452301; m->s[1] = 0xefcdab89; m->s[2] = 0x98badcfe; m->s[3] = 0x10325476; m->len = 0; m->fill = 0; }
static void md5_update(md5_t *m, const uint8_t *p, size_t n)
{
  m->len += n;
  while (n) {
    size_t k = 64 - (size_t)m->fill; if (k > n) k = n;
    memcpy(m->buf + m->fill, p, k); m->fill += (int)k; p += k; n -= k;
    if (m->fill == 64) { md5_block(m->s, m->buf); m->fill = 0; }
  }
}
static void md5_final(md5_t *m, uint8_t out[16])
{
  uint64_t bits = m->len * 8; uint8_t pad = 0x80, z = 0;
  md5_update(m, &pad, 1);
  while (m->fill != 56) md5_update(m, &z, 1);
  uint8_t l[8]; for (int i = 0; i < 8; i++) l[i] = (uint8_t)(bits >> (8 * i));
  md5_update(m, l, 8);
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) out[4 * i + j] = (uint8_t)(m->s[i] >> (8 * j));
}

static size_t write_hash_sei(const orc_encoder_t *e, uint8_t *out, size_t cap)
{
  uint8_t rbsp[64]; size_t n = 0;
  rbsp[n++] = 132;            /* payloadType decoded picture hash */
  rbsp[n++] = 1 + 3 * 16;     /* payloadSize */
  rbsp[n++] = 0;              /* hash_type MD5 */
  for (int c = 0; c < 3; c++) {
    md5_t m; md5_init(&m);
    int pw = c ? e->cw : e->w, ph = c ? e->ch : e->h;
    md5_update(&m, plane(e->rec, e->w, e->h, c), (size_t)pw * ph);
    md5_final(&m, rbsp + n); n += 16;
  }
  rbsp[n++] = 0x80;           /* rbsp_trailing_bits */
  return write_nal(out, cap, 40, rbsp, n);     /* SUFFIX_SEI_NUT */
}

/* ---- slice ------------------------------------------------------------------------------------ */

/* slice segment header (7.3.6.1) + slice data -> one NAL.  n_sub substreams, sub_esc[i] = size of
 * substream i once emulation prevention bytes are in, data = the unescaped substreams back to back. */
static int assemble_slice(const orc_encoder_t *e, int n_sub, const size_t *sub_esc, const uint8_t *data, size_t pos,
                          uint8_t *out, size_t cap, size_t *written)
{
  uint8_t hdr[8192];
  orc_bits_t b;
  orc_bits_init(&b, hdr, sizeof(hdr));
  orc_bits_put(&b, 1, 1);                               /* first_slice_segment_in_pic_flag */
  if (e->is_idr) orc_bits_put(&b, 0, 1);                /* no_output_of_prior_pics_flag */
  orc_bits_ue(&b, 0);                                   /* slice_pic_parameter_set_id */
  orc_bits_ue(&b, e->is_idr ? 2 : 1);                   /* slice_type */
  if (!e->is_idr) {
    orc_bits_put(&b, (uint32_t)(e->poc & 255), 8);      /* slice_pic_order_cnt_lsb */
    if (e->n_refs == imax(1, e->cfg.refs)) {
      orc_bits_put(&b, 1, 1);                           /* short_term_ref_pic_set_sps_flag */
    } else {                                            /* fewer pictures exist yet: the set is written here */
      orc_bits_put(&b, 0, 1);
      orc_bits_put(&b, 0, 1);                           /* inter_ref_pic_set_prediction_flag (stRpsIdx = 1) */
      orc_bits_ue(&b, (uint32_t)e->n_refs); orc_bits_ue(&b, 0);
      for (int r = 0; r < e->n_refs; r++) { orc_bits_ue(&b, 0); orc_bits_put(&b, 1, 1); }
    }
    if (e->cfg.tmvp) orc_bits_put(&b, 1, 1);            /* slice_temporal_mvp_enabled_flag */
  }
  if (e->cfg.sao) orc_bits_put(&b, 3, 2);               /* slice_sao_luma_flag, slice_sao_chroma_flag */
  if (!e->is_idr) {
    orc_bits_put(&b, e->n_refs != 1, 1);                /* num_ref_idx_active_override_flag */
    if (e->n_refs != 1) orc_bits_ue(&b, (uint32_t)(e->n_refs - 1));
    if (e->cfg.cabac_init) orc_bits_put(&b, 1, 1);      /* cabac_init_flag: P slices start from the B-slice tables */
    if (e->cfg.tmvp && e->n_refs > 1) orc_bits_ue(&b, 0);   /* collocated_ref_idx */
    orc_bits_ue(&b, 5 - MAX_MERGE);                     /* five_minus_max_num_merge_cand */
  }
  orc_bits_se(&b, e->cfg.qp - 26);                      /* slice_qp_delta */
  if (e->cfg.deblock || e->cfg.sao) orc_bits_put(&b, 1, 1);   /* slice_loop_filter_across_slices_enabled_flag */
  orc_bits_ue(&b, (uint32_t)(n_sub - 1));               /* num_entry_point_offsets */
  if (n_sub > 1) {
    size_t mx = 1;
    for (int r = 0; r < n_sub - 1; r++) if (sub_esc[r] > mx) mx = sub_esc[r];
    int len = 1;
    while (((mx - 1) >> len) > 0) len++;
    orc_bits_ue(&b, (uint32_t)(len - 1));               /* offset_len_minus1 */
    for (int r = 0; r < n_sub - 1; r++) orc_bits_put(&b, (uint32_t)(sub_esc[r] - 1), len);
  }
  orc_bits_trailing(&b);                                /* byte_alignment() */
  if (b.overflow) return -1;
  size_t hl = orc_bits_bytes(&b);
  /* assemble the NAL: header + substreams, escaped as one RBSP */
  uint8_t *rbsp = (uint8_t *)malloc(hl + pos);
  memcpy(rbsp, hdr, hl);
  memcpy(rbsp + hl, data, pos);
  size_t n = write_nal(out, cap, e->is_idr ? 19 : 1, rbsp, hl + pos);
  free(rbsp);
  *written = n;
  return n <= cap ? 0 : -1;
}

static int encode_slice(orc_encoder_t *e, uint8_t *out, size_t cap, size_t *written)
{
  const int rows = e->ctb_rows, cols = e->ctb_cols;
  size_t *sub_len = (size_t *)calloc((size_t)rows, sizeof(size_t));
  size_t *sub_esc = (size_t *)calloc((size_t)rows, sizeof(size_t));
  orc_cabac_t cab, saved;
  memset(&cab, 0, sizeof(cab));
  memset(&saved, 0, sizeof(saved));
  size_t pos = 0;
  e->bins = 0;
  const int wpp = !e->cfg.no_wpp;
  const int n_sub = wpp ? rows : 1;
  orc_bits_t bits;
  for (int r = 0; r < rows; r++) {
    if (wpp || r == 0) {
      orc_bits_init(&bits, e->sub + pos, e->sub_cap - pos);
      if (r == 0 || cols < 2) orc_cabac_init_contexts(&cab, e->is_idr ? 0 : (e->cfg.cabac_init ? 2 : 1), e->cfg.qp);
      else memcpy(cab.ctx, saved.ctx, sizeof(cab.ctx));          /* WPP sync from CTU 1 of the row above */
      orc_cabac_start(&cab, &bits);
    }
    for (int cidx = 0; cidx < cols; cidx++) {
      if (e->cfg.sao) code_sao(e, &cab, cidx, r);
      e->delta_coded = 0;
      code_quadtree(e, &cab, cidx * CTB, r * CTB, CTB_LOG2, 0);
      if (cidx == 1) memcpy(saved.ctx, cab.ctx, sizeof(cab.ctx));
      int last_in_slice = r == rows - 1 && cidx == cols - 1 && !e->cfg.more_tiles;
      int last_in_sub = cidx == cols - 1 && (wpp || r == rows - 1);
      orc_cabac_terminate(&cab, last_in_slice);                  /* end_of_slice_segment_flag */
      if (last_in_sub && !last_in_slice) orc_cabac_terminate(&cab, 1);        /* end_of_subset_one_bit */
    }
    if (!wpp && r < rows - 1) continue;                          /* one substream: keep coding */
    orc_cabac_finish(&cab);
    if (bits.overflow) { free(sub_len); free(sub_esc); return -1; }
    const int si = wpp ? r : 0;
    sub_len[si] = orc_bits_bytes(&bits);
    /* emulation prevention bytes this substream will receive inside the NAL */
    sub_esc[si] = orc_nal_escape(e->sub + pos, sub_len[si], NULL, 0);
    pos += sub_len[si];
    e->bins += cab.bins; cab.bins = 0;
  }
  if (e->cfg.raw_slice_data) {
    size_t need = pos + 4 * (size_t)n_sub, o = 0, off = 0;
    free(sub_esc);
    if (need > cap) { free(sub_len); return -1; }
    for (int r = 0; r < n_sub; r++) {
      uint32_t n = (uint32_t)sub_len[r];
      out[o++] = (uint8_t)n; out[o++] = (uint8_t)(n >> 8); out[o++] = (uint8_t)(n >> 16); out[o++] = (uint8_t)(n >> 24);
      memcpy(out + o, e->sub + off, n);
      o += n; off += n;
    }
    free(sub_len);
    *written = o;
    return 0;
  }
  int rc = assemble_slice(e, n_sub, sub_esc, e->sub, pos, out, cap, written);
  free(sub_len); free(sub_esc);
  return rc;
}

int orc_enc_encode(orc_encoder_t *e, const uint8_t *i420, uint8_t *out, int cap)
{
  if (!e || !i420 || !out || cap <= 0) return -1;
  const size_t fsz = (size_t)e->w * e->h * 3 / 2;
  e->src = i420;
  e->is_idr = e->frame_idx == 0 || (e->cfg.intra_period > 0 && e->frame_idx % e->cfg.intra_period == 0);
  if (e->is_idr) e->poc = 0;
  e->n_refs = e->is_idr ? 0 : imin(imax(1, e->cfg.refs), e->n_dpb);
  memset(e->levels, 0, fsz * sizeof(int16_t));
  if (e->cfg.vaq) orc_vaq_offsets(i420, e->w, e->h, e->cfg.vaq, e->vaq_dqp);
  if (e->is_idr) {
    for (int cy = 0; cy < e->h; cy += CTB)
      for (int cx = 0; cx < e->w; cx += CTB) intra_quadtree(e, cx, cy, CTB_LOG2);
  } else {
    inter_frame(e);
  }
  memcpy(e->rec_pre, e->rec, fsz);
  if (e->cfg.qp_delta) derive_cu_qps(e);
  if (e->cfg.deblock) deblock_frame(e);
  if (e->cfg.sao) sao_frame(e);
  size_t o = 0, n = 0;
  if (e->is_idr && !e->cfg.raw_slice_data) o += write_parameter_sets(e, out, (size_t)cap);
  if (o > (size_t)cap) return -1;
  if (encode_slice(e, out + o, (size_t)cap - o, &n) != 0) return -1;
  o += n;
  if (e->cfg.hash_sei && !e->cfg.raw_slice_data) o += write_hash_sei(e, out + o, (size_t)cap - o);
  if (o > (size_t)cap) return -1;
  if (e->is_idr) e->n_dpb = 0;                  /* an IDR empties the decoded picture buffer */
  dpb_insert(e);
  e->frame_idx++;
  e->poc++;
  return (int)o;
}


/* ------------------------------------------------------------------------------------------ */
/* tiles (uniform columns x rows) as independent pictures                                        */

#define MAX_TILES 64
struct orc_tiled {
  orc_encoder_t *hdr;             /* full-size instance: parameter sets, slice header, hash SEI, composite recon */
  int tiles, x0[MAX_TILES], y0[MAX_TILES], wd[MAX_TILES], ht[MAX_TILES];
  orc_encoder_t *strip[MAX_TILES];
  uint8_t *strip_src, *raw;
  size_t raw_cap;
};

orc_tiled_t *orc_tiled_open2(const orc_enc_cfg_t *cfg, int tile_cols, int tile_rows)
{
  if (!cfg || tile_cols < 1 || tile_rows < 1 || tile_cols * tile_rows > MAX_TILES) return NULL;
  const int ctb_cols = (cfg->width + CTB - 1) / CTB, ctb_rows = (cfg->height + CTB - 1) / CTB;
  if (tile_cols > 1 && tile_cols > ctb_cols / 2) return NULL;   /* every tile at least two CTUs wide (WPP sync) */
  if (tile_rows > ctb_rows) return NULL;
  if (cfg->qp_delta) return NULL;                         /* the QP prediction chain of derive_cu_qps assumes WPP rows */
  orc_tiled_t *t = (orc_tiled_t *)calloc(1, sizeof(*t));
  orc_enc_cfg_t hc = *cfg;
  hc.tile_cols = tile_cols; hc.tile_rows = tile_rows; hc.mv_edges = 0; hc.more_tiles = 0; hc.raw_slice_data = 0;
  t->hdr = orc_enc_open(&hc);
  t->tiles = tile_cols * tile_rows;
  if (!t->hdr) { free(t); return NULL; }
  for (int i = 0; i < t->tiles; i++) {                    /* tiles in raster order (6.5.1) */
    const int tc = i % tile_cols, tr = i / tile_cols;
    const int c0 = tc * ctb_cols / tile_cols, c1 = (tc + 1) * ctb_cols / tile_cols;   /* colBd, uniform spacing */
    const int r0 = tr * ctb_rows / tile_rows, r1 = (tr + 1) * ctb_rows / tile_rows;   /* rowBd */
    t->x0[i] = c0 * CTB; t->wd[i] = imin(cfg->width, c1 * CTB) - t->x0[i];
    t->y0[i] = r0 * CTB; t->ht[i] = imin(cfg->height, r1 * CTB) - t->y0[i];
    orc_enc_cfg_t sc = *cfg;
    sc.width = t->wd[i]; sc.height = t->ht[i]; sc.hash_sei = 0; sc.tile_cols = 0; sc.tile_rows = 0; sc.raw_slice_data = 1;
    sc.mv_edges = (tc > 0 ? 1 : 0) | (tc < tile_cols - 1 ? 2 : 0) | (tr > 0 ? 4 : 0) | (tr < tile_rows - 1 ? 8 : 0) | (cfg->mv_edges & 15);
    sc.more_tiles = i < t->tiles - 1;
    t->strip[i] = orc_enc_open(&sc);
    if (!t->strip[i]) { orc_tiled_close(t); return NULL; }
  }
  t->strip_src = (uint8_t *)malloc((size_t)cfg->width * cfg->height * 3 / 2);
  t->raw_cap = (size_t)cfg->width * cfg->height * 3 + 65536;
  t->raw = (uint8_t *)malloc(t->raw_cap);
  return t;
}

orc_tiled_t *orc_tiled_open(const orc_enc_cfg_t *cfg, int tile_cols) { return orc_tiled_open2(cfg, tile_cols, 1); }

void orc_tiled_close(orc_tiled_t *t)
{
  if (!t) return;
  for (int i = 0; i < t->tiles; i++) if (t->strip[i]) orc_enc_close(t->strip[i]);
  orc_enc_close(t->hdr);
  free(t->strip_src); free(t->raw); free(t);
}

const uint8_t *orc_tiled_recon(const orc_tiled_t *t) { return t->hdr->rec; }

/* copies the rectangle (x0, y0, wd, ht) of a packed I420 picture into / out of a packed I420 tile */
static void strip_copy(uint8_t *pic, int w, int h, uint8_t *strip, int x0, int y0, int wd, int ht, int to_strip)
{
  for (int c = 0; c < 3; c++) {
    const int pw = c ? w / 2 : w, sx = c ? x0 / 2 : x0, sy = c ? y0 / 2 : y0, sw = c ? wd / 2 : wd, sh = c ? ht / 2 : ht;
    uint8_t *pp = plane(pic, w, h, c), *sp = plane(strip, wd, ht, c);
    for (int y = 0; y < sh; y++) {
      if (to_strip) memcpy(sp + (size_t)y * sw, pp + (size_t)(sy + y) * pw + sx, sw);
      else memcpy(pp + (size_t)(sy + y) * pw + sx, sp + (size_t)y * sw, sw);
    }
  }
}

int orc_tiled_encode(orc_tiled_t *t, const uint8_t *i420, uint8_t *out, int cap)
{
  orc_encoder_t *e = t->hdr;
  e->is_idr = e->frame_idx == 0 || (e->cfg.intra_period > 0 && e->frame_idx % e->cfg.intra_period == 0);
  if (e->is_idr) e->poc = 0;
  size_t *sub_esc = (size_t *)calloc((size_t)MAX_TILES * 64, sizeof(size_t)), total = 0;
  int n_sub = 0;
  uint8_t *data = (uint8_t *)malloc(t->raw_cap);
  for (int i = 0; i < t->tiles; i++) {
    strip_copy((uint8_t *)i420, e->w, e->h, t->strip_src, t->x0[i], t->y0[i], t->wd[i], t->ht[i], 1);
    int n = orc_enc_encode(t->strip[i], t->strip_src, t->raw, (int)t->raw_cap);
    if (n < 0) { free(data); free(sub_esc); return -1; }
    for (int o = 0; o < n;) {                                /* 4-byte length + substream, repeated */
      uint32_t len = t->raw[o] | (t->raw[o + 1] << 8) | (t->raw[o + 2] << 16) | ((uint32_t)t->raw[o + 3] << 24);
      o += 4;
      memcpy(data + total, t->raw + o, len);
      sub_esc[n_sub++] = orc_nal_escape(t->raw + o, len, NULL, 0);
      total += len; o += (int)len;
    }
    strip_copy(e->rec, e->w, e->h, (uint8_t *)orc_enc_recon(t->strip[i]), t->x0[i], t->y0[i], t->wd[i], t->ht[i], 0);
  }
  size_t o = 0, n = 0;
  e->n_refs = t->strip[0]->n_refs;                        /* what the slice header says about the reference list */
  if (e->is_idr) o += write_parameter_sets(e, out, (size_t)cap);
  int rc = o > (size_t)cap ? -1 : assemble_slice(e, n_sub, sub_esc, data, total, out + o, (size_t)cap - o, &n);
  free(data); free(sub_esc);
  if (rc != 0) return -1;
  o += n;
  if (e->cfg.hash_sei) o += write_hash_sei(e, out + o, (size_t)cap - o);
  if (o > (size_t)cap) return -1;
  e->frame_idx++;
  e->poc++;
  return (int)o;
}
