/* TEST INFRASTRUCTURE ONLY -- CPU oracle of the B200 HEVC encoder: the same decision
 * algorithm the CUDA encoder runs, written as plain sequential C directly from the
 * H.265 text, so that (a) its streams can be validated by an independent decoder and
 * (b) the GPU's cu-map, coefficients, reconstruction and bitstream can be compared
 * with it bit for bit.  See hevc_tables.h for scope and pinning status. */
#ifndef ORACLE_HEVC_ENC_H_
#define ORACLE_HEVC_ENC_H_
#include <stdint.h>

/* One entry per 8x8 luma unit (all units of a CU carry the same values). */
typedef struct {
  int16_t mvx, mvy;        /* quarter-sample motion vector (inter) */
  uint8_t log2_size;       /* CU size: 3..6 */
  uint8_t pred_mode;       /* 0 inter, 1 intra */
  uint8_t intra_mode;      /* luma intra prediction mode 0..34 */
  uint8_t cbf;             /* bit0 Y, bit1 Cb, bit2 Cr */
  uint8_t skip;            /* filled by the entropy stage */
  uint8_t merge_idx;       /* 0xff = not merged */
  uint8_t mvp_idx;
  uint8_t qp;              /* luma QP of the CU when cu_qp_delta is enabled (8.6.1), else 0 */
  uint8_t ref_idx;         /* index into reference picture list 0 (inter) */
  uint8_t chroma_mode;     /* intra: resolved chroma prediction mode */
  uint8_t tu_log2;         /* luma size of the transform unit covering this unit */
  uint8_t flags;           /* bit 0: intra NxN (modes of parts 1..3: mvx & 0xff, mvx >> 8, mvy & 0xff); bit 1: coded residual */
} orc_cu_t;

typedef struct {
  int width, height;       /* multiples of 8 */
  int qp;                  /* constant QP, 0..51 */
  int intra_period;        /* IDR every n frames; 0 = first frame only */
  int search_range;        /* full-sample search range (+-), 1..32 */
  int deblock;             /* 1 = in-loop deblocking on */
  int hash_sei;            /* 1 = append MD5 decoded-picture-hash SEI (test use) */
  int qp_delta;            /* 1 = cu_qp_delta_enabled_flag, one quantisation group per CTU (ROI) */
  /* Tile-column mode: this encoder codes one tile (a column strip) of a larger picture as if it were
   * a picture of its own.  mv_edges: bit 0 / bit 1 = the left / right edge is an interior tile edge,
   * so no reference sample beyond it may be touched (the neighbouring tile's samples are there, not
   * edge padding).  more_tiles: tiles follow in the slice, so the last CTU does not end the slice
   * segment.  raw_slice_data: orc_enc_encode returns only the substreams, each prefixed by its
   * 4-byte little-endian length (the compositor writes parameter sets and the slice header). */
  int mv_edges, more_tiles, raw_slice_data;
  int no_wpp;              /* 1 = entropy_coding_sync off: one substream for the whole (strip) picture */
  int subme_satd;          /* 1 = fractional motion refinement by SATD (Kvazaar / HM style) instead of SAD; oracle only */
  int sao;                 /* sample adaptive offset (8.7.3) after deblocking: 1 = on, 2 = on and sao_merge_left / _up
                            * flags are used where a CTU's parameters repeat its neighbour's */
  int tile_cols;           /* > 1: PPS / slice header of a picture with that many uniform tile columns (compositor only) */
  int tr_depth;            /* max_transform_hierarchy_depth_inter = _intra (0..2): transform units may be split into four,
                            * down to 8x8 luma; decided by squared error + lambda * estimated bits */
  int cabac_init;          /* 1 = cabac_init_present_flag, P slices signal cabac_init_flag = 1 (initialisation type 2) */
  int refs;                /* reference pictures (0 or 1 = the previous picture only, up to 4): every CU picks the
                            * reference with the smallest cost; ref_idx_l0, the short-term RPS in the slice header
                            * while fewer pictures exist, AMVP with vector scaling, bS from reference pictures */
  int tmvp;                /* 1 = temporal motion vector prediction (8.5.3.2.8): collocated candidates in the merge and
                            * AMVP lists, from the previous picture's motion field */
  int me_coarse;           /* > 0: two-level motion search -- a coarse level on quarter-resolution pictures (+- me_coarse
                            * coarse samples = 4 * me_coarse luma samples) gives every 32x32 block a second search
                            * centre besides the zero vector; search_range (<= 16) is the window around each centre */
  int intra_in_p;          /* 1 = 16x16 intra CUs in P pictures where inter prediction is poor (scene cuts, uncovered areas) */
  int fps_num, fps_den;    /* both > 0: VUI timing info in the SPS (vui_time_scale / vui_num_units_in_tick); 0 = no VUI */
  int intra_satd;          /* 1 = I pictures: the 35-mode search of blocks >= 8x8 compares Hadamard SATD (8x8 tiles) instead of SAD */
  /* Syntax the GPU encoder does not produce but a Kvazaar-family peer may: streams for the decoder tests. */
  int tu4;                 /* 1 = an 8x8 transform unit may split into four 4x4 luma blocks (and one 4x4 block per chroma
                            * plane, coded after the fourth luma block): DST-VII for intra luma, DCT otherwise; needs
                            * tr_depth >= 1 (NxN CUs split by inference) */
  int intra_sizes;         /* I pictures, each where its search cost is smaller: bit 0 = 8x8 CUs, bit 1 = 32x32 CUs,
                            * bit 2 = NxN partition of 8x8 CUs (four 4x4 prediction blocks) */
  int chroma_modes;        /* 1 = intra_chroma_pred_mode chosen among planar / vertical / horizontal / DC / derived */
  int sign_hiding;         /* 1 = sign_data_hiding_enabled_flag: the sign of the first coefficient of a 4x4 group whose
                            * significant coefficients span more than three scan positions is the parity of the sum */
  int strong_intra;        /* 1 = strong_intra_smoothing_enabled_flag */
  int cb_qp_offset, cr_qp_offset;         /* pps_cb_qp_offset / pps_cr_qp_offset (-12..12) */
  int beta_offset_div2, tc_offset_div2;   /* pps_beta_offset_div2 / pps_tc_offset_div2 (-6..6) */
  int tile_rows;           /* > 1 (or tile_cols > 1): PPS of a picture with that many uniform tile rows (compositor only);
                            * mv_edges bits 2 / 3 then mark the top / bottom edge of a tile as interior */
  int vaq;                 /* variance adaptive quantisation strength (Kvazaar --vaq), 0 = off, 1..20; needs qp_delta:
                            * every picture, each CTU's QP moves by orc_vaq_offsets() on top of the ROI offsets */
  int scaling_list;        /* scaling_list_enabled_flag: 1 = the default lists (Kvazaar --scaling-list default; Tables 7-5 / 7-6),
                            * 2 = test lists carried in the SPS, 3 = test lists carried in the PPS (the SPS then says "default") */
  int conf_right, conf_bottom;   /* conformance window (7.4.3.2.1): that many luma samples (0, 2, 4, 6) at the right / bottom of the
                                  * coded picture are padding and are cropped on output -- a source size that is not a multiple of 8 */
} orc_enc_cfg_t;

typedef struct orc_encoder orc_encoder_t;

void orc_set_threads(int n);      /* OpenMP threads for the P-picture loops (default: 1 via binding) */
int orc_max_threads(void);

orc_encoder_t *orc_enc_open(const orc_enc_cfg_t *cfg);
void orc_enc_close(orc_encoder_t *e);
/* Encode one packed I420 frame; writes one Annex-B access unit (4-byte start codes;
 * VPS+SPS+PPS precede every IDR).  Returns the AU size in bytes or <0 (-needed if cap is short). */
int orc_enc_encode(orc_encoder_t *e, const uint8_t *i420, uint8_t *out, int cap);
/* Per-CTU QP offsets for the pictures that follow (ctb_cols*ctb_rows entries, raster; NULL = none).
 * Needs cfg.qp_delta; CTU QP = clip(qp + dqp, 0, 51). */
int orc_enc_set_ctu_dqp(orc_encoder_t *e, const int8_t *dqp);
/* Slice QP (0..51) of the following pictures. */
int orc_enc_set_qp(orc_encoder_t *e, int qp);
/* Variance adaptive quantisation: the QP offset (-12..12) of every CTU (raster) of one I420 picture
 * (even w, h) at `strength` (1..20), from the CTU's and the picture's sample variance. */
void orc_vaq_offsets(const uint8_t *i420, int w, int h, int strength, int8_t *out);

/* Tile columns (H.265 6.5.1, uniform spacing) coded as independent strips: one encoder per tile with
 * mv_edges / more_tiles / raw_slice_data set, loop filtering across tiles off, substreams
 * concatenated in tile order behind one slice header.  Same call shape as the plain encoder. */
typedef struct orc_tiled orc_tiled_t;
orc_tiled_t *orc_tiled_open(const orc_enc_cfg_t *cfg, int tile_cols);
orc_tiled_t *orc_tiled_open2(const orc_enc_cfg_t *cfg, int tile_cols, int tile_rows);   /* uniform tile grid, tiles in raster order */
void orc_tiled_close(orc_tiled_t *t);
int orc_tiled_encode(orc_tiled_t *t, const uint8_t *i420, uint8_t *out, int cap);
const uint8_t *orc_tiled_recon(const orc_tiled_t *t);            /* packed I420 of the whole picture */

const uint8_t *orc_enc_recon(const orc_encoder_t *e);            /* packed I420, after deblocking */
const uint8_t *orc_enc_recon_predeblock(const orc_encoder_t *e); /* packed I420, before deblocking */
const orc_cu_t *orc_enc_cu_map(const orc_encoder_t *e);          /* (w/8)*(h/8) entries, raster */
const int16_t *orc_enc_levels(const orc_encoder_t *e);           /* quantised levels, I420-shaped int16 */
int orc_enc_last_was_idr(const orc_encoder_t *e);
unsigned long long orc_enc_bins(const orc_encoder_t *e);         /* CABAC bins of the last frame */

#endif
