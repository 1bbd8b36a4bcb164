/*
 * TEST INFRASTRUCTURE ONLY -- constant tables of ITU-T H.265 used by the CPU
 * oracle (restated from the specification text; HM conventions where the
 * standard is silent).  The product keeps its own copies in
 * kvazzup_b200/csrc/hevc_tables.cuh; nothing here is linked into the product.
 *
 * PARITY STATUS for everything under oracle/hevc_*: the third-party encoder
 * the reference calls (Kvazaar 2.3.1, git f49af6386c9c7cddaa9d00e85cfe30cfe1d6a60f,
 * /root/reference/dependencies/kvazaar.cmake:13; call sites
 * /root/reference/src/media/processing/kvazaarfilter.cpp:145-299,435-476) is not
 * vendored and the reference holds no golden vectors for it: PARITY vs KVAZAAR IS
 * UNPINNED.  What IS pinned: every normative element (CABAC, syntax, intra/inter
 * prediction, inverse transform, dequantisation, deblocking) is checked by
 * decoding the oracle's streams with an independent conformant decoder (FFmpeg's
 * native hevc, tests/ffhevc.py) and comparing its output with the oracle's
 * reconstruction bit for bit, plus the per-picture MD5 SEI.
 */
#ifndef ORACLE_HEVC_TABLES_H_
#define ORACLE_HEVC_TABLES_H_
#include <stdint.h>

extern const uint8_t  orc_range_tab_lps[64][4];   /* H.265 Table 9-46 */
extern const uint8_t  orc_trans_idx_lps[64];      /* H.265 Table 9-47 */
extern const int8_t   orc_intra_pred_angle[35];   /* H.265 Table 8-4 (modes 2..34) */
extern const int16_t  orc_inv_angle[35];          /* H.265 Table 8-5 (modes 11..25) */
extern const int8_t   orc_luma_filter[4][8];      /* H.265 Table 8-11 */
extern const int8_t   orc_chroma_filter[8][4];    /* H.265 Table 8-12 */
extern const uint8_t  orc_beta_table[52];         /* H.265 Table 8-12 (beta') */
extern const uint8_t  orc_tc_table[54];           /* H.265 Table 8-12 (tc') */
extern const uint8_t  orc_chroma_qp_table[58];    /* H.265 Table 8-10, ChromaArrayType 1 */
extern const int16_t  orc_quant_scales[6];        /* HM g_quantScales */
extern const uint8_t  orc_level_scale[6];         /* H.265 8.6.4.2 levelScale */
extern const int8_t   orc_dst4[4][4];             /* H.265 (8-xx) DST-VII 4x4 */
extern const uint8_t  orc_sig_ctx_map_4x4[16];    /* H.265 Table 9-41 ctxIdxMap */

/* DCT coefficient transMatrix[k][n] of the N-point transform (N=4,8,16,32) --
 * generated from the 33 distinct magnitudes of the 32-point matrix (8.6.4.2). */
int orc_dct_coef(int N, int k, int n);

/* Context-table layout (one flat array per slice). */
enum {
  CTX_SPLIT_CU = 0,             /* 3 */
  CTX_SKIP = 3,                 /* 3 */
  CTX_MERGE_FLAG = 6,           /* 1 */
  CTX_MERGE_IDX = 7,            /* 1 */
  CTX_PART_MODE = 8,            /* 4 */
  CTX_PRED_MODE = 12,           /* 1 */
  CTX_PREV_INTRA_LUMA = 13,     /* 1 */
  CTX_INTRA_CHROMA = 14,        /* 1 */
  CTX_MVD_GT0 = 15,             /* 1 */
  CTX_MVD_GT1 = 16,             /* 1 */
  CTX_MVP_IDX = 17,             /* 1 */
  CTX_RQT_ROOT_CBF = 18,        /* 1 */
  CTX_SPLIT_TRANSFORM = 19,     /* 3 */
  CTX_CBF_LUMA = 22,            /* 2 */
  CTX_CBF_CHROMA = 24,          /* 4 */
  CTX_LAST_X = 28,              /* 18 */
  CTX_LAST_Y = 46,              /* 18 */
  CTX_CSBF = 64,                /* 4 */
  CTX_SIG = 68,                 /* 42 */
  CTX_GT1 = 110,                /* 24 */
  CTX_GT2 = 134,                /* 6 */
  CTX_CU_QP_DELTA = 140,        /* 2 */
  CTX_SAO_MERGE = 142,          /* 1 */
  CTX_SAO_TYPE = 143,           /* 1 */
  CTX_REF_IDX = 144,            /* 2 */
  CTX_COUNT = 146
};

/* init values, [initType 0=I,1=P,2=B][CTX_COUNT]; 154 where the element cannot occur */
extern const uint8_t orc_ctx_init[3][CTX_COUNT];

/* scan tables: orc_scan[log2-1? see .c] -- position index -> (x | y<<4)?  Provided as functions. */
/* scan_idx: 0 diagonal, 1 horizontal, 2 vertical; blk_log2 in 1..3 (2x2, 4x4, 8x8 grids). */
void orc_scan_pos(int scan_idx, int blk_log2, int i, int *x, int *y);

#endif
