/* TEST INFRASTRUCTURE ONLY -- scaling lists (H.265 7.3.4 scaling_list_data, 7.4.5, 8.6.4.2).
 * Kvazaar: --scaling-list default (the reference sets it at kvazaarfilter.cpp:236-243 from
 * video/scalingList).  Pinned by FFmpeg: streams coded with the default lists and with the test lists
 * below (carried in the SPS or in the PPS) must reconstruct there to the oracle's pictures. */
#ifndef ORACLE_HEVC_SCALING_H_
#define ORACLE_HEVC_SCALING_H_
#include <stdint.h>
#include "hevc_cabac.h"

/* How list (size_id, matrix_id) is signalled: inferred default (pred_mode_flag 0, delta 0), copied from
 * the list `ref` positions back (pred_mode_flag 0, delta = ref) or coded (DPCM in diagonal scan order). */
enum { ORC_SL_DEFAULT = 0, ORC_SL_COPY = 1, ORC_SL_CODED = 2 };

typedef struct {
  uint8_t list[4][6][64];        /* ScalingList[sizeId][matrixId][i], i in up-right diagonal order (16 entries for sizeId 0) */
  uint8_t dc[4][6];              /* scaling_list_dc_coef_minus8 + 8 (sizeId 2, 3), 16 otherwise */
  uint8_t how[4][6], ref[4][6];  /* signalling (writer only) */
  uint8_t m[4][6][64];           /* expanded: ScalingFactor of sizeId at raster position (y >> s) * 8 + (x >> s)
                                  * (sizeId 0: y * 4 + x); the DC entry of sizeId 2, 3 is dc[][] */
} orc_scaling_t;

/* matrix id: 0..2 = intra Y / Cb / Cr, 3..5 = inter Y / Cb / Cr */
void orc_scaling_default(orc_scaling_t *s);           /* Tables 7-5 / 7-6 */
void orc_scaling_test_lists(orc_scaling_t *s);        /* a mix of coded, copied and default lists (decoder tests) */
int  orc_scaling_factor(const orc_scaling_t *s, int log2n, int matrix, int x, int y);
void orc_scaling_write(orc_bits_t *b, const orc_scaling_t *s);          /* scaling_list_data() */
/* the factors of the default (mode 1) / the test lists (mode 2, 3) as 4 x 6 x 64 raster-order bytes followed by the
 * DC factors of sizeId 2 and 3 (2 x 6 bytes): 1548 bytes */
void orc_scaling_table(int mode, uint8_t *out);

/* quantisation with per-coefficient scale (HM: quantCoef = (quantScale << 4) / m), dequantisation 8.6.4.2 */
int  orc_quant_sl(const int16_t *coeff, int16_t *level, int log2n, int qp, int intra_slice, const orc_scaling_t *s, int matrix);
void orc_dequant_sl(const int16_t *level, int16_t *coeff, int log2n, int qp, const orc_scaling_t *s, int matrix);
#endif
