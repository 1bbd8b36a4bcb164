/* TEST INFRASTRUCTURE ONLY -- scalar restatement of the HEVC block primitives
 * (SURVEY.md 8a-K rows K1-K7).  See hevc_tables.h for scope and pinning status. */
#ifndef ORACLE_HEVC_PRIMS_H_
#define ORACLE_HEVC_PRIMS_H_
#include <stdint.h>

/* K1: sum of absolute differences over a w x h block. */
uint32_t orc_sad(const uint8_t *a, int sa, const uint8_t *b, int sb, int w, int h);
/* K2: Hadamard SATD, HM xCalcHADs convention: 4x4 tiles when w or h is 4
 * ((sum|h|+1)>>1 each), otherwise 8x8 tiles ((sum|h|+2)>>2 each). */
uint32_t orc_satd(const uint8_t *a, int sa, const uint8_t *b, int sb, int w, int h);

/* K5: transforms on contiguous N x N int16 blocks (row-major), N = 1<<log2n, 8-bit video. */
void orc_fdct(const int16_t *resid, int16_t *coeff, int log2n);
void orc_idct(const int16_t *coeff, int16_t *resid, int log2n);
void orc_fdst4(const int16_t *resid, int16_t *coeff);
void orc_idst4(const int16_t *coeff, int16_t *resid);

/* K6: flat quantisation / dequantisation (no scaling lists, no RDOQ, no sign hiding).
 * orc_quant returns the number of non-zero levels. */
int  orc_quant(const int16_t *coeff, int16_t *level, int log2n, int qp, int intra_slice);
void orc_dequant(const int16_t *level, int16_t *coeff, int log2n, int qp);
int  orc_chroma_qp(int qp_y);

/* K4: intra prediction from an already substituted reference array.
 * refs has 4N+1 samples: refs[0..2N-1] = left column from the BOTTOM (p[-1][2N-1]) up to
 * p[-1][0], refs[2N] = corner p[-1][-1], refs[2N+1..4N] = top row p[0][-1]..p[2N-1][-1].
 * Applies 8.4.4.2.3 neighbour filtering when cidx==0 (strong_intra_smoothing off) and the
 * DC / horizontal / vertical edge filters of 8.4.4.2.5-6. */
void orc_intra_predict(const uint8_t *refs, int log2n, int mode, int cidx, uint8_t *dst, int dstride);
/* ... with strong_intra_smoothing_enabled_flag = `strong` (bi-linear smoothing of 32x32 luma neighbours, 8.4.4.2.3) */
void orc_intra_predict2(const uint8_t *refs, int log2n, int mode, int cidx, int strong, uint8_t *dst, int dstride);
/* K3: motion compensated prediction, uni-directional, 8-bit; mv in quarter luma samples.
 * Reference samples outside the picture are edge-clamped (8.5.3.3.3.1). */
void orc_mc_luma(const uint8_t *ref, int stride, int pic_w, int pic_h, int x0, int y0, int w, int h,
                 int mvx, int mvy, uint8_t *dst, int dstride);
/* chroma plane of a 4:2:0 picture: x0,y0,w,h,pic_w,pic_h in chroma samples, mv still in
 * quarter LUMA samples (= eighth chroma samples). */
void orc_mc_chroma(const uint8_t *ref, int stride, int pic_w, int pic_h, int x0, int y0, int w, int h,
                   int mvx, int mvy, uint8_t *dst, int dstride);

/* K7: deblocking of one edge segment of 4 lines (luma) / chroma, in place.
 * `pix` points at q0 of the first line; xstride moves across the edge, ystride along it. */
void orc_deblock_luma_segment(uint8_t *pix, int xstride, int ystride, int bs, int qp);
void orc_deblock_chroma_segment(uint8_t *pix, int xstride, int ystride, int qp_y, int lines);
/* ... with slice_beta_offset_div2 / slice_tc_offset_div2 and the chroma QP offset of the PPS (cQpPicOffset) */
void orc_deblock_luma_segment2(uint8_t *pix, int xstride, int ystride, int bs, int qp, int beta_off, int tc_off);
void orc_deblock_chroma_segment2(uint8_t *pix, int xstride, int ystride, int qp_y, int lines, int c_off, int tc_off);

#endif
