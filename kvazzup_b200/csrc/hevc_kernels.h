// Launchers of the HEVC kernels (internal header).  Every launcher enqueues on `s` and returns
// the launch status; each counts as one kernel launch for b200_launch_count().
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "hevc_common.h"

namespace b200 {

// P pictures
cudaError_t launch_inter_me(const FrameParams &fp, const uint8_t *src, const uint8_t *ref, CuInfo *cu, cudaStream_t s);
cudaError_t launch_inter_recon(const FrameParams &fp, const uint8_t *src, const uint8_t *ref, uint8_t *rec,
                               int16_t *levels, CuInfo *cu, cudaStream_t s);
cudaError_t launch_inter_modes(const FrameParams &fp, CuInfo *cu, cudaStream_t s);
// quarter-resolution luma (rounded 4x4 means) for the coarse level of the motion search
cudaError_t launch_down4(const uint8_t *plane, int w, int h, uint8_t *out, cudaStream_t s);

// Intra CUs: wavefront over CTUs in ticket order.  `ticket` = 1 int and fp.ctu_done = one int per CTU,
// both zeroed by the launcher.  launch_intra_frame: I picture (mode decision + reconstruction of every
// CU).  launch_intra_in_p: the intra CUs the motion search put into the cu map of a P picture, after
// launch_inter_recon (returns at once when *fp.any_intra == 0).
void intra_wavefront_order(int cols, int rows, int *out);      // host: CTU indices sorted by (col + 2*row, row)
cudaError_t launch_intra_frame(const FrameParams &fp, const uint8_t *src, uint8_t *rec, int16_t *levels, CuInfo *cu,
                               int *ticket, const int *order, cudaStream_t s);
cudaError_t launch_intra_in_p(const FrameParams &fp, const uint8_t *src, uint8_t *rec, int16_t *levels, CuInfo *cu,
                              int *ticket, const int *order, cudaStream_t s);

// decoder-side reconstruction (levels / modes / motion from the parser)
cudaError_t launch_inter_decode(const FrameParams &fp, const RefList &refs, uint8_t *rec, const int16_t *levels,
                                const CuInfo *cu, cudaStream_t s);
cudaError_t launch_intra_decode(const FrameParams &fp, uint8_t *rec, const int16_t *levels, const CuInfo *cu,
                                int *ticket, const int *order, cudaStream_t s);
// CABAC parse: one warp per substream.  data = unescaped slice data, bases[r] = offset of row r
// (bases[rows] = end).  status[0] = first error (0 ok), status[1] = largest |mv| component.
cudaError_t launch_parse(const FrameParams &fp, const uint8_t *data, const uint32_t *bases, CuInfo *cu, int16_t *levels,
                         uint8_t *sync_ctx, int *sync_flag, int *progress, int *status, cudaStream_t s);

// motion field of a parsed picture at 16x16 granularity ((w+15)/16 x (h+15)/16 entries), what later
// pictures' temporal motion vector candidates read (FrameParams::col_mvf)
cudaError_t launch_store_mvf(const FrameParams &fp, const CuInfo *cu, MvField *out, cudaStream_t s);

// cu_qp_delta (fp.ctu_qp != 0): per-CU luma QP into the cu map, per-CTU coded delta and the CU that
// codes it into fp.ctu_delta / fp.ctu_first.  After reconstruction (needs the cbf), before
// deblocking and binarisation.
cudaError_t launch_cu_qps(const FrameParams &fp, CuInfo *cu, cudaStream_t s);

// margins of a coded picture (w x h) whose source is src_w x src_h (even, within 7 samples of w x h): every plane's
// last source column / row repeated to the right / below (hevc_pad.cu).  No launch when the sizes are equal.
cudaError_t launch_pad_edges(uint8_t *pic, int w, int h, int src_w, int src_h, cudaStream_t s);

// variance adaptive quantisation (hevc_vaq.cu), two launches: per-CTU sample statistics of the source picture
// into `stats` (6 words per CTU), then ctu_qp[i] (staged as QP + kVaqBias) += the CTU's offset, clipped to 0..51
cudaError_t launch_vaq(const FrameParams &fp, const uint8_t *src, int strength, uint32_t *stats, uint8_t *ctu_qp, cudaStream_t s);

// in-loop deblocking, in place on `rec` (vertical edges of the whole picture, then horizontal)
cudaError_t launch_deblock(const FrameParams &fp, uint8_t *rec, const CuInfo *cu, cudaStream_t s);

// sample adaptive offset (hevc_sao.cu), one launch: deblocked picture `dbk` -> output picture `out`.
// encode: statistics against `src`, per-CTU decision into `params`, apply.  decode: apply `params`.
cudaError_t launch_sao_encode(const FrameParams &fp, const uint8_t *src, const uint8_t *dbk, uint8_t *out, SaoCtu *params,
                              cudaStream_t s);
cudaError_t launch_sao_decode(const FrameParams &fp, const uint8_t *dbk, uint8_t *out, const SaoCtu *params, cudaStream_t s);

// CABAC, two kernels meeting in the bin-record buffer `recs`
// (ctb_cols*ctb_rows*64*kRecUnitCap words): k_binarise (one warp per CU, whole picture in parallel)
// and the two entropy phases per CTU row / WPP substream (context resolution, then the range coder): as the
// two halves of one launch, the range coder one CTU behind (`fused`: lowest latency, what a stream with
// nothing in flight wants), or as two launches (a deep pipeline: no polling next to the prediction chain).
// rows[r*row_cap ..] receives the escaped bytes of row r, row_len[r] its length (0xffffffff on overflow).
cudaError_t launch_binarise(const FrameParams &fp, const CuInfo *cu, const int16_t *levels, uint32_t *recs, cudaStream_t s);
cudaError_t launch_arith(const FrameParams &fp, const CuInfo *cu, uint32_t *recs, uint8_t *rows, uint32_t row_cap,
                         uint32_t *row_len, uint8_t *sync_ctx, int *sync_flag, unsigned long long *bins, bool fused, cudaStream_t s);

// substreams -> one contiguous buffer + header {total, row_len[rows]} (dst/hdr may be mapped host memory)
cudaError_t launch_pack_rows(int rows, const uint8_t *src, uint32_t row_cap, const uint32_t *row_len, uint8_t *dst,
                             uint32_t dst_cap, uint32_t *hdr, cudaStream_t s);

}  // namespace b200
