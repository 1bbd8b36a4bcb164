// Shared host/device definitions of the B200 HEVC codec (internal header).
//
// Data layout in HBM (per encoder / decoder instance), all picture-shaped so that every
// kernel addresses it with coalesced row segments:
//   frame planes   packed I420 (Y w*h, Cb, Cr), uint8
//   levels         quantised transform levels, int16, same I420 geometry: the level of the
//                  coefficient (u,v) of the TU whose top-left sample is (x0,y0) lives at
//                  plane[(y0+v)*stride + x0+u]
//   cu map         one 16-byte CuInfo per 8x8 luma unit, raster order, every unit of a CU
//                  carries the CU's values (what the entropy coder, the deblocking filter
//                  and the decoder's reconstruction all index by sample position)
//   substreams     one byte buffer per CTU row (WPP substream), already escaped
#pragma once
#include <stdint.h>

namespace b200 {

constexpr int kCtbLog2 = 6;
constexpr int kCtb = 64;
constexpr int kMaxMerge = 5;
constexpr int kCuOverheadBits = 3;
// Bin-record buffer: kRecUnitCap 32-bit words per 8x8 luma unit (CTU-major, z-order inside the
// CTU), so the record list of a CU starts at a fixed address and owns the capacity of all its
// units.  Worst case of an 8x8 CU: 6 sub-blocks x 75 records + CU header + 3 last positions < 640.
constexpr int kRecUnitCap = 640;
// Variance adaptive quantisation: the host stages picture QP + ROI offset of every CTU biased by
// kVaqBias (unclipped), k_vaq_qp adds its offset and clips to 0..51.
constexpr int kVaqBias = 76;

struct CuInfo {
  int16_t mvx, mvy;     // quarter-sample motion vector (inter).  Intra NxN CUs keep the modes of parts 1..3 here:
                        // mvx & 0xff, mvx >> 8, mvy & 0xff
  uint8_t log2_size;    // CU size 3..6 (0 = outside the picture / not yet decided)
  uint8_t pred_mode;    // 0 inter, 1 intra
  uint8_t intra_mode;   // 0..34 (part 0 of an NxN CU)
  uint8_t cbf;          // bit0 Y, bit1 Cb, bit2 Cr of the transform unit covering this 8x8 unit; bits 4-7: luma cbf
                        // of its four 4x4 transform blocks when tu_log2 == 2
  uint8_t skip;
  uint8_t merge_idx;    // 0xff = not merged
  uint8_t mvp_idx;
  uint8_t qp;           // luma QP of the CU when cu_qp_delta is enabled (FrameParams::ctu_qp != 0), else 0
  uint8_t ref_idx;      // index into reference picture list 0 (inter)
  uint8_t chroma_mode;  // intra: resolved chroma prediction mode 0..34 (mode 4 "derived" already replaced)
  uint8_t tu_log2;      // luma size of the transform unit covering this 8x8 unit: 2 (four 4x4 blocks) .. 5
  uint8_t flags;        // bit 0: intra NxN partition (8x8 CU, four 4x4 prediction blocks)
};
static_assert(sizeof(CuInfo) == 16, "CuInfo layout is part of the test ABI");

// Motion of one 16x16 block of a decoded picture (the 8x8 unit at its top-left corner), kept for the
// temporal motion vector prediction of later pictures (8.5.3.2.8; the standard compresses to 16x16).
struct MvField {
  int16_t mvx, mvy;
  int16_t poc_diff;     // POC(picture) - POC(the reference picture the vector points to)
  uint8_t inter, pad;
};
static_assert(sizeof(MvField) == 8, "MvField layout");

// Reference picture list 0 of a picture (packed I420 pictures in HBM); the encoder has one entry.
struct RefList {
  const uint8_t *pic[16];
  int n;
};

// SAO parameters of one CTU (7.4.9.3): [0] luma, [1] chroma (type and class shared by Cb and Cr)
struct SaoCtu {
  uint8_t type[2];        // 0 off, 1 band offset, 2 edge offset
  uint8_t eo_class[2];
  uint8_t band_pos[3];
  int8_t offset[3][4];    // signed SaoOffsetVal of categories 1..4 / the four bands
  uint8_t pad;
};
static_assert(sizeof(SaoCtu) == 20, "SaoCtu layout");

// CABAC context layout (one flat table per substream)
enum CtxOffset {
  CTX_SPLIT_CU = 0, CTX_SKIP = 3, CTX_MERGE_FLAG = 6, CTX_MERGE_IDX = 7, CTX_PART_MODE = 8,
  CTX_PRED_MODE = 12, CTX_PREV_INTRA_LUMA = 13, CTX_INTRA_CHROMA = 14, CTX_MVD_GT0 = 15,
  CTX_MVD_GT1 = 16, CTX_MVP_IDX = 17, CTX_RQT_ROOT_CBF = 18, CTX_SPLIT_TRANSFORM = 19,
  CTX_CBF_LUMA = 22, CTX_CBF_CHROMA = 24, CTX_LAST_X = 28, CTX_LAST_Y = 46, CTX_CSBF = 64,
  CTX_SIG = 68, CTX_GT1 = 110, CTX_GT2 = 134, CTX_CU_QP_DELTA = 140, CTX_SAO_MERGE = 142, CTX_SAO_TYPE = 143,
  CTX_REF_IDX = 144, CTX_COUNT = 146
};

struct FrameParams {
  int w, h;             // luma size, multiples of 8
  int w8, h8;           // size in 8x8 units
  int ctb_cols, ctb_rows;
  int qp, qp_c;         // luma / chroma QP
  int lambda_q4;        // 16 * sqrt(lambda) for SAD-domain costs
  int search_range;
  int is_idr;           // quantiser rounding offset and CABAC init type follow the slice type
  int deblock;
  // cu_qp_delta (ROI) state, all null when the PPS flag is off.  ctu_qp: luma QP each CTU quantises
  // with (encoder: slice QP + ROI offset; decoder: written by the parser).  ctu_delta / ctu_first:
  // CuQpDeltaVal coded in the CTU and the z-index (8x8 units) of the CU that codes it (64 = none).
  uint8_t *ctu_qp;
  int8_t *ctu_delta;
  uint8_t *ctu_first;
  // Tile mode: this picture is one tile of a larger one, coded as a picture of its own.  mv_edges bit
  // 0 / 1 / 2 / 3: the left / right / top / bottom edge is an interior tile edge, so no reference
  // sample beyond it may be used (the neighbouring tile is there, not edge padding).  more_tiles:
  // tiles follow in the slice, the last CTU does not end the slice segment.  no_wpp: one substream
  // for the whole picture (entropy_coding_sync off; HEVC Main allows tiles or WPP, not both).
  int mv_edges, more_tiles, no_wpp;
  // SAO: per-CTU parameters (raster) -- written by k_sao_ctu (encoder) or the parser (decoder), read by
  // the binariser / the decoder's apply pass; null when SAO is off.  sao_flags: bit 0 slice_sao_luma_flag,
  // bit 1 slice_sao_chroma_flag, bit 2 (encoder) use the merge flags where parameters repeat.
  SaoCtu *sao;
  int sao_flags;
  // Intra CUs in P pictures.  intra_in_p (encoder): the motion search also tries 16x16 intra CUs.
  // any_intra: set to 1 by whoever finds an intra CU in a P picture (motion search / parser), so that
  // the intra pass over a P picture without any returns at once; ctu_done: one flag per CTU for the
  // intra pass's wavefront (hevc_intra.cu).
  int intra_in_p;
  int subme_satd;       // encoder, P pictures: the fractional motion refinement compares SATD instead of SAD (row K2)
  int intra_satd;       // encoder, I pictures: the 35-mode search compares Hadamard SATD instead of SAD (row K2)
  int *any_intra;
  int *ctu_done;
  // Two-level motion search (encoder): me_coarse > 0 = range of the coarse level in coarse samples (a
  // multiple of 4); src_q / ref_q = quarter-resolution luma of the source and of the reference.
  int me_coarse;
  const uint8_t *src_q, *ref_q;
  // Reference picture list 0 as the entropy stage and the deblocking filter see it.  n_refs =
  // num_ref_idx_l0_active (1 in the encoder); ref_dist[i] = POC(current) - POC(RefPicList0[i]): two
  // indices name the same picture iff their distances are equal, and vector scaling uses the ratio.
  // max_merge = MaxNumMergeCand.  col_mvf (decoder, slice_temporal_mvp_enabled_flag): motion field of
  // the collocated picture, ((w + 15) / 16) entries per row; null = no temporal candidates.
  int n_refs, max_merge;
  // CABAC initialisation type (9.3.2.2): 0 I slices, 1 P slices, 2 P slices with cabac_init_flag; decoder:
  // max_transform_hierarchy_depth_inter / _intra of the SPS
  int init_type, tr_depth_inter, tr_depth_intra;
  int16_t ref_dist[16];
  const MvField *col_mvf;
  // Decoder: syntax of foreign (Kvazaar-family) streams that this encoder does not produce -- all zero in
  // the encoder.  sign_hiding / strong_intra: the PPS / SPS flags.  cb / cr_qp_offset: PPS + slice chroma QP
  // offsets (8.6.1); *_pps: the PPS part alone, what chroma deblocking uses (cQpPicOffset, 8.7.2.5.5).
  // beta / tc_offset_div2: deblocking offsets in force for the slice.
  int sign_hiding, strong_intra;
  int cb_qp_offset, cr_qp_offset, cb_qp_offset_pps, cr_qp_offset_pps;
  int beta_offset_div2, tc_offset_div2;
  // Scaling lists (scaling_list_enabled_flag): the ScalingTable (hevc_headers.h) in force, in device memory;
  // null = flat (every factor 16).  Encoder: the default lists ("scaling-list default").
  const uint8_t *scaling;
  // optional work counters of the motion search (profiling): [0] CTUs, [1] 32x32 quadrants whose second
  // centre set was searched, [2] 16x16 intra mode searches, [3] intra CUs chosen (16x16)
  unsigned long long *me_stats;
};

}  // namespace b200
