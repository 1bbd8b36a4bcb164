// Parameter-set and slice-header parsing of the B200 HEVC decoder (host code, internal header).
//
// The parsers read the complete syntax of H.265 7.3.2.2 (SPS incl. profile_tier_level, short-term
// RPS list with inter-RPS prediction, VUI up to the timing info), 7.3.2.3 (PPS) and 7.3.6.1 (slice
// segment header) into plain structures; WHICH of it the CUDA decoder can reconstruct is decided
// separately (hevc_decoder.cu: supported()), so that "parsed" and "supported" do not get mixed up.
// This is what OpenHEVC does behind libOpenHevcDecode for the streams a Kvazaar peer sends
// (reference src/media/processing/openhevcfilter.cpp:145).
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <string>
#include <vector>

namespace b200 {

struct BitReader {
  const uint8_t *p;
  size_t n, pos = 0;       // pos in bits
  bool bad = false;
  BitReader(const uint8_t *d, size_t len) : p(d), n(len) {}
  uint32_t u(int bits)
  {
    uint32_t v = 0;
    for (int i = 0; i < bits; i++) {
      if (pos >= n * 8) { bad = true; return 0; }
      v = (v << 1) | ((p[pos >> 3] >> (7 - (pos & 7))) & 1);
      pos++;
    }
    return v;
  }
  void skip(int bits) { pos += (size_t)bits; if (pos > n * 8) bad = true; }
  uint32_t ue()
  {
    // values of 2^30 and more are refused: no syntax element this decoder reads comes near, and what callers add to
    // or multiply with a value (+ 8, * 2) then stays inside an int
    int z = 0;
    while (!bad && u(1) == 0 && z < 30) z++;
    if (z >= 30) { bad = true; return 0; }
    return z ? ((1u << z) - 1 + u(z)) : 0;
  }
  int32_t se()
  {
    uint32_t k = ue();
    return (k & 1) ? (int32_t)((k + 1) >> 1) : -(int32_t)(k >> 1);
  }
  void align() { pos = (pos + 7) & ~(size_t)7; }
};

// short-term reference picture set (7.4.8): delta POCs in decoding order of the lists
struct ShortTermRps {
  int num_neg = 0, num_pos = 0;
  int delta_poc[16] = {};      // [0, num_neg): S0 (negative, closest first), then num_pos entries of S1
  uint8_t used[16] = {};
  int num_delta() const { return num_neg + num_pos; }
};

// Scaling factors (7.3.4 scaling_list_data, 7.4.5, 8.6.4.2) the way the kernels read them: the factor of a
// coefficient at (x, y) of a 4 << sizeId block is m[sizeId][matrixId][y * 4 + x] for 4x4 blocks and
// m[sizeId][matrixId][(y >> (sizeId - 1)) * 8 + (x >> (sizeId - 1))] above, except the DC coefficient
// of 16x16 / 32x32 blocks, dc[sizeId - 2][matrixId].  matrixId: 0..2 intra Y / Cb / Cr, 3..5 inter.
struct ScalingTable {
  uint8_t m[4][6][64];
  uint8_t dc[2][6];
  uint8_t pad[4];
  void set_default();                               // Tables 7-5 / 7-6
};
static_assert(sizeof(ScalingTable) == 1552, "the kernels index this layout");

struct Sps {
  bool valid = false;
  int id = 0, chroma_format_idc = 1;
  int width = 0, height = 0;                        // coded size in luma samples
  int conf_left = 0, conf_right = 0, conf_top = 0, conf_bottom = 0;   // conformance window, luma samples
  int bit_depth_luma = 8, bit_depth_chroma = 8;
  int log2_max_poc = 8;
  int max_dec_pic_buffering = 1, max_num_reorder = 0;
  int log2_min_cb = 3, log2_ctb = 6, log2_min_tb = 2, log2_max_tb = 5;
  int max_tr_depth_inter = 0, max_tr_depth_intra = 0;
  int scaling_list = 0, amp = 0, sao = 0, pcm = 0;
  int scaling_list_data = 0;                        // sps_scaling_list_data_present_flag: `lists` holds them (else the defaults apply)
  ScalingTable lists;
  std::vector<ShortTermRps> rps;
  int long_term_refs = 0, num_lt_sps = 0, tmvp = 0, strong_intra_smoothing = 0;
  int fps_num = 0, fps_den = 0;                     // from the VUI timing info (0 = absent)
};

struct Pps {
  bool valid = false;
  int id = 0, sps_id = 0;
  int dependent_slices = 0, output_flag_present = 0, extra_slice_header_bits = 0;
  int sign_hiding = 0, cabac_init_present = 0;
  int num_ref_idx_l0_default = 1, num_ref_idx_l1_default = 1;
  int init_qp = 26;
  int constrained_intra = 0, transform_skip = 0;
  int qp_delta = 0, diff_cu_qp_delta_depth = 0;
  int cb_qp_offset = 0, cr_qp_offset = 0, slice_chroma_qp_offsets = 0;
  int weighted_pred = 0, weighted_bipred = 0, transquant_bypass = 0;
  int tiles = 0, wpp = 0;
  int tile_cols = 1, tile_rows = 1, uniform_spacing = 1, loop_filter_across_tiles = 1;
  std::vector<int> col_width, row_height;           // explicit sizes in CTBs when !uniform_spacing
  int loop_across_slices = 0;
  int deblock_ctrl = 0, deblock_override_enabled = 0, deblock_disabled = 0, beta_offset_div2 = 0, tc_offset_div2 = 0;
  int scaling_list = 0, lists_modification = 0, log2_parallel_merge_level = 2, slice_header_extension = 0;
  ScalingTable lists;                               // scaling_list != 0 (pps_scaling_list_data_present_flag): these replace the SPS's
};

struct SliceHeader {
  int first_slice_in_pic = 1, pps_id = 0, dependent = 0, segment_address = 0;
  int slice_type = 2;                               // 0 B, 1 P, 2 I
  int poc_lsb = 0;
  ShortTermRps rps;                                 // the set in force for this picture
  int tmvp = 0, sao_luma = 0, sao_chroma = 0;
  int num_ref_idx_l0 = 0;
  int cabac_init_flag = 0, collocated_ref_idx = 0, max_merge_cand = 5;
  int qp = 26, cb_qp_offset = 0, cr_qp_offset = 0;
  int deblock_disabled = 0, beta_offset_div2 = 0, tc_offset_div2 = 0, loop_across_slices = 0;
  std::vector<uint32_t> entry;                      // entry_point_offset_minus1 + 1 (escaped bytes)
  size_t data_offset = 0;                           // first byte of slice data inside the unescaped NAL payload
};

// All return false with `err` set when the syntax is malformed (or uses what cannot even be
// skipped); they do not judge decodability.
bool parse_sps_rbsp(const uint8_t *rbsp, size_t n, Sps &sps, std::string &err);
bool parse_pps_rbsp(const uint8_t *rbsp, size_t n, Pps &pps, std::string &err);
bool parse_slice_header_rbsp(const uint8_t *rbsp, size_t n, int nal_type, const Sps &sps, const Pps &pps,
                             SliceHeader &sh, std::string &err);

}  // namespace b200
