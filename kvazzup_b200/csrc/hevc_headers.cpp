// H.265 parameter-set and slice-header syntax (7.3.2.2, 7.3.2.3, 7.3.6.1, 7.3.7, E.2.1) -> structures.
// Host code of the B200 decoder; see hevc_headers.h.
#include "hevc_headers.h"

#include <string.h>

#include <algorithm>

namespace b200 {

namespace {

// profile_tier_level(1, max_sub_layers_minus1), 7.3.3: nothing in it changes how a picture decodes
void skip_profile_tier_level(BitReader &b, int max_sub_layers_minus1)
{
  b.skip(2 + 1 + 5); b.skip(32); b.skip(4); b.skip(43); b.skip(1);
  b.skip(8);                                       // general_level_idc
  uint8_t prof[8] = {}, lev[8] = {};
  for (int i = 0; i < max_sub_layers_minus1; i++) { prof[i] = (uint8_t)b.u(1); lev[i] = (uint8_t)b.u(1); }
  if (max_sub_layers_minus1 > 0)
    for (int i = max_sub_layers_minus1; i < 8; i++) b.skip(2);
  for (int i = 0; i < max_sub_layers_minus1; i++) {
    if (prof[i]) b.skip(88);
    if (lev[i]) b.skip(8);
  }
}

// st_ref_pic_set(idx), 7.3.7 + the derivation of 7.4.8.  `sets` holds the idx sets before this one
// (for the slice header's own set: all sets of the SPS, idx == their count).
bool parse_st_rps(BitReader &b, int idx, int num_sps_sets, const std::vector<ShortTermRps> &sets, ShortTermRps &out)
{
  out = ShortTermRps();
  int inter = 0;
  if (idx != 0) inter = (int)b.u(1);
  if (inter) {
    int delta_idx = 1;
    if (idx == num_sps_sets) delta_idx = (int)b.ue() + 1;
    if (delta_idx > idx) return false;
    const ShortTermRps &ref = sets[idx - delta_idx];
    const int sign = (int)b.u(1);
    const int delta_rps = (1 - 2 * sign) * ((int)b.ue() + 1);
    const int nref = ref.num_delta();
    uint8_t used[17], use_delta[17];
    for (int j = 0; j <= nref; j++) {
      used[j] = (uint8_t)b.u(1);
      use_delta[j] = 1;
      if (!used[j]) use_delta[j] = (uint8_t)b.u(1);
    }
    int s0[16], s1[16], n0 = 0, n1 = 0;
    uint8_t u0[16], u1[16];
    auto ref_s0 = [&](int j) { return ref.delta_poc[j]; };
    auto ref_s1 = [&](int j) { return ref.delta_poc[ref.num_neg + j]; };
    for (int j = ref.num_pos - 1; j >= 0; j--) {
      const int d = ref_s1(j) + delta_rps;
      if (d < 0 && use_delta[ref.num_neg + j] && n0 < 16) { s0[n0] = d; u0[n0++] = used[ref.num_neg + j]; }
    }
    if (delta_rps < 0 && use_delta[nref] && n0 < 16) { s0[n0] = delta_rps; u0[n0++] = used[nref]; }
    for (int j = 0; j < ref.num_neg; j++) {
      const int d = ref_s0(j) + delta_rps;
      if (d < 0 && use_delta[j] && n0 < 16) { s0[n0] = d; u0[n0++] = used[j]; }
    }
    for (int j = ref.num_neg - 1; j >= 0; j--) {
      const int d = ref_s0(j) + delta_rps;
      if (d > 0 && use_delta[j] && n1 < 16) { s1[n1] = d; u1[n1++] = used[j]; }
    }
    if (delta_rps > 0 && use_delta[nref] && n1 < 16) { s1[n1] = delta_rps; u1[n1++] = used[nref]; }
    for (int j = 0; j < ref.num_pos; j++) {
      const int d = ref_s1(j) + delta_rps;
      if (d > 0 && use_delta[ref.num_neg + j] && n1 < 16) { s1[n1] = d; u1[n1++] = used[ref.num_neg + j]; }
    }
    if (n0 + n1 > 16) return false;
    out.num_neg = n0; out.num_pos = n1;
    for (int i = 0; i < n0; i++) { out.delta_poc[i] = s0[i]; out.used[i] = u0[i]; }
    for (int i = 0; i < n1; i++) { out.delta_poc[n0 + i] = s1[i]; out.used[n0 + i] = u1[i]; }
  } else {
    const uint32_t neg = b.ue(), pos = b.ue();
    if (neg + pos > 16) return false;
    out.num_neg = (int)neg; out.num_pos = (int)pos;
    int poc = 0;
    for (uint32_t i = 0; i < neg; i++) { poc -= (int)b.ue() + 1; out.delta_poc[i] = poc; out.used[i] = (uint8_t)b.u(1); }
    poc = 0;
    for (uint32_t i = 0; i < pos; i++) { poc += (int)b.ue() + 1; out.delta_poc[neg + i] = poc; out.used[neg + i] = (uint8_t)b.u(1); }
  }
  return !b.bad;
}

// vui_parameters(), E.2.1, up to the timing info (nothing after it is used; HRD parameters in front
// of the bitstream restriction are not parsed)
void parse_vui(BitReader &b, Sps &sps)
{
  if (b.u(1)) {                                   // aspect_ratio_info_present_flag
    if (b.u(8) == 255) b.skip(32);
  }
  if (b.u(1)) b.skip(1);                          // overscan
  if (b.u(1)) {                                   // video_signal_type_present_flag
    b.skip(3 + 1);
    if (b.u(1)) b.skip(24);
  }
  if (b.u(1)) { b.ue(); b.ue(); }                 // chroma_loc_info
  b.skip(3);                                      // neutral_chroma_indication, field_seq, frame_field_info_present
  if (b.u(1)) { b.ue(); b.ue(); b.ue(); b.ue(); } // default display window
  if (b.u(1)) {                                   // vui_timing_info_present_flag
    const uint32_t units = b.u(32), scale = b.u(32);
    if (!b.bad && units > 0 && scale > 0 && units < (1u << 31) && scale < (1u << 31)) { sps.fps_num = (int)scale; sps.fps_den = (int)units; }
  }
}

}  // namespace

namespace {

// Table 7-6 in up-right diagonal order (shared by the colour components and by every size from 8x8 up)
const uint8_t kDefaultIntra[64] = {
    16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 17, 16, 17, 16, 17, 18, 17, 18, 18, 17, 18, 21, 19, 20, 21, 20, 19, 21, 24, 22, 22, 24,
    24, 22, 22, 24, 25, 25, 27, 30, 27, 25, 25, 29, 31, 35, 35, 31, 29, 36, 41, 44, 41, 36, 47, 54, 54, 47, 65, 70, 65, 88, 88, 115};
const uint8_t kDefaultInter[64] = {
    16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 17, 17, 17, 17, 17, 18, 18, 18, 18, 18, 18, 20, 20, 20, 20, 20, 20, 20, 24, 24, 24, 24,
    24, 24, 24, 24, 25, 25, 25, 25, 25, 25, 25, 28, 28, 28, 28, 28, 28, 33, 33, 33, 33, 33, 41, 41, 41, 41, 54, 54, 54, 71, 71, 91};

// raster position of entry i of the up-right diagonal scan (6.5.3) of an n x n array
void diag_order(int n, uint8_t *pos)
{
  int i = 0, x = 0, y = 0;
  while (i < n * n) {
    while (y >= 0) {
      if (x < n && y < n) pos[i++] = (uint8_t)(y * n + x);
      y--; x++;
    }
    y = x; x = 0;
  }
}

void store_list(ScalingTable &t, int sid, int mid, const uint8_t *list, int dc)
{
  uint8_t pos[64];
  const int n = sid == 0 ? 4 : 8;
  diag_order(n, pos);
  memset(t.m[sid][mid], 16, 64);
  for (int i = 0; i < n * n; i++) t.m[sid][mid][pos[i]] = list[i];
  if (sid >= 2) t.dc[sid - 2][mid] = (uint8_t)dc;
}

void store_default(ScalingTable &t, int sid, int mid)
{
  uint8_t flat[16];
  memset(flat, 16, sizeof(flat));
  store_list(t, sid, mid, sid == 0 ? flat : (mid < 3 ? kDefaultIntra : kDefaultInter), 16);
}

// scaling_list_data() (7.3.4).  32x32 lists exist for luma only in 4:2:0 (matrixId 0 and 3).
bool parse_scaling_list_data(BitReader &b, ScalingTable &t)
{
  t.set_default();
  for (int sid = 0; sid < 4; sid++)
    for (int mid = 0; mid < 6; mid += sid == 3 ? 3 : 1) {
      if (!b.u(1)) {                                  // scaling_list_pred_mode_flag = 0: copy
        uint32_t delta = b.ue();
        if (delta > 5) return false;
        delta *= sid == 3 ? 3 : 1;
        if (delta > (uint32_t)mid) return false;
        if (delta == 0) { store_default(t, sid, mid); continue; }
        memcpy(t.m[sid][mid], t.m[sid][mid - delta], 64);
        if (sid >= 2) t.dc[sid - 2][mid] = t.dc[sid - 2][mid - delta];
        continue;
      }
      uint8_t list[64];
      const int cnt = sid == 0 ? 16 : 64;
      int next = 8, dc = 16;
      if (sid > 1) {
        const int d = b.se();
        if (d < -7 || d > 247) return false;
        next = dc = d + 8;
      }
      for (int i = 0; i < cnt; i++) {
        const int d = b.se();
        if (d < -128 || d > 127) return false;
        next = (next + d + 256) & 255;
        if (next == 0) return false;                  // scaling factors are 1..255
        list[i] = (uint8_t)next;
      }
      store_list(t, sid, mid, list, dc);
    }
  return !b.bad;
}

}  // namespace

void ScalingTable::set_default()
{
  memset(this, 0, sizeof(*this));
  for (int sid = 0; sid < 4; sid++)
    for (int mid = 0; mid < 6; mid++) store_default(*this, sid, mid);
}

bool parse_sps_rbsp(const uint8_t *rbsp, size_t n, Sps &sps, std::string &err)
{
  BitReader b(rbsp, n);
  sps = Sps();
  b.skip(4);
  const int max_sub = (int)b.u(3);
  b.skip(1);
  skip_profile_tier_level(b, max_sub);
  sps.id = (int)b.ue();
  sps.chroma_format_idc = (int)b.ue();
  if (sps.chroma_format_idc == 3) b.skip(1);
  sps.width = (int)b.ue(); sps.height = (int)b.ue();
  if (b.u(1)) {
    const int sw = sps.chroma_format_idc == 1 || sps.chroma_format_idc == 2 ? 2 : 1, shh = sps.chroma_format_idc == 1 ? 2 : 1;
    // offsets come from the network: bound them before they are multiplied (the window is checked against the picture later)
    auto off = [&]() { const uint32_t v = b.ue(); if (v > 8192) { b.bad = true; return 0; } return (int)v; };
    sps.conf_left = sw * off(); sps.conf_right = sw * off();
    sps.conf_top = shh * off(); sps.conf_bottom = shh * off();
  }
  sps.bit_depth_luma = 8 + (int)b.ue(); sps.bit_depth_chroma = 8 + (int)b.ue();
  sps.log2_max_poc = 4 + (int)b.ue();
  const int ordering = (int)b.u(1);
  for (int i = ordering ? 0 : max_sub; i <= max_sub; i++) {
    sps.max_dec_pic_buffering = (int)b.ue() + 1; sps.max_num_reorder = (int)b.ue(); b.ue();
  }
  sps.log2_min_cb = 3 + (int)b.ue();
  sps.log2_ctb = sps.log2_min_cb + (int)b.ue();
  sps.log2_min_tb = 2 + (int)b.ue();
  sps.log2_max_tb = sps.log2_min_tb + (int)b.ue();
  sps.max_tr_depth_inter = (int)b.ue(); sps.max_tr_depth_intra = (int)b.ue();
  sps.scaling_list = (int)b.u(1);
  if (sps.scaling_list) {
    sps.scaling_list_data = (int)b.u(1);
    if (sps.scaling_list_data && !parse_scaling_list_data(b, sps.lists)) { err = "malformed SPS (scaling list data)"; return false; }
  }
  sps.amp = (int)b.u(1); sps.sao = (int)b.u(1); sps.pcm = (int)b.u(1);
  if (sps.pcm) { b.skip(8); b.ue(); b.ue(); b.skip(1); }
  const uint32_t num_rps = b.ue();
  if (b.bad || num_rps > 64) { err = "malformed SPS (short-term RPS count)"; return false; }
  sps.rps.resize(num_rps);
  for (uint32_t i = 0; i < num_rps; i++)
    if (!parse_st_rps(b, (int)i, (int)num_rps, sps.rps, sps.rps[i])) { err = "malformed SPS (short-term RPS)"; return false; }
  sps.long_term_refs = (int)b.u(1);
  if (sps.long_term_refs) {
    const uint32_t nlt = b.ue();
    if (nlt > 32) { err = "malformed SPS (long-term pictures)"; return false; }
    sps.num_lt_sps = (int)nlt;
    for (uint32_t i = 0; i < nlt; i++) { b.skip(sps.log2_max_poc); b.skip(1); }
  }
  sps.tmvp = (int)b.u(1);
  sps.strong_intra_smoothing = (int)b.u(1);
  if (b.bad) { err = "malformed SPS"; return false; }
  if (b.u(1)) parse_vui(b, sps);                  // a VUI cut short only loses the (optional) timing
  // limits: sizes come from the network (level 6.2 tops out at 8192x4320 luma samples)
  if (sps.width <= 0 || sps.height <= 0 || sps.width > 8192 || sps.height > 8192 || (long long)sps.width * sps.height > 8192LL * 4320) {
    err = "SPS picture size out of range"; return false;
  }
  if (sps.log2_max_poc > 16 || sps.log2_ctb > 6 || sps.log2_ctb < 4 || sps.log2_max_tb > 5 || sps.log2_max_tb > sps.log2_ctb ||
      sps.log2_min_cb > sps.log2_ctb || sps.log2_min_tb >= sps.log2_min_cb) {
    err = "SPS block sizes / POC length out of range"; return false;
  }
  sps.valid = true;
  return true;
}

bool parse_pps_rbsp(const uint8_t *rbsp, size_t n, Pps &pps, std::string &err)
{
  BitReader b(rbsp, n);
  pps = Pps();
  pps.id = (int)b.ue(); pps.sps_id = (int)b.ue();
  pps.dependent_slices = (int)b.u(1); pps.output_flag_present = (int)b.u(1); pps.extra_slice_header_bits = (int)b.u(3);
  pps.sign_hiding = (int)b.u(1); pps.cabac_init_present = (int)b.u(1);
  pps.num_ref_idx_l0_default = (int)b.ue() + 1; pps.num_ref_idx_l1_default = (int)b.ue() + 1;
  pps.init_qp = 26 + b.se();
  pps.constrained_intra = (int)b.u(1); pps.transform_skip = (int)b.u(1);
  pps.qp_delta = (int)b.u(1);
  if (pps.qp_delta) pps.diff_cu_qp_delta_depth = (int)b.ue();
  pps.cb_qp_offset = b.se(); pps.cr_qp_offset = b.se();
  pps.slice_chroma_qp_offsets = (int)b.u(1);
  pps.weighted_pred = (int)b.u(1); pps.weighted_bipred = (int)b.u(1);
  pps.transquant_bypass = (int)b.u(1);
  pps.tiles = (int)b.u(1); pps.wpp = (int)b.u(1);
  if (pps.tiles) {
    pps.tile_cols = (int)b.ue() + 1; pps.tile_rows = (int)b.ue() + 1;
    if (pps.tile_cols > 64 || pps.tile_rows > 64) { err = "malformed PPS (tile counts)"; return false; }
    pps.uniform_spacing = (int)b.u(1);
    if (!pps.uniform_spacing) {
      for (int i = 0; i < pps.tile_cols - 1; i++) pps.col_width.push_back((int)b.ue() + 1);
      for (int i = 0; i < pps.tile_rows - 1; i++) pps.row_height.push_back((int)b.ue() + 1);
    }
    pps.loop_filter_across_tiles = (int)b.u(1);
  }
  pps.loop_across_slices = (int)b.u(1);
  pps.deblock_ctrl = (int)b.u(1);
  if (pps.deblock_ctrl) {
    pps.deblock_override_enabled = (int)b.u(1);
    pps.deblock_disabled = (int)b.u(1);
    if (!pps.deblock_disabled) { pps.beta_offset_div2 = b.se(); pps.tc_offset_div2 = b.se(); }
  }
  pps.scaling_list = (int)b.u(1);
  if (pps.scaling_list && !parse_scaling_list_data(b, pps.lists)) { err = "malformed PPS (scaling list data)"; return false; }
  pps.lists_modification = (int)b.u(1);
  pps.log2_parallel_merge_level = 2 + (int)b.ue();
  pps.slice_header_extension = (int)b.u(1);
  if (b.bad || pps.id > 63 || pps.sps_id > 15 || pps.init_qp < 0 || pps.init_qp > 51 || pps.num_ref_idx_l0_default > 15) {
    err = "malformed PPS"; return false;
  }
  pps.valid = true;
  return true;
}

bool parse_slice_header_rbsp(const uint8_t *rbsp, size_t n, int nal_type, const Sps &sps, const Pps &pps,
                             SliceHeader &sh, std::string &err)
{
  BitReader b(rbsp, n);
  sh = SliceHeader();
  const bool irap = nal_type >= 16 && nal_type <= 23;
  const bool idr = nal_type == 19 || nal_type == 20;
  sh.first_slice_in_pic = (int)b.u(1);
  if (irap) b.skip(1);
  sh.pps_id = (int)b.ue();
  if (!sh.first_slice_in_pic) {
    if (pps.dependent_slices) sh.dependent = (int)b.u(1);
    const int ctbs = ((sps.width + (1 << sps.log2_ctb) - 1) >> sps.log2_ctb) * ((sps.height + (1 << sps.log2_ctb) - 1) >> sps.log2_ctb);
    int bits = 0;
    while ((1 << bits) < ctbs) bits++;
    sh.segment_address = (int)b.u(bits);
  }
  if (sh.dependent) { err = "dependent slice segment"; return false; }
  b.skip(pps.extra_slice_header_bits);
  sh.slice_type = (int)b.ue();
  if (sh.slice_type > 2) { err = "malformed slice header (slice_type)"; return false; }
  if (pps.output_flag_present) b.skip(1);
  if (!idr) {
    sh.poc_lsb = (int)b.u(sps.log2_max_poc);
    if (!b.u(1)) {
      if (!parse_st_rps(b, (int)sps.rps.size(), (int)sps.rps.size(), sps.rps, sh.rps)) { err = "malformed slice header (short-term RPS)"; return false; }
    } else {
      if (sps.rps.empty()) { err = "slice refers to a missing RPS"; return false; }
      int bits = 0;
      while ((1u << bits) < sps.rps.size()) bits++;
      const uint32_t idx = bits ? b.u(bits) : 0;
      if (idx >= sps.rps.size()) { err = "malformed slice header (RPS index)"; return false; }
      sh.rps = sps.rps[idx];
    }
    if (sps.long_term_refs) {
      // no long-term pictures are ever marked by the encoders this decoder serves; a slice that
      // signals any is rejected by the caller (num_long_term_* must be 0)
      uint32_t lt = 0;
      if (sps.num_lt_sps > 0) lt += b.ue();         // num_long_term_sps
      lt += b.ue();                                 // num_long_term_pics
      if (lt != 0) { err = "long-term reference pictures"; return false; }
    }
    if (sps.tmvp) sh.tmvp = (int)b.u(1);
  }
  if (sps.sao) { sh.sao_luma = (int)b.u(1); sh.sao_chroma = (int)b.u(1); }
  sh.num_ref_idx_l0 = 0;
  if (sh.slice_type != 2) {
    sh.num_ref_idx_l0 = pps.num_ref_idx_l0_default;
    if (b.u(1)) {
      sh.num_ref_idx_l0 = (int)b.ue() + 1;
      if (sh.slice_type == 0) b.ue();
    }
    int total_curr = 0;
    for (int i = 0; i < sh.rps.num_delta(); i++) total_curr += sh.rps.used[i];
    if (pps.lists_modification && total_curr > 1) { err = "reference picture list modification"; return false; }
    if (sh.slice_type == 0) b.skip(1);              // mvd_l1_zero_flag
    if (pps.cabac_init_present) sh.cabac_init_flag = (int)b.u(1);
    if (sh.tmvp) {
      int from_l0 = 1;
      if (sh.slice_type == 0) from_l0 = (int)b.u(1);
      if (from_l0 && sh.num_ref_idx_l0 > 1) sh.collocated_ref_idx = (int)b.ue();
    }
    if ((pps.weighted_pred && sh.slice_type == 1) || (pps.weighted_bipred && sh.slice_type == 0)) { err = "weighted prediction"; return false; }
    sh.max_merge_cand = 5 - (int)b.ue();
    if (sh.max_merge_cand < 1 || sh.max_merge_cand > 5 || sh.num_ref_idx_l0 > 15) { err = "malformed slice header (merge candidates / reference count)"; return false; }
  }
  sh.qp = pps.init_qp + b.se();
  if (pps.slice_chroma_qp_offsets) { sh.cb_qp_offset = b.se(); sh.cr_qp_offset = b.se(); }
  sh.deblock_disabled = pps.deblock_disabled; sh.beta_offset_div2 = pps.beta_offset_div2; sh.tc_offset_div2 = pps.tc_offset_div2;
  int override_flag = 0;
  if (pps.deblock_override_enabled) override_flag = (int)b.u(1);
  if (override_flag) {
    sh.deblock_disabled = (int)b.u(1);
    if (!sh.deblock_disabled) { sh.beta_offset_div2 = b.se(); sh.tc_offset_div2 = b.se(); }
  }
  sh.loop_across_slices = pps.loop_across_slices;
  if (pps.loop_across_slices && (sh.sao_luma || sh.sao_chroma || !sh.deblock_disabled)) sh.loop_across_slices = (int)b.u(1);
  if (pps.tiles || pps.wpp) {
    const uint32_t n_entry = b.ue();
    if (b.bad || n_entry > 8192) { err = "implausible number of entry points"; return false; }
    sh.entry.resize(n_entry);
    if (n_entry > 0) {
      const int len = (int)b.ue() + 1;
      if (len > 32) { err = "malformed slice header (entry point length)"; return false; }
      for (uint32_t i = 0; i < n_entry; i++) sh.entry[i] = b.u(len) + 1;
    }
  }
  if (pps.slice_header_extension) {
    const uint32_t len = b.ue();
    if (len > 256) { err = "malformed slice header (extension length)"; return false; }
    b.skip(8 * (int)len);
  }
  if (!b.u(1)) { err = "malformed slice header (alignment bit)"; return false; }
  b.align();
  if (b.bad || sh.qp < 0 || sh.qp > 51) { err = "malformed slice header"; return false; }
  sh.data_offset = b.pos >> 3;
  return true;
}

}  // namespace b200
