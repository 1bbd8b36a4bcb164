// Host side of the B200 HEVC decoder and the libOpenHevc* boundary (include/b200_openhevc.h):
// NAL splitting, parameter-set and slice-header parsing, substream extraction, kernel sequencing
// (CABAC parse -> reconstruction -> deblocking), picture output.  Replaces what
// reference src/media/processing/openhevcfilter.cpp:38-56,145,195-199 calls in OpenHEVC.
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <deque>
#include <memory>
#include <string>
#include <vector>

#include "../../include/b200_openhevc.h"
#include "hevc_headers.h"
#include "hevc_kernels.h"
#include "runtime.h"

namespace b200 {

namespace {

std::vector<uint8_t> unescape(const uint8_t *p, size_t n)
{
  std::vector<uint8_t> out;
  out.reserve(n);
  int zeros = 0;
  for (size_t i = 0; i < n; i++) {
    if (zeros >= 2 && p[i] == 3) { zeros = 0; continue; }
    out.push_back(p[i]);
    zeros = p[i] == 0 ? zeros + 1 : 0;
  }
  return out;
}

const uint8_t kChromaQpD[58] = {
  0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29,
  29, 30, 31, 32, 33, 33, 34, 34, 35, 35, 36, 36, 37, 37, 38, 39, 40, 41, 42, 43, 44, 45, 46, 47, 48, 49, 50, 51};

}  // namespace

// A tile (of a uniform grid of tile columns x tile rows) is decoded as a picture of its own ("strip", see
// hevc_tiles.cu for why that is exact when motion stays inside the tile and tiles are not loop-filtered
// across); a picture without tiles is a single strip covering everything.  Per decoder: the strip's geometry, reconstruction
// ping-pong and reconstruction stream.
struct StripGeom {
  int x0 = 0, y0 = 0, wd = 0, ht = 0;
  FrameParams fp{};                       // strip-sized
  size_t bytes = 0;                       // packed I420 of the strip
  // decoded picture buffer of the strip: kPool pictures (index shared by all strips of a picture), each
  // with its 16x16 motion field for temporal MV prediction and the event that says the field is written
  uint8_t *d_pic[8] = {};
  MvField *d_mvf[8] = {};
  cudaEvent_t ev_mvf[8] = {};
  uint8_t *d_dbk = nullptr;               // deblocked picture of a slice with SAO (SAO reads it, writes d_rec)
  int *d_order = nullptr;
  cudaStream_t stream = nullptr;          // reconstruction chain of this strip
  cudaEvent_t ev_done = nullptr;
};

// Per picture in flight and strip: what the CABAC parse of the strip reads and writes.
struct StripBufs {
  cudaStream_t stream = nullptr;
  cudaEvent_t ev_parsed = nullptr;
  uint8_t *d_small = nullptr, *d_ctu_qp = nullptr;
  SaoCtu *d_sao = nullptr;                // per-CTU SAO parameters from the parser
  int *d_ctu_done = nullptr;              // intra wavefront: one flag per CTU, then "the P picture has intra CUs"
  CuInfo *d_cu = nullptr;
  int16_t *d_levels = nullptr;
  uint32_t *h_bases = nullptr;
  int *h_status = nullptr;
  FrameParams fp{};
};

// One picture in flight.  Parsing needs nothing from other pictures (no TMVP), so the parses of up
// to `frame_delay + 1` pictures run concurrently on their own streams; only reconstruction is
// chained from picture to picture.
struct DecSlot {
  cudaStream_t stream = nullptr;          // slice data upload
  cudaEvent_t ev_uploaded = nullptr;
  uint8_t *d_data = nullptr, *h_data = nullptr, *h_out = nullptr;
  uint8_t *d_scaling = nullptr, *h_scaling = nullptr;   // the ScalingTable this picture dequantises with (scaling lists on)
  std::vector<StripBufs> strips;
  int64_t pts = 0;
  int cur_idx = 0;                        // pool entry this picture is reconstructed into
  int n_refs = 0, ref_pool[16] = {};      // reference picture list 0 as pool entries
};

struct Decoder {
  Sps sps_tab[16];                        // parameter sets by id (7.4.2.4.2: a slice activates its PPS, the PPS its SPS)
  Pps pps_tab[64];
  Sps sps;                                // the active ones
  Pps pps;
  bool vps_seen = false, started = false;
  int prev_poc = 0, prev_poc_lsb = 0, prev_poc_msb = 0;   // POC of the last decoded picture (8.3.1)
  FrameParams fp{};                       // whole picture
  size_t frame_bytes = 0;
  cudaStream_t stream = nullptr;          // output assembly / copy
  cudaEvent_t ev_out = nullptr;           // the picture is in host memory (what libOpenHevcDecode waits for)
  std::vector<StripGeom> geom;
  uint8_t *d_full = nullptr;              // whole picture on the device when there are several strips
  // Conformance window (7.4.3.2.1): the decoder works on the coded size and hands out the window.
  int crop_l = 0, crop_t = 0, out_w = 0, out_h = 0;   // window origin and size, luma samples
  size_t out_bytes = 0;
  uint8_t *d_crop = nullptr;              // the window as a packed picture (only when it differs from the coded size)
  bool cropped() const { return out_w != fp.w || out_h != fp.h; }
  int conf_w = 0, conf_h = 0, conf_tiles = 0, conf_tile_cols = 0, conf_tile_rows = 0;
  std::vector<DecSlot> slots;
  std::deque<int> pending;                // slots whose parse was launched, oldest first
  size_t data_cap = 0, small_bytes = 0, off_flag = 0, off_prog = 0, off_ticket = 0, off_status = 0, off_bases = 0, off_ctx = 0;
  int frame_delay = 0;                    // pictures held back (OpenHEVC frame threads - 1)
  int next_slot = 0, out_slot = -1;
  bool host_output = true;                // false: pictures stay on the GPU (b200_dec_output_dev)
  const uint8_t *d_out = nullptr;         // device copy of the last output picture
  // Decoded picture buffer bookkeeping (8.3.2): which pool entries hold a picture that may still be
  // referenced, and its POC.  Managed in decoding order when a slice is submitted.
  static constexpr int kPool = 8;
  struct DpbEntry { int poc = 0; bool valid = false; } dpb[kPool];
  int pictures = 0;
  bool have_submitted = false;            // a picture whose POC is prev_poc went into the pipeline
  int missing_refs = 0;                   // P pictures whose reference (POC - 1) was not the previous decoded picture
  int64_t out_pts = 0;
  int fr_num = 0, fr_den = 0;

  ~Decoder() { release(); }
  void release()
  {
    if (stream) cudaStreamSynchronize(stream);
    for (StripGeom &g : geom) {
      if (g.stream) { cudaStreamSynchronize(g.stream); cudaStreamDestroy(g.stream); }
      if (g.ev_done) cudaEventDestroy(g.ev_done);
      for (int i = 0; i < kPool; i++) {
        if (g.d_pic[i]) cudaFree(g.d_pic[i]);
        if (g.d_mvf[i]) cudaFree(g.d_mvf[i]);
        if (g.ev_mvf[i]) cudaEventDestroy(g.ev_mvf[i]);
      }
      if (g.d_dbk) cudaFree(g.d_dbk);
      if (g.d_order) cudaFree(g.d_order);
    }
    geom.clear();
    for (DecSlot &s : slots) {
      if (s.stream) cudaStreamSynchronize(s.stream);
      for (StripBufs &t : s.strips) {
        if (t.stream) { cudaStreamSynchronize(t.stream); cudaStreamDestroy(t.stream); }
        if (t.ev_parsed) cudaEventDestroy(t.ev_parsed);
        if (t.d_small) cudaFree(t.d_small);
        if (t.d_ctu_qp) cudaFree(t.d_ctu_qp);
        if (t.d_sao) cudaFree(t.d_sao);
        if (t.d_ctu_done) cudaFree(t.d_ctu_done);
        if (t.d_cu) cudaFree(t.d_cu);
        if (t.d_levels) cudaFree(t.d_levels);
        if (t.h_bases) cudaFreeHost(t.h_bases);
        if (t.h_status) cudaFreeHost(t.h_status);
      }
      if (s.d_data) cudaFree(s.d_data);
      if (s.h_data) cudaFreeHost(s.h_data);
      if (s.h_out) cudaFreeHost(s.h_out);
      if (s.d_scaling) cudaFree(s.d_scaling);
      if (s.h_scaling) cudaFreeHost(s.h_scaling);
      if (s.ev_uploaded) cudaEventDestroy(s.ev_uploaded);
      if (s.stream) cudaStreamDestroy(s.stream);
    }
    slots.clear();
    pending.clear();
    if (d_full) cudaFree(d_full);
    if (d_crop) cudaFree(d_crop);
    d_crop = nullptr;
    if (ev_out) cudaEventDestroy(ev_out);
    if (stream) cudaStreamDestroy(stream);
    d_full = nullptr; stream = nullptr; ev_out = nullptr;
    next_slot = 0; out_slot = -1; conf_tiles = 0; conf_tile_cols = 0; conf_tile_rows = 0;
  }

  // (Re)allocate for a picture size and tile grid; pictures in flight are dropped.
  bool configure(int w, int h, int tile_cols, int tile_rows, int cl = 0, int cr = 0, int ct = 0, int cb = 0)
  {
    release();
    crop_l = cl; crop_t = ct; out_w = w - cl - cr; out_h = h - ct - cb;
    out_bytes = (size_t)out_w * out_h * 3 / 2;
    const int tiles = tile_cols * tile_rows;
    fp = FrameParams{};
    fp.w = w; fp.h = h; fp.w8 = w / 8; fp.h8 = h / 8;
    fp.ctb_cols = (w + kCtb - 1) / kCtb; fp.ctb_rows = (h + kCtb - 1) / kCtb;
    fp.deblock = 1; fp.search_range = 8;
    frame_bytes = (size_t)w * h * 3 / 2;
    data_cap = frame_bytes * 2 + 65536;
    const int rows = fp.ctb_rows;
    off_flag = 0; off_prog = sizeof(int) * rows; off_ticket = 2 * sizeof(int) * rows;
    off_status = off_ticket + 2 * sizeof(int); off_bases = off_status + 2 * sizeof(int);
    off_ctx = off_bases + sizeof(uint32_t) * (rows + 1);
    small_bytes = off_ctx + (size_t)rows * CTX_COUNT;
    if (!cuda_ok(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking), "cudaStreamCreate")) return false;
    if (!cuda_ok(cudaEventCreateWithFlags(&ev_out, wait_event_flags(frame_delay > 0)), "cudaEventCreate")) return false;
    geom.resize(tiles);
    for (int i = 0; i < tiles; i++) {
      StripGeom &g = geom[i];
      const int tc = i % tile_cols, tr = i / tile_cols;                                 // tiles in raster order
      const int c0 = tc * fp.ctb_cols / tile_cols, c1 = (tc + 1) * fp.ctb_cols / tile_cols;   // colBd, uniform spacing (6.5.1)
      const int r0 = tr * fp.ctb_rows / tile_rows, r1 = (tr + 1) * fp.ctb_rows / tile_rows;   // rowBd
      g.x0 = c0 * kCtb;
      g.wd = std::min(w, c1 * kCtb) - g.x0;
      g.y0 = r0 * kCtb;
      g.ht = std::min(h, r1 * kCtb) - g.y0;
      g.fp = fp;
      g.fp.w = g.wd; g.fp.w8 = g.wd / 8; g.fp.ctb_cols = c1 - c0;
      g.fp.h = g.ht; g.fp.h8 = g.ht / 8; g.fp.ctb_rows = r1 - r0;
      g.fp.mv_edges = (tc > 0 ? 1 : 0) | (tc < tile_cols - 1 ? 2 : 0) | (tr > 0 ? 4 : 0) | (tr < tile_rows - 1 ? 8 : 0);
      g.fp.more_tiles = i < tiles - 1 ? 1 : 0;
      g.bytes = (size_t)g.wd * g.ht * 3 / 2;
      if (!cuda_ok(cudaStreamCreateWithFlags(&g.stream, cudaStreamNonBlocking), "cudaStreamCreate")) return false;
      if (!cuda_ok(cudaEventCreateWithFlags(&g.ev_done, cudaEventDisableTiming), "cudaEventCreate")) return false;
      for (int k = 0; k < kPool; k++) {
        const size_t n16 = (size_t)((g.wd + 15) / 16) * ((g.ht + 15) / 16);
        if (!cuda_ok(cudaMalloc((void **)&g.d_pic[k], g.bytes), "cudaMalloc")) return false;
        if (!cuda_ok(cudaMalloc((void **)&g.d_mvf[k], n16 * sizeof(MvField)), "cudaMalloc")) return false;
        if (!cuda_ok(cudaMemset(g.d_mvf[k], 0, n16 * sizeof(MvField)), "memset")) return false;
        if (!cuda_ok(cudaEventCreateWithFlags(&g.ev_mvf[k], cudaEventDisableTiming), "cudaEventCreate")) return false;
      }
      if (!cuda_ok(cudaMalloc((void **)&g.d_dbk, g.bytes), "cudaMalloc")) return false;
      std::vector<int> order((size_t)g.fp.ctb_cols * g.fp.ctb_rows);
      intra_wavefront_order(g.fp.ctb_cols, g.fp.ctb_rows, order.data());
      if (!cuda_ok(cudaMalloc((void **)&g.d_order, order.size() * sizeof(int)), "cudaMalloc")) return false;
      if (!cuda_ok(cudaMemcpy(g.d_order, order.data(), order.size() * sizeof(int), cudaMemcpyHostToDevice), "H2D order")) return false;
    }
    if (tiles > 1 && !cuda_ok(cudaMalloc((void **)&d_full, frame_bytes), "cudaMalloc")) return false;
    if ((out_w != w || out_h != h) && !cuda_ok(cudaMalloc((void **)&d_crop, out_bytes), "cudaMalloc")) return false;
    slots.resize(frame_delay + 1);
    for (DecSlot &s : slots) {
      if (!cuda_ok(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking), "cudaStreamCreate")) return false;
      if (!cuda_ok(cudaEventCreateWithFlags(&s.ev_uploaded, cudaEventDisableTiming), "cudaEventCreate")) return false;
      if (!cuda_ok(cudaMalloc((void **)&s.d_data, data_cap), "cudaMalloc")) return false;
      if (!cuda_ok(cudaMallocHost((void **)&s.h_out, frame_bytes), "cudaMallocHost")) return false;
      if (!cuda_ok(cudaMallocHost((void **)&s.h_data, data_cap), "cudaMallocHost")) return false;
      if (!cuda_ok(cudaMalloc((void **)&s.d_scaling, sizeof(ScalingTable)), "cudaMalloc")) return false;
      if (!cuda_ok(cudaMallocHost((void **)&s.h_scaling, sizeof(ScalingTable)), "cudaMallocHost")) return false;
      s.strips.resize(tiles);
      for (int i = 0; i < tiles; i++) {
        StripBufs &t = s.strips[i];
        const StripGeom &g = geom[i];
        if (!cuda_ok(cudaStreamCreateWithFlags(&t.stream, cudaStreamNonBlocking), "cudaStreamCreate")) return false;
        if (!cuda_ok(cudaEventCreateWithFlags(&t.ev_parsed, wait_event_flags(frame_delay > 0)), "cudaEventCreate")) return false;
        if (!cuda_ok(cudaMalloc((void **)&t.d_small, small_bytes), "cudaMalloc")) return false;
        if (!cuda_ok(cudaMalloc((void **)&t.d_ctu_qp, (size_t)g.fp.ctb_cols * rows), "cudaMalloc")) return false;
        if (!cuda_ok(cudaMalloc((void **)&t.d_sao, sizeof(SaoCtu) * g.fp.ctb_cols * rows), "cudaMalloc")) return false;
        if (!cuda_ok(cudaMalloc((void **)&t.d_ctu_done, sizeof(int) * (g.fp.ctb_cols * rows + 1)), "cudaMalloc")) return false;
        if (!cuda_ok(cudaMalloc((void **)&t.d_cu, sizeof(CuInfo) * g.fp.w8 * g.fp.h8), "cudaMalloc")) return false;
        if (!cuda_ok(cudaMalloc((void **)&t.d_levels, g.bytes * sizeof(int16_t)), "cudaMalloc")) return false;
        if (!cuda_ok(cudaMallocHost((void **)&t.h_bases, sizeof(uint32_t) * (rows + 1)), "cudaMallocHost")) return false;
        if (!cuda_ok(cudaMallocHost((void **)&t.h_status, sizeof(int) * 2), "cudaMallocHost")) return false;
      }
    }
    for (DpbEntry &d : dpb) d = DpbEntry();
    have_submitted = false;
    conf_w = w; conf_h = h; conf_tiles = tiles; conf_tile_cols = tile_cols; conf_tile_rows = tile_rows;
    return true;
  }

  bool parse_sps(const std::vector<uint8_t> &rbsp)
  {
    Sps t;
    std::string err;
    if (!parse_sps_rbsp(rbsp.data(), rbsp.size(), t, err)) { set_error("decoder: %s", err.c_str()); return false; }
    if (t.id > 15) { set_error("decoder: SPS id out of range"); return false; }
    sps_tab[t.id] = t;
    return true;
  }

  bool parse_pps(const std::vector<uint8_t> &rbsp)
  {
    Pps t;
    std::string err;
    if (!parse_pps_rbsp(rbsp.data(), rbsp.size(), t, err)) { set_error("decoder: %s", err.c_str()); return false; }
    pps_tab[t.id] = t;
    return true;
  }

  // What the CUDA kernels can reconstruct, checked once per slice against the active parameter sets
  // and the slice header.  Returns nullptr when everything is supported.
  const char *unsupported(const Sps &s, const Pps &p, const SliceHeader &sh) const
  {
    if (s.chroma_format_idc != 1) return "only 4:2:0 is supported";
    if (s.bit_depth_luma != 8 || s.bit_depth_chroma != 8) return "only 8-bit video is supported";
    if ((s.width & 7) || (s.height & 7)) return "picture sizes that are not multiples of 8 are not supported";
    if (s.conf_left < 0 || s.conf_right < 0 || s.conf_top < 0 || s.conf_bottom < 0 || s.conf_left + s.conf_right >= s.width ||
        s.conf_top + s.conf_bottom >= s.height) return "conformance window larger than the picture";
    if (s.log2_min_cb != 3 || s.log2_ctb != 6) return "coding block sizes other than 8..64 are not supported";
    if (s.log2_min_tb != 2 || s.log2_max_tb != 5) return "transform block sizes other than 4..32 are not supported";
    if (s.max_tr_depth_inter > 3 || s.max_tr_depth_intra > 3) return "transform hierarchy depth > 3 is not supported";
    if (s.amp) return "AMP is not supported";
    if (s.pcm) return "PCM is not supported";
    if (s.long_term_refs) return "long-term reference pictures are not supported";
    if (p.dependent_slices) return "dependent slice segments are not supported";
    if (p.constrained_intra) return "constrained intra prediction is not supported";
    if (p.transform_skip) return "transform skip is not supported";
    if (p.qp_delta && p.diff_cu_qp_delta_depth != 0) return "quantisation groups smaller than the CTU are not supported";
    if (abs(p.cb_qp_offset + sh.cb_qp_offset) > 12 || abs(p.cr_qp_offset + sh.cr_qp_offset) > 12) return "chroma QP offsets out of range";
    if (p.transquant_bypass) return "transquant bypass is not supported";
    if (p.tiles) {
      if (!p.uniform_spacing) return "only uniformly spaced tiles are supported";
      if (p.loop_filter_across_tiles) return "loop filtering across tiles is not supported";
      if (p.tile_cols > 32 || p.tile_rows > 32 || p.tile_cols * p.tile_rows > 64) return "too many tiles";
    }
    if (abs(sh.beta_offset_div2) > 6 || abs(sh.tc_offset_div2) > 6) return "deblocking offsets out of range";
    if (p.log2_parallel_merge_level != 2) return "parallel merge level > 2 is not supported";
    if (!sh.first_slice_in_pic) return "multiple slice segments per picture are not supported";
    if (sh.slice_type == 0) return "B slices are not supported";
    // reference pictures: any short-term set of earlier pictures; pictures that FOLLOW in output order
    // would need output reordering, which a low-delay call never uses
    if (sh.rps.num_pos > 0) return "reference picture sets with pictures that follow in output order are not supported";
    if (sh.rps.num_delta() > kPool - 1) return "too many pictures in the reference picture set";
    if (sh.slice_type == 1 && (sh.num_ref_idx_l0 < 1 || sh.num_ref_idx_l0 > 16)) return "bad number of reference indices";
    if (sh.slice_type == 1 && sh.collocated_ref_idx >= sh.num_ref_idx_l0) return "collocated_ref_idx out of range";
    return nullptr;
  }

  // returns 1 picture decoded, 0 nothing, -1 error
  int decode_slice(int nal_type, const uint8_t *payload, size_t n, int64_t pts)
  {
    // The slice header sits in the first bytes; unescape the whole NAL payload once and keep a map
    // from escaped to unescaped offsets for the entry points (which count escaped bytes).
    std::vector<uint8_t> rbsp;
    rbsp.reserve(n);
    std::vector<uint32_t> epb_pos;            // escaped offsets of removed bytes
    int zeros = 0;
    for (size_t i = 0; i < n; i++) {
      if (zeros >= 2 && payload[i] == 3) { epb_pos.push_back((uint32_t)i); zeros = 0; continue; }
      rbsp.push_back(payload[i]);
      zeros = payload[i] == 0 ? zeros + 1 : 0;
    }
    // activate the parameter sets the slice names (slice_pic_parameter_set_id is the second or third
    // syntax element; parse it ahead of the full header, which needs the sets)
    {
      BitReader pb(rbsp.data(), rbsp.size());
      pb.u(1);
      if (nal_type >= 16 && nal_type <= 23) pb.u(1);
      const uint32_t pid = pb.ue();
      if (pb.bad || pid > 63 || !pps_tab[pid].valid || !sps_tab[pps_tab[pid].sps_id].valid) { set_error("decoder: slice before parameter sets"); return -1; }
      pps = pps_tab[pid];
      sps = sps_tab[pps.sps_id];
    }
    SliceHeader sh;
    {
      std::string err;
      if (!parse_slice_header_rbsp(rbsp.data(), rbsp.size(), nal_type, sps, pps, sh, err)) { set_error("decoder: %s", err.c_str()); return -1; }
    }
    if (const char *why = unsupported(sps, pps, sh)) { set_error("decoder: %s", why); return -1; }
    const bool idr = nal_type == 19 || nal_type == 20;
    const int slice_type = sh.slice_type, qp = sh.qp;
    const int deblock = !sh.deblock_disabled;
    // picture order count (8.3.1) and the reference it names: this decoder keeps one reference, the
    // previously decoded picture, so the picture with POC - 1 must be exactly that one.  A gap (a P
    // picture lost on the way) is reported so that the application can wait for / request an IDR
    // instead of silently predicting from the wrong picture.
    int poc = 0, poc_msb = 0;
    if (!idr) {
      const int max_lsb = 1 << sps.log2_max_poc;
      if (sh.poc_lsb < prev_poc_lsb && prev_poc_lsb - sh.poc_lsb >= max_lsb / 2) poc_msb = prev_poc_msb + max_lsb;
      else if (sh.poc_lsb > prev_poc_lsb && sh.poc_lsb - prev_poc_lsb > max_lsb / 2) poc_msb = prev_poc_msb - max_lsb;
      else poc_msb = prev_poc_msb;
      if (nal_type >= 16 && nal_type <= 18) poc_msb = 0;       // BLA
      poc = poc_msb + sh.poc_lsb;
    }
    // Reference picture set (8.3.2): pictures it does not name leave the buffer; the pictures it marks
    // "used by the current picture" form list 0 (closest first, repeated cyclically up to
    // num_ref_idx_l0_active).  A named picture that is not there was lost on the way: like OpenHEVC the
    // decoder conceals with the nearest picture it has -- the reference application just keeps feeding
    // NALs (openhevcfilter.cpp:145-152) -- and counts the event (b200_dec_missing_refs) so that an
    // application can ask the sender for an IDR instead of drifting unknowingly until the next one.
    int n_refs = 0, ref_pool[16], ref_poc[16], cur_idx = -1;
    {
      if (idr) for (DpbEntry &d : dpb) d.valid = false;
      std::vector<int> curr;
      for (int i = 0; i < sh.rps.num_delta(); i++) if (sh.rps.used[i]) curr.push_back(poc + sh.rps.delta_poc[i]);
      bool concealing[kPool] = {};
      if (slice_type != 2) {
        if (curr.empty()) { set_error("decoder: P slice with an empty reference picture set"); return -1; }
        bool counted = false;
        for (int i = 0; i < sh.num_ref_idx_l0; i++) {
          const int want = curr[i % curr.size()];
          int found = -1, best = -1;
          for (int k = 0; k < kPool; k++) {
            if (!dpb[k].valid) continue;
            if (dpb[k].poc == want) found = k;
            if (best < 0 || std::abs(dpb[k].poc - want) < std::abs(dpb[best].poc - want)) best = k;
          }
          if (found < 0) {
            if (best < 0) { set_error("decoder: P slice without a reference picture (waiting for an IDR)"); return -1; }
            if (!counted) { missing_refs++; counted = true; }
            found = best;
            concealing[best] = true;
          }
          ref_pool[n_refs] = found; ref_poc[n_refs++] = want;
        }
      }
      if (!idr)
        for (int k = 0; k < kPool; k++) {
          bool keep = concealing[k];               // a stand-in for a lost picture stays until it is not needed
          for (int i = 0; i < sh.rps.num_delta(); i++) keep |= dpb[k].poc == poc + sh.rps.delta_poc[i];
          if (!keep) dpb[k].valid = false;
        }
      for (int k = 0; k < kPool && cur_idx < 0; k++) if (!dpb[k].valid) cur_idx = k;
      if (cur_idx < 0) { set_error("decoder: decoded picture buffer full"); return -1; }
    }
    fr_num = sps.fps_num; fr_den = sps.fps_den;
    // geometry: picture size from the SPS, tile columns from the PPS (pictures in flight are dropped
    // when either changes)
    if (conf_w != sps.width || conf_h != sps.height || conf_tile_cols != pps.tile_cols || conf_tile_rows != pps.tile_rows ||
        crop_l != sps.conf_left || crop_t != sps.conf_top || out_w != sps.width - sps.conf_left - sps.conf_right ||
        out_h != sps.height - sps.conf_top - sps.conf_bottom) {
      const int ctb_cols = (sps.width + kCtb - 1) / kCtb, ctb_rows = (sps.height + kCtb - 1) / kCtb;
      if (pps.tile_cols > 1 && pps.tile_cols > ctb_cols / 2) { set_error("decoder: tile columns narrower than two CTUs are not supported"); return -1; }
      if (pps.tile_rows > ctb_rows) { set_error("decoder: more tile rows than CTU rows"); return -1; }
      if (!configure(sps.width, sps.height, pps.tile_cols, pps.tile_rows, sps.conf_left, sps.conf_right, sps.conf_top, sps.conf_bottom)) return -1;
    }
    const int rows = fp.ctb_rows, tiles = conf_tiles;
    // substreams: one per CTU row of every tile with WPP, else one per tile
    int n_sub = 0;
    std::vector<int> first_sub(tiles + 1, 0);
    for (int i = 0; i < tiles; i++) { first_sub[i] = n_sub; n_sub += pps.wpp ? geom[i].fp.ctb_rows : 1; }
    first_sub[tiles] = n_sub;
    const int n_entry = (int)sh.entry.size();
    const std::vector<uint32_t> &entry = sh.entry;
    if (n_entry != n_sub - 1) {
      set_error("decoder: %d entry points for %d tile(s) with %d substream(s)", n_entry, tiles, n_sub);
      return -1;
    }
    prev_poc = poc; prev_poc_lsb = idr ? 0 : sh.poc_lsb; prev_poc_msb = poc_msb; have_submitted = true;
    dpb[cur_idx].valid = true; dpb[cur_idx].poc = poc;
    const size_t hdr_unesc_pos = sh.data_offset;
    // escaped offset of the first slice-data byte
    const size_t hdr_unesc = hdr_unesc_pos;
    size_t hdr_esc = hdr_unesc;
    for (uint32_t e : epb_pos) { if (e < hdr_esc + 1) hdr_esc++; else break; }
    // substream boundaries in escaped bytes -> unescaped offsets
    auto to_unesc = [&](size_t esc) {
      size_t k = std::lower_bound(epb_pos.begin(), epb_pos.end(), (uint32_t)esc) - epb_pos.begin();
      return esc - k;
    };
    DecSlot &sl = slots[next_slot];
    const size_t data_len = rbsp.size() - hdr_unesc;
    if (data_len > data_cap) { set_error("decoder: slice larger than the staging buffer"); return -1; }
    {
      size_t esc = hdr_esc;
      uint32_t prev = 0;
      for (int i = 0, k = 0; i < tiles; i++) {
        const int per_tile = first_sub[i + 1] - first_sub[i];
        for (int j = 0; j < per_tile; j++, k++) {
          const uint32_t base = (uint32_t)(to_unesc(esc) - hdr_unesc);
          if (base < prev || base > data_len) { set_error("decoder: entry points run past the slice data"); return -1; }
          prev = base;
          sl.strips[i].h_bases[j] = base;
          if (j == 0 && i > 0) sl.strips[i - 1].h_bases[first_sub[i] - first_sub[i - 1]] = base;
          if (k < n_sub - 1) esc += entry[k];
        }
      }
      sl.strips[tiles - 1].h_bases[first_sub[tiles] - first_sub[tiles - 1]] = (uint32_t)data_len;
    }
    memcpy(sl.h_data, rbsp.data() + hdr_unesc, data_len);
    sl.pts = pts;
#define DEC_CHECK(expr, what) do { if (!cuda_ok((expr), (what))) return -1; } while (0)
    DEC_CHECK(cudaMemcpyAsync(sl.d_data, sl.h_data, data_len, cudaMemcpyHostToDevice, sl.stream), "H2D slice");
    if (sps.scaling_list) {              // 7.4.5: lists of the PPS, else of the SPS, else the default ones
      ScalingTable def;
      if (!pps.scaling_list && !sps.scaling_list_data) def.set_default();
      const ScalingTable &lists = pps.scaling_list ? pps.lists : (sps.scaling_list_data ? sps.lists : def);
      memcpy(sl.h_scaling, &lists, sizeof(ScalingTable));
      DEC_CHECK(cudaMemcpyAsync(sl.d_scaling, sl.h_scaling, sizeof(ScalingTable), cudaMemcpyHostToDevice, sl.stream), "H2D scaling lists");
    }
    DEC_CHECK(cudaEventRecord(sl.ev_uploaded, sl.stream), "record upload");
    for (int i = 0; i < tiles; i++) {
      StripBufs &t = sl.strips[i];
      const StripGeom &g = geom[i];
      t.fp = g.fp;
      t.fp.qp = qp; t.fp.qp_c = kChromaQpD[qp]; t.fp.is_idr = slice_type == 2 ? 1 : 0; t.fp.deblock = deblock;
      t.fp.no_wpp = pps.wpp ? 0 : 1;
      t.fp.ctu_qp = pps.qp_delta ? t.d_ctu_qp : nullptr; t.fp.ctu_delta = nullptr; t.fp.ctu_first = nullptr;
      t.fp.sao_flags = (sh.sao_luma ? 1 : 0) | (sh.sao_chroma ? 2 : 0);
      t.fp.sao = t.fp.sao_flags ? t.d_sao : nullptr;
      t.fp.init_type = slice_type == 2 ? 0 : (sh.cabac_init_flag ? 2 : 1);
      t.fp.tr_depth_inter = sps.max_tr_depth_inter; t.fp.tr_depth_intra = sps.max_tr_depth_intra;
      t.fp.sign_hiding = pps.sign_hiding; t.fp.strong_intra = sps.strong_intra_smoothing;
      t.fp.cb_qp_offset = pps.cb_qp_offset + sh.cb_qp_offset; t.fp.cr_qp_offset = pps.cr_qp_offset + sh.cr_qp_offset;
      t.fp.cb_qp_offset_pps = pps.cb_qp_offset; t.fp.cr_qp_offset_pps = pps.cr_qp_offset;
      t.fp.beta_offset_div2 = sh.beta_offset_div2; t.fp.tc_offset_div2 = sh.tc_offset_div2;
      t.fp.scaling = sps.scaling_list ? sl.d_scaling : nullptr;
      t.fp.n_refs = std::max(n_refs, 1); t.fp.max_merge = sh.max_merge_cand;
      for (int k = 0; k < 16; k++) t.fp.ref_dist[k] = (int16_t)(k < n_refs ? poc - ref_poc[k] : 1);
      t.fp.col_mvf = (slice_type != 2 && sh.tmvp) ? g.d_mvf[ref_pool[sh.collocated_ref_idx]] : nullptr;
      if (t.fp.col_mvf) DEC_CHECK(cudaStreamWaitEvent(t.stream, g.ev_mvf[ref_pool[sh.collocated_ref_idx]], 0), "stream wait");
      t.fp.ctu_done = t.d_ctu_done; t.fp.any_intra = t.d_ctu_done + g.fp.ctb_cols * rows; t.fp.intra_in_p = 0;
      DEC_CHECK(cudaMemsetAsync(t.fp.any_intra, 0, sizeof(int), t.stream), "memset any_intra");
      int *sync_flag = (int *)(t.d_small + off_flag), *progress = (int *)(t.d_small + off_prog);
      int *status = (int *)(t.d_small + off_status);
      uint32_t *d_bases = (uint32_t *)(t.d_small + off_bases);
      DEC_CHECK(cudaStreamWaitEvent(t.stream, sl.ev_uploaded, 0), "stream wait");
      DEC_CHECK(cudaMemcpyAsync(d_bases, t.h_bases, sizeof(uint32_t) * (first_sub[i + 1] - first_sub[i] + 1), cudaMemcpyHostToDevice, t.stream), "H2D bases");
      DEC_CHECK(cudaMemsetAsync(t.d_levels, 0, g.bytes * sizeof(int16_t), t.stream), "memset levels");
      DEC_CHECK(launch_parse(t.fp, sl.d_data, d_bases, t.d_cu, t.d_levels, t.d_small + off_ctx, sync_flag, progress, status, t.stream), "parse launch");
      count_launch(1);
      if (sps.tmvp) {                      // the motion field later pictures' temporal candidates read
        DEC_CHECK(launch_store_mvf(t.fp, t.d_cu, g.d_mvf[cur_idx], t.stream), "mvf launch");
        DEC_CHECK(cudaEventRecord(g.ev_mvf[cur_idx], t.stream), "record mvf");
        count_launch(1);
      }
      DEC_CHECK(cudaMemcpyAsync(t.h_status, status, sizeof(int) * 2, cudaMemcpyDeviceToHost, t.stream), "D2H status");
      DEC_CHECK(cudaEventRecord(t.ev_parsed, t.stream), "record parse");
    }
    sl.cur_idx = cur_idx; sl.n_refs = n_refs;
    for (int k = 0; k < n_refs; k++) sl.ref_pool[k] = ref_pool[k];
    pending.push_back(next_slot);
    next_slot = (next_slot + 1) % (int)slots.size();
    if ((int)pending.size() <= frame_delay) return 0;
    return finish_oldest();
  }

  // Reconstructs the oldest parsed picture (prediction + residual, deblocking, strip by strip) and
  // copies it to the host.  Returns 1 with the picture in slots[out_slot].h_out, -1 on error.
  int finish_oldest()
  {
    const int idx = pending.front();
    pending.pop_front();
    DecSlot &sl = slots[idx];
    const int tiles = (int)sl.strips.size();
    for (int i = 0; i < tiles; i++) {
      StripBufs &t = sl.strips[i];
      if (!cuda_ok(cudaEventSynchronize(t.ev_parsed), "sync parse")) { dpb[sl.cur_idx].valid = false; return -1; }
      if (t.h_status[0] != 0) {
        static const char *const why[] = {"", "escape code too long", "(unused)", "inter partition other than 2Nx2N", "mvd too long",
          "(unused)", "(unused)", "(unused)", "end_of_slice_segment_flag mismatch",
          "end_of_subset_one_bit missing", "(unused)", "cu_qp_delta out of range",
          "motion vector reaches across a tile boundary", "(unused)"};
        int c = t.h_status[0];
        set_error("decoder: unsupported or corrupt slice data (%s)", c > 0 && c <= 13 ? why[c] : "unknown");
        dpb[sl.cur_idx].valid = false;     // never a reference: what follows it conceals and counts
        return -1;
      }
    }
    const size_t ysz = (size_t)fp.w * fp.h;
    for (int i = 0; i < tiles; i++) {      // the strips reconstruct concurrently, each on its own stream
      StripBufs &t = sl.strips[i];
      StripGeom &g = geom[i];
      FrameParams &f = t.fp;
      int *ticket = (int *)(t.d_small + off_ticket);
      // with SAO the picture is reconstructed and deblocked in g.d_dbk; SAO writes the output picture
      uint8_t *const out_rec = g.d_pic[sl.cur_idx];
      uint8_t *rec = f.sao_flags ? g.d_dbk : out_rec;
      if (f.is_idr) {
        DEC_CHECK(launch_intra_decode(f, rec, t.d_levels, t.d_cu, ticket, g.d_order, g.stream), "intra decode launch");
        count_launch(1);
      } else {
        RefList refs{};
        refs.n = sl.n_refs;
        for (int k = 0; k < sl.n_refs; k++) refs.pic[k] = g.d_pic[sl.ref_pool[k]];
        cudaError_t e = launch_inter_decode(f, refs, rec, t.d_levels, t.d_cu, g.stream);
        DEC_CHECK(e, "inter decode launch");
        // intra CUs of the P picture predict from the reconstructed inter CUs (the kernel returns at
        // once when the parser met none)
        DEC_CHECK(launch_intra_decode(f, rec, t.d_levels, t.d_cu, ticket, g.d_order, g.stream), "intra decode launch");
        count_launch(2);
      }
      if (f.deblock) {
        DEC_CHECK(launch_deblock(f, rec, t.d_cu, g.stream), "deblock launch");
        count_launch(2);
      }
      if (f.sao_flags) {
        DEC_CHECK(launch_sao_decode(f, rec, out_rec, t.d_sao, g.stream), "sao launch");
        count_launch(1);
        rec = out_rec;
      }
      if (tiles > 1) {                     // place the tile in the whole picture
        const size_t sy = (size_t)g.wd * g.ht, cw = fp.w / 2;
        DEC_CHECK(cudaMemcpy2DAsync(d_full + (size_t)g.y0 * fp.w + g.x0, fp.w, rec, g.wd, g.wd, g.ht, cudaMemcpyDeviceToDevice, g.stream), "assemble Y");
        DEC_CHECK(cudaMemcpy2DAsync(d_full + ysz + (size_t)(g.y0 / 2) * cw + g.x0 / 2, cw, rec + sy, g.wd / 2, g.wd / 2, g.ht / 2, cudaMemcpyDeviceToDevice, g.stream), "assemble U");
        DEC_CHECK(cudaMemcpy2DAsync(d_full + ysz + ysz / 4 + (size_t)(g.y0 / 2) * cw + g.x0 / 2, cw, rec + sy + sy / 4, g.wd / 2, g.wd / 2, g.ht / 2, cudaMemcpyDeviceToDevice, g.stream), "assemble V");
      }
      DEC_CHECK(cudaEventRecord(g.ev_done, g.stream), "record strip");
      DEC_CHECK(cudaStreamWaitEvent(stream, g.ev_done, 0), "stream wait");
    }
    d_out = tiles > 1 ? d_full : geom[0].d_pic[sl.cur_idx];
    // host output: the coded picture; libOpenHevcGetOutput points into it at the window origin with the coded
    // pitches (OpenHEVC does the same: linesize > width), which keeps the chroma pitch even -- the reference
    // reads chroma rows at i * (nUPitch / 2) for even i (openhevcfilter.cpp:213-227)
    if (host_output) DEC_CHECK(cudaMemcpyAsync(sl.h_out, d_out, frame_bytes, cudaMemcpyDeviceToHost, stream), "D2H picture");
    if (cropped()) {                       // device-resident hand-over: the conformance window as a packed picture
      const size_t oy = (size_t)out_w * out_h;
      const int cw = fp.w / 2, ow = out_w / 2;
      DEC_CHECK(cudaMemcpy2DAsync(d_crop, out_w, d_out + (size_t)crop_t * fp.w + crop_l, fp.w, out_w, out_h, cudaMemcpyDeviceToDevice, stream), "crop Y");
      DEC_CHECK(cudaMemcpy2DAsync(d_crop + oy, ow, d_out + ysz + (size_t)(crop_t / 2) * cw + crop_l / 2, cw, ow, out_h / 2, cudaMemcpyDeviceToDevice, stream), "crop U");
      DEC_CHECK(cudaMemcpy2DAsync(d_crop + oy + oy / 4, ow, d_out + ysz + ysz / 4 + (size_t)(crop_t / 2) * cw + crop_l / 2, cw, ow, out_h / 2, cudaMemcpyDeviceToDevice, stream), "crop V");
      d_out = d_crop;
    }
    DEC_CHECK(cudaEventRecord(ev_out, stream), "record picture");
    DEC_CHECK(cudaEventSynchronize(ev_out), "sync picture");
#undef DEC_CHECK
    out_slot = idx; out_pts = sl.pts; pictures++;
    return 1;
  }

  // libOpenHevcDecode(h, NULL, 0, pts): hand over the next held-back picture, 0 when none is left.
  int drain_one() { return pending.empty() ? 0 : finish_oldest(); }

  int decode_nal(const uint8_t *nal, size_t n, int64_t pts)
  {
    if (n < 2) return 0;
    if (nal[0] & 0x80) { set_error("decoder: forbidden_zero_bit set"); return -1; }
    const int type = (nal[0] >> 1) & 63;
    if (type == 32) { vps_seen = true; return 0; }
    if (type == 33) return parse_sps(unescape(nal + 2, n - 2)) ? 0 : -1;
    if (type == 34) return parse_pps(unescape(nal + 2, n - 2)) ? 0 : -1;
    if (type <= 9 || (type >= 16 && type <= 21)) return decode_slice(type, nal + 2, n - 2, pts);
    return 0;      // SEI, AUD, ... ignored
  }
};

}  // namespace b200

using b200::Decoder;

extern "C" {

OpenHevc_Handle libOpenHevcInit(int nb_pthreads, int thread_type)
{
  // Frame threading (thread_type 1 or 3) in OpenHEVC keeps nb_pthreads pictures in flight and
  // delays output by nb_pthreads - 1 pictures; here the pictures in flight are concurrent CABAC
  // parses on the GPU.  Slice threading (2) has no delay: WPP substreams are always parallel.
  Decoder *d = new Decoder();
  if (thread_type & 1) d->frame_delay = std::min(std::max(nb_pthreads, 1), 64) - 1;
  return d;
}

int libOpenHevcStartDecoder(OpenHevc_Handle h)
{
  Decoder *d = (Decoder *)h;
  if (!d) return -1;
  if (b200_device_count() <= 0) { b200::set_error("no CUDA device: the B200 decoder has no CPU fallback"); return -1; }
  d->started = true;
  return 0;
}

int libOpenHevcDecode(OpenHevc_Handle h, const unsigned char *buff, int nal_len, int64_t pts)
{
  Decoder *d = (Decoder *)h;
  if (!d || !d->started) { b200::set_error("libOpenHevcDecode: decoder not started"); return -1; }
  if (!buff || nal_len <= 0) return d->drain_one();
  // split on start codes (00 00 01 / 00 00 00 01)
  int got = 0;
  size_t n = (size_t)nal_len;
  auto find_sc = [&](size_t from, size_t &sc_len) -> size_t {
    for (size_t k = from; k + 3 <= n; k++)
      if (buff[k] == 0 && buff[k + 1] == 0 && buff[k + 2] == 1) { sc_len = 3; return k; }
    return n;
  };
  size_t sc_len = 0;
  size_t pos = find_sc(0, sc_len);
  if (pos == n) { b200::set_error("libOpenHevcDecode: no start code in the buffer"); return -1; }
  while (pos < n) {
    size_t start = pos + sc_len;
    size_t next_len = 0;
    size_t next = find_sc(start, next_len);
    size_t end = next;
    while (end > start && next < n && buff[end - 1] == 0) end--;     // zero_byte of a 4-byte start code / trailing zeros
    int rc = d->decode_nal(buff + start, end - start, pts);
    if (rc < 0) return -1;
    got |= rc;
    pos = next; sc_len = next_len;
  }
  return got;
}

int libOpenHevcGetOutput(OpenHevc_Handle h, int got_picture, OpenHevc_Frame *frame)
{
  Decoder *d = (Decoder *)h;
  if (!d || !frame || !got_picture || d->out_slot < 0) return 0;
  const size_t ysz = (size_t)d->fp.w * d->fp.h;
  const int cw = d->fp.w / 2;
  uint8_t *out = d->slots[d->out_slot].h_out;
  frame->pvY = out + (size_t)d->crop_t * d->fp.w + d->crop_l;
  frame->pvU = out + ysz + (size_t)(d->crop_t / 2) * cw + d->crop_l / 2;
  frame->pvV = out + ysz + ysz / 4 + (size_t)(d->crop_t / 2) * cw + d->crop_l / 2;
  libOpenHevcGetPictureInfo(h, &frame->frameInfo);
  return 1;
}

void libOpenHevcGetPictureInfo(OpenHevc_Handle h, OpenHevc_FrameInfo *info)
{
  Decoder *d = (Decoder *)h;
  if (!d || !info) return;
  memset(info, 0, sizeof(*info));
  info->nWidth = d->out_w; info->nHeight = d->out_h;       // the conformance window (= the coded size without one)
  info->nYPitch = d->fp.w; info->nUPitch = d->fp.w / 2; info->nVPitch = d->fp.w / 2;       // pitches of the coded picture
  info->nBitDepth = 8; info->chromat_format = 1;
  info->sample_aspect_ratio.num = 1; info->sample_aspect_ratio.den = 1;
  // VUI timing when the stream carries it; never 0/0 -- the reference copies this into vInfo
  // (openhevcfilter.cpp:232-233) and DisplayFilter divides by it (displayfilter.cpp:153)
  info->frameRate.num = d->fr_num > 0 && d->fr_den > 0 ? d->fr_num : 30;
  info->frameRate.den = d->fr_num > 0 && d->fr_den > 0 ? d->fr_den : 1;
  info->display_picture_number = d->pictures - 1;
  info->nTimeStamp = d->out_pts;
}

void libOpenHevcSetTemporalLayer_id(OpenHevc_Handle, int) {}
void libOpenHevcSetActiveDecoders(OpenHevc_Handle, int) {}
void libOpenHevcSetViewLayers(OpenHevc_Handle, int) {}
void libOpenHevcSetDebugMode(OpenHevc_Handle, int) {}
void libOpenHevcSetCheckMD5(OpenHevc_Handle, int) {}
const char *libOpenHevcVersion(OpenHevc_Handle) { return "b200-hevc-dec 0.1 (sm_100a)"; }
void libOpenHevcFlush(OpenHevc_Handle h)
{
  Decoder *d = (Decoder *)h;
  if (!d) return;
  for (b200::DecSlot &s : d->slots) {
    cudaStreamSynchronize(s.stream);
    for (b200::StripBufs &t : s.strips) cudaStreamSynchronize(t.stream);
  }
  d->pending.clear();
  for (Decoder::DpbEntry &e : d->dpb) e.valid = false;
  d->have_submitted = false; d->out_slot = -1; d->d_out = nullptr;
}
void libOpenHevcClose(OpenHevc_Handle h) { delete (Decoder *)h; }

const uint8_t *b200_dec_output_dev(OpenHevc_Handle h)
{
  Decoder *d = (Decoder *)h;
  return d && d->out_slot >= 0 ? d->d_out : NULL;
}

int b200_dec_probe(const uint8_t *buf, size_t n, b200_stream_info *info, uint8_t *scaling_table)
{
  if (!buf || !info || info->struct_size < (int)(3 * sizeof(int))) { b200::set_error("b200_dec_probe: bad arguments"); return B200_ERR_ARG; }
  std::unique_ptr<Decoder> d(new Decoder());
  int last_pps = -1;
  b200_stream_info o;
  memset(&o, 0, sizeof(o));
  const char *why = nullptr;
  size_t pos = 0;
  auto next_sc = [&](size_t from) { for (size_t k = from; k + 3 <= n; k++) if (buf[k] == 0 && buf[k + 1] == 0 && buf[k + 2] == 1) return k; return n; };
  for (pos = next_sc(0); pos < n;) {
    const size_t start = pos + 3, next = next_sc(start);
    size_t end = next;
    while (end > start && next < n && buf[end - 1] == 0) end--;
    if (end - start >= 2) {
      const int type = (buf[start] >> 1) & 63;
      if (type == 33 && !d->parse_sps(b200::unescape(buf + start + 2, end - start - 2))) return B200_ERR_ARG;
      if (type == 34) {
        b200::Pps t;
        std::string err;
        const std::vector<uint8_t> r = b200::unescape(buf + start + 2, end - start - 2);
        if (!b200::parse_pps_rbsp(r.data(), r.size(), t, err)) { b200::set_error("b200_dec_probe: %s", err.c_str()); return B200_ERR_ARG; }
        d->pps_tab[t.id] = t;
        last_pps = t.id;
      }
      if (type <= 9 || (type >= 16 && type <= 21)) {           // a slice segment: its header against the sets it names
        const std::vector<uint8_t> r = b200::unescape(buf + start + 2, end - start - 2);
        b200::BitReader pb(r.data(), r.size());
        pb.u(1);
        if (type >= 16 && type <= 23) pb.u(1);
        const uint32_t pid = pb.ue();
        if (pb.bad || pid > 63 || !d->pps_tab[pid].valid || !d->sps_tab[d->pps_tab[pid].sps_id].valid) {
          b200::set_error("b200_dec_probe: slice before its parameter sets");
          return B200_ERR_ARG;
        }
        const b200::Pps &sp = d->pps_tab[pid];
        const b200::Sps &ss = d->sps_tab[sp.sps_id];
        b200::SliceHeader sh;
        std::string err;
        if (!b200::parse_slice_header_rbsp(r.data(), r.size(), type, ss, sp, sh, err)) { b200::set_error("b200_dec_probe: %s", err.c_str()); return B200_ERR_ARG; }
        last_pps = (int)pid;
        o.slices++;
        o.slice_type = sh.slice_type; o.slice_qp = sh.qp; o.entry_points = (int)sh.entry.size();
        o.num_ref_idx_l0 = sh.slice_type == 2 ? 0 : sh.num_ref_idx_l0; o.rps_pictures = sh.rps.num_delta();
        if (!why) why = d->unsupported(ss, sp, sh);
      }
    }
    pos = next;
  }
  if (last_pps < 0 || !d->pps_tab[last_pps].valid || !d->sps_tab[d->pps_tab[last_pps].sps_id].valid) {
    b200::set_error("b200_dec_probe: no SPS / PPS pair in the buffer");
    return B200_ERR_ARG;
  }
  const b200::Pps &p = d->pps_tab[last_pps];
  const b200::Sps &s = d->sps_tab[p.sps_id];
  o.coded_width = s.width; o.coded_height = s.height;
  o.width = s.width - s.conf_left - s.conf_right; o.height = s.height - s.conf_top - s.conf_bottom;
  o.crop_left = s.conf_left; o.crop_top = s.conf_top;
  o.fps_num = s.fps_num; o.fps_den = s.fps_den;
  o.tile_cols = p.tile_cols; o.tile_rows = p.tile_rows; o.wpp = p.wpp;
  o.sao = s.sao; o.sign_hiding = p.sign_hiding; o.qp_delta = p.qp_delta; o.tmvp = s.tmvp; o.strong_intra = s.strong_intra_smoothing;
  o.cabac_init_present = p.cabac_init_present;
  o.scaling_list = !s.scaling_list ? 0 : (p.scaling_list ? 3 : (s.scaling_list_data ? 2 : 1));
  o.max_tr_depth_inter = s.max_tr_depth_inter; o.max_tr_depth_intra = s.max_tr_depth_intra;
  o.max_dec_pic_buffering = s.max_dec_pic_buffering;
  if (!o.slices) {
    b200::SliceHeader sh;                            // a slice that uses nothing beyond the parameter sets
    sh.slice_type = 1; sh.num_ref_idx_l0 = 1;
    why = d->unsupported(s, p, sh);
  }
  o.decodable = why ? 0 : 1;
  if (why) snprintf(o.reason, sizeof(o.reason), "%s", why);
  if (scaling_table && o.scaling_list) {
    b200::ScalingTable def;
    if (o.scaling_list == 1) def.set_default();
    const b200::ScalingTable &t = o.scaling_list == 3 ? p.lists : (o.scaling_list == 2 ? s.lists : def);
    memcpy(scaling_table, &t, sizeof(t));
  }
  const int keep = info->struct_size;
  memcpy(info, &o, std::min<size_t>((size_t)keep, sizeof(o)));
  info->struct_size = keep;
  return B200_OK;
}

int b200_dec_missing_refs(OpenHevc_Handle h)
{
  Decoder *d = (Decoder *)h;
  return d ? d->missing_refs : 0;
}

void b200_dec_set_host_output(OpenHevc_Handle h, int on)
{
  Decoder *d = (Decoder *)h;
  if (d) d->host_output = on != 0;
}

int b200_dec_last_picture(OpenHevc_Handle h, uint8_t *dst, int cap)
{
  Decoder *d = (Decoder *)h;
  if (!d || !dst || d->out_slot < 0 || (size_t)cap < d->out_bytes) return -1;
  OpenHevc_Frame fr;
  libOpenHevcGetOutput(h, 1, &fr);                           // the window, packed
  const int w = d->out_w, hh = d->out_h;
  for (int r = 0; r < hh; r++) memcpy(dst + (size_t)r * w, (const uint8_t *)fr.pvY + (size_t)r * d->fp.w, w);
  for (int r = 0; r < hh / 2; r++) {
    memcpy(dst + (size_t)w * hh + (size_t)r * (w / 2), (const uint8_t *)fr.pvU + (size_t)r * (d->fp.w / 2), w / 2);
    memcpy(dst + (size_t)w * hh * 5 / 4 + (size_t)r * (w / 2), (const uint8_t *)fr.pvV + (size_t)r * (d->fp.w / 2), w / 2);
  }
  return (int)d->out_bytes;
}

}  // extern "C"
