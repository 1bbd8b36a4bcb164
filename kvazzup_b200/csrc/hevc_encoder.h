// Host-side encoder object (internal header).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <deque>
#include <vector>

#include "hevc_common.h"

namespace b200 {

struct EncoderConfig {
  int width = 0, height = 0;
  int qp = 32;
  int intra_period = 64;     // IDR every n pictures; 0 = first picture only
  int search_range = 8;      // full-sample motion search range (+-)
  int deblock = 1;
  int debug = 0;             // keep a copy of the reconstruction before deblocking
  int overlap_idr = 1;       // run an IDR on its own stream, concurrently with the P pictures queued before it
                             // (B200, 1080p GOP 64: 1812 -> 2200 pictures/s)
  int qp_delta = 0;          // cu_qp_delta_enabled_flag: per-CTU QP offsets (ROI), one quantisation group per CTU
  // Tile-column mode (hevc_tiles.cu): this encoder codes one tile of a larger picture as a picture of
  // its own.  mv_edges bit 0 / 1 / 2 / 3 = left / right / top / bottom edge is an interior tile edge; more_tiles = tiles
  // follow in the slice; no_wpp = one substream per tile (Main profile: tiles or WPP, not both);
  // raw = collect() hands back the substreams instead of an access unit.
  int mv_edges = 0, more_tiles = 0, no_wpp = 0, raw = 0;
  int fps_num = 0, fps_den = 0;   // both > 0: VUI timing info in the SPS (the reference's DisplayFilter divides by
                             // the frame rate the decoder reports, displayfilter.cpp:153)
  int sao = 0;               // sample adaptive offset after deblocking (hevc_sao.cu); 2 = with merge flags
  int intra_in_p = 0;        // 16x16 intra CUs in P pictures where inter prediction is poor
  int intra_satd = 0;        // I pictures: SATD instead of SAD in the intra mode search
  int subme_satd = 0;        // P pictures: SATD instead of SAD in the fractional motion refinement
  int vaq = 0;               // variance adaptive quantisation strength (Kvazaar --vaq), 1..20; needs qp_delta
  int scaling_list = 0;      // 1 = scaling_list_enabled_flag with the default lists (Kvazaar --scaling-list default)
  int vps_period = 1;        // Kvazaar's vps-period: VPS / SPS / PPS before every n-th IDR picture (the first always); 0 =
                             // before the first picture only
  int src_width = 0, src_height = 0;   // size of the pictures passed in when it is not a multiple of 8 (even, at most 6
                             // samples short of width / height; 0 = width / height): the encoder pads by edge
                             // repetition and the SPS carries a conformance window
  int me_coarse = 0;         // two-level motion search: range of the coarse level in coarse (4x4-mean) samples,
                             // a multiple of 4; search_range (<= 16) is then the window around each centre
  int depth = 1;             // pictures in flight (Kvazaar's owf + 1): output of picture n is
                             // returned by the call that submits picture n + depth - 1
};

// Everything one in-flight picture owns.  The prediction chain (ME, reconstruction, deblocking)
// runs on the encoder's main stream in picture order; entropy coding of a finished picture runs
// on the slot's own stream, overlapping the next pictures' prediction chain.
struct FrameSlot {
  CuInfo *d_cu = nullptr;
  int16_t *d_levels = nullptr;
  uint8_t *d_rows = nullptr;       // per-row substreams
  uint32_t *d_recs = nullptr;      // bin records (binariser -> arithmetic coder)
  uint8_t *d_qpinfo = nullptr;     // qp_delta: ctu_qp | ctu_delta | ctu_first, one byte per CTU each
  uint8_t *h_ctu_qp = nullptr;     // pinned staging of ctu_qp
  uint32_t *d_vaq = nullptr;       // vaq: per-CTU sample statistics (6 words each)
  uint8_t *d_small = nullptr;      // row_len | sync flags | progress | ticket | bins | sync contexts
  int *d_ctu_done = nullptr;       // intra wavefront: one flag per CTU, then the picture's "any intra CU" flag
  uint8_t *d_dbk = nullptr;        // sao: the deblocked picture (SAO reads it and writes the reconstruction ring)
  SaoCtu *d_sao = nullptr;         // sao: per-CTU parameters (k_sao_ctu -> k_binarise)
  uint8_t *d_src = nullptr;        // device copy of a host-supplied picture
  uint8_t *h_src = nullptr;        // pinned staging of the input
  uint8_t *h_pack = nullptr;       // mapped pinned: packed substreams
  uint32_t *h_hdr = nullptr;       // mapped pinned: {total, row_len[rows]}
  cudaStream_t stream = nullptr;
  cudaEvent_t ev_pred = nullptr, ev_done = nullptr;
  cudaEvent_t pev[18] = {};        // profiling: begin/end event per kernel slot (kernel ids below)
  unsigned prof_mask = 0;          // which kernel ids were recorded for the picture in this slot
  bool idr = false;
  bool headers = false;            // the access unit starts with the parameter sets
  int poc = 0, qp = 0;
  long long seq = 0;
};

// What the parameter sets and the slice header describe (shared by Encoder and TiledEncoder).
struct StreamLayout {
  int w = 0, h = 0, deblock = 1, qp_delta = 0;
  int fps_num = 0, fps_den = 0;   // VUI timing info when both > 0
  int sao = 0;               // sample_adaptive_offset_enabled_flag; slices switch luma and chroma SAO on
  int tile_cols = 1;         // > 1: uniform tile columns, no loop filtering across tiles
  int tile_rows = 1;         // > 1: uniform tile rows
  int wpp = 1;               // entropy_coding_sync_enabled_flag
  int scaling_list = 0;      // scaling_list_enabled_flag, default lists (no list data in the SPS / PPS)
  int conf_right = 0, conf_bottom = 0;   // conformance window: luma samples cropped at the right / bottom
};
void write_parameter_sets(const StreamLayout &l, std::vector<uint8_t> &out);
// layout of the stream a tiled encoder (b200_tiled_open*) writes
const StreamLayout &tiled_layout(void *tiled_encoder);
// Slice segment header with the entry points of `sub_len` (escaped sizes) followed by `data_len`
// bytes of already escaped slice data.
void write_slice_nal(const StreamLayout &l, bool idr, int poc, int qp, const uint32_t *sub_len, int n_sub,
                     const uint8_t *data, size_t data_len, std::vector<uint8_t> &out);

class Encoder {
 public:
  ~Encoder();
  bool open(const EncoderConfig &c);
  // Submit one packed I420 picture (host / device resident).  Returns false on error.  `au`
  // receives the next finished access unit in submission order, or stays empty while the
  // pipeline is filling.
  // `pinned` = the caller's buffer is page-locked and stays untouched until this picture's access
  // unit has been returned (kvz_api pictures from picture_alloc): it is then uploaded in place.
  bool encode_host(const uint8_t *i420, std::vector<uint8_t> &au, bool pinned = false);
  bool encode_device(const uint8_t *d_i420, std::vector<uint8_t> &au);
  // Columns [x0, x0 + width) of a packed I420 host picture that is pic_w wide (tile-column mode).
  bool encode_host_strip(const uint8_t *pic, int pic_w, int pic_h, int x0, int y0, std::vector<uint8_t> &au);
  // raw mode: substreams of the picture the last collect() returned (valid until its slot is reused)
  std::vector<uint32_t> last_sub_len;
  const uint8_t *last_data = nullptr;
  size_t last_data_len = 0;
  StreamLayout layout() const;
  // Drain: returns the next pending access unit (empty when none is left).
  bool flush(std::vector<uint8_t> &au);
  int pending() const { return (int)inflight.size(); }
  // QP of the pictures submitted from now on (frame-level rate control hook)
  void set_qp(int qp);
  // per-CTU QP offsets (raster, ctb_cols*ctb_rows) of the pictures submitted from now on; null clears
  bool set_ctu_dqp(const int8_t *dqp, int n);
  std::vector<int8_t> ctu_dqp;
  int qp() const { return cur_qp; }

  EncoderConfig cfg;
  FrameParams fp{};
  size_t frame_bytes = 0;
  int src_w = 0, src_h = 0;       // source picture size (= fp.w x fp.h unless a conformance window is in use)
  size_t src_bytes = 0;
  bool padded() const { return src_w != fp.w || src_h != fp.h; }
  bool stage_source(FrameSlot &s, const uint8_t *pic, cudaMemcpyKind kind, cudaStream_t st);
  uint32_t row_cap = 0, pack_cap = 0;
  int frame_idx = 0, poc = 0, cur = 0, last_idr = 0, last_qp = 0, last_poc = 0, cur_qp = 32;
  int idr_count = 0;              // IDR pictures submitted so far (vps_period)
  bool last_headers = false;      // the last returned access unit carries (tile mode: needs) the parameter sets
  unsigned long long last_bins = 0;
  std::vector<uint8_t> au;

  // Reconstruction ring: picture n is written to ring[n % kRecRing] and predicts from
  // ring[(n-1) % kRecRing].  More than two buffers let an IDR (which depends on nothing) run on
  // its own stream concurrently with the P pictures queued before it.
  static constexpr int kRecRing = 32;
  uint8_t *d_rec[kRecRing] = {}, *d_rec_pre = nullptr;
  uint8_t *d_scaling = nullptr;                     // scaling_list: the ScalingTable the kernels read (FrameParams::scaling)
  unsigned long long *d_me_stats = nullptr;         // profiling: work counters of k_me_ctu (FrameParams::me_stats)
  uint8_t *d_src_q = nullptr, *d_ref_q = nullptr;   // me_coarse: quarter-resolution source / previous reconstruction (main stream order)
  cudaEvent_t ev_ring[kRecRing] = {};    // "picture n finished reading its reference" (main stream)
  cudaStream_t intra_stream = nullptr, upload_stream = nullptr;
  cudaEvent_t ev_upload = nullptr;
  int *d_order = nullptr;                // CTU indices in wavefront order (intra kernel tickets)
  cudaEvent_t ev_intra = nullptr;
  uint8_t *last_rec() const { return d_rec[(frame_idx + kRecRing - 1) % kRecRing]; }
  std::vector<FrameSlot> slots;
  std::deque<int> inflight;          // slot indices, oldest first
  int last_slot = 0;
  size_t small_bytes = 0, off_flag = 0, off_prog = 0, off_ticket = 0, off_bins = 0, off_ctx = 0;
  cudaStream_t stream = nullptr;     // main (prediction chain) stream

  // per-kernel device time, measured with CUDA events on the launching stream (profile != 0)
  enum { K_INTRA = 0, K_ME, K_RECON, K_MODES, K_DEBLOCK, K_BINARISE, K_ARITH, K_PACK, K_SAO, K_COUNT };
  int profile = 0;
  cudaEvent_t ev_base = nullptr;      // time origin of the timeline below
  float timeline[2 * K_COUNT] = {};  // begin/end (ms since ev_base) of each kernel of the last collected picture
  double prof_ms[K_COUNT] = {};
  unsigned long long prof_cnt[K_COUNT] = {};

 private:
  void release();
  cudaStream_t input_stream() const;
  bool submit(FrameSlot &s, const uint8_t *d_i420);
  bool collect(FrameSlot &s, std::vector<uint8_t> &au);
};

}  // namespace b200
