// Host-side encoder object (internal header).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

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
};

class Encoder {
 public:
  ~Encoder();
  bool open(const EncoderConfig &c);
  // Encode one packed I420 picture (host / device resident).  `au` receives one Annex-B access unit.
  bool encode_host(const uint8_t *i420, std::vector<uint8_t> &au);
  bool encode_device(const uint8_t *d_i420, std::vector<uint8_t> &au);

  EncoderConfig cfg;
  FrameParams fp{};
  size_t frame_bytes = 0;
  uint32_t row_cap = 0;
  int frame_idx = 0, poc = 0, cur = 0, last_idr = 0;
  std::vector<uint8_t> au;

  // HBM-resident state
  uint8_t *d_src = nullptr, *d_rec[2] = {nullptr, nullptr}, *d_rec_pre = nullptr, *d_rows = nullptr;
  CuInfo *d_cu = nullptr;
  int16_t *d_levels = nullptr;
  uint8_t *d_small = nullptr;
  size_t small_bytes = 0, off_flag = 0, off_prog = 0, off_ticket = 0, off_bins = 0, off_ctx = 0;
  // pinned host staging
  uint8_t *h_src = nullptr, *h_rows = nullptr;
  uint32_t *h_small = nullptr;
  cudaStream_t stream = nullptr;

  uint32_t *d_row_len() const { return (uint32_t *)d_small; }
  int *d_sync_flag() const { return (int *)(d_small + off_flag); }
  int *d_progress() const { return (int *)(d_small + off_prog); }
  int *d_ticket() const { return (int *)(d_small + off_ticket); }
  unsigned long long *d_bins() const { return (unsigned long long *)(d_small + off_bins); }
  uint8_t *d_sync_ctx() const { return d_small + off_ctx; }

 private:
  void release();
  void write_parameter_sets(std::vector<uint8_t> &out) const;
  void write_slice(std::vector<uint8_t> &out, bool idr, const uint32_t *row_len) const;
};

}  // namespace b200
