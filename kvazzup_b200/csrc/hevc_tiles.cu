// HEVC tiles (uniform columns x rows) as independent encoders ("4K call, tiles").
//
// A tile whose motion vectors never reach across its interior edges (Kvazaar's
// mv-constraint=frametilemargin, which the reference exposes, kvazaarfilter.cpp:246-276) and that is
// not loop-filtered across those edges is coded exactly like a picture of its own: neighbours in
// another tile are unavailable just as they are beyond a picture edge, CABAC restarts at the tile, and
// the slice data is the tiles' substreams back to back behind one slice header.  So the tiled encoder
// is T ordinary encoders -- each may sit on its own GPU, and nothing is exchanged between them
// (option A of SURVEY 8e) -- plus this compositor, which writes the parameter sets (PPS with uniform
// tile columns, loop_filter_across_tiles off) and the slice header with all entry points.
//
// HEVC Main allows tiles or WPP in a picture, not both, and FFmpeg's decoder (the independent pin
// of this repository) follows that; `wpp = 0` (one substream per tile) is therefore the mode whose
// streams are verified externally.  `wpp = 1` keeps one substream per CTU row inside every tile --
// what Kvazaar emits with --tiles and --wpp and what keeps the entropy coder parallel -- and is
// checked against the oracle only.
#include <string.h>

#include <algorithm>
#include <memory>
#include <string>
#include <thread>

#include "../../include/b200_hevc.h"
#include "../../include/b200media.h"
#include "hevc_encoder.h"
#include "runtime.h"

namespace b200 {

struct TiledEncoder {
  struct Strip { std::unique_ptr<Encoder> enc; int x0 = 0, y0 = 0, wd = 0, ht = 0, device = 0; };
  std::vector<Strip> strips;
  StreamLayout layout;
  int width = 0, height = 0, pending_pics = 0;
  std::vector<uint8_t> au, tmp, data;
  std::vector<uint32_t> sub_len;

  bool open(const EncoderConfig &c, int tile_cols, int tile_rows, int wpp, const int *devices, int n_devices)
  {
    const int ctb_cols = (c.width + kCtb - 1) / kCtb, ctb_rows = (c.height + kCtb - 1) / kCtb;
    if (tile_cols < 1 || (tile_cols > 1 && tile_cols > ctb_cols / 2)) { set_error("tiled encoder: %d tile columns for %d CTU columns (each tile must be at least two CTUs wide)", tile_cols, ctb_cols); return false; }
    if (tile_rows < 1 || tile_rows > ctb_rows || tile_cols * tile_rows > 64) { set_error("tiled encoder: %d tile rows for %d CTU rows", tile_rows, ctb_rows); return false; }
    const int tiles = tile_cols * tile_rows;
    if (c.qp_delta) { set_error("tiled encoder: per-CTU QP is not available together with tiles"); return false; }
    width = c.width; height = c.height;
    layout.w = c.width; layout.h = c.height; layout.deblock = c.deblock; layout.qp_delta = 0;
    layout.tile_cols = tile_cols; layout.tile_rows = tile_rows; layout.wpp = wpp ? 1 : 0;
    layout.fps_num = c.fps_num; layout.fps_den = c.fps_den; layout.sao = c.sao; layout.scaling_list = c.scaling_list;
    int prev = 0;
    cudaGetDevice(&prev);
    strips.resize(tiles);
    for (int i = 0; i < tiles; i++) {                                              // tiles in raster order (6.5.1)
      Strip &s = strips[i];
      const int tc = i % tile_cols, tr = i / tile_cols;
      const int c0 = tc * ctb_cols / tile_cols, c1 = (tc + 1) * ctb_cols / tile_cols;   // colBd of uniform spacing
      const int r0 = tr * ctb_rows / tile_rows, r1 = (tr + 1) * ctb_rows / tile_rows;   // rowBd
      s.x0 = c0 * kCtb;
      s.wd = std::min(c.width, c1 * kCtb) - s.x0;
      s.y0 = r0 * kCtb;
      s.ht = std::min(c.height, r1 * kCtb) - s.y0;
      s.device = n_devices > 0 ? devices[i % n_devices] : prev;
      EncoderConfig sc = c;
      sc.width = s.wd; sc.height = s.ht;
      sc.mv_edges = (tc > 0 ? 1 : 0) | (tc < tile_cols - 1 ? 2 : 0) | (tr > 0 ? 4 : 0) | (tr < tile_rows - 1 ? 8 : 0) | (c.mv_edges & 15);   // interior tile edges + the constrained picture edges
      sc.more_tiles = i < tiles - 1 ? 1 : 0;
      sc.no_wpp = wpp ? 0 : 1;
      sc.raw = 1;
      if (cudaSetDevice(s.device) != cudaSuccess) { set_error("tiled encoder: cannot select CUDA device %d", s.device); cudaSetDevice(prev); return false; }
      s.enc.reset(new Encoder());
      if (!s.enc->open(sc)) { cudaSetDevice(prev); return false; }
    }
    cudaSetDevice(prev);
    return true;
  }

  // Assemble the access unit of the picture every strip has just handed back.
  void compose()
  {
    sub_len.clear();
    data.clear();
    for (Strip &s : strips) {
      sub_len.insert(sub_len.end(), s.enc->last_sub_len.begin(), s.enc->last_sub_len.end());
      data.insert(data.end(), s.enc->last_data, s.enc->last_data + s.enc->last_data_len);
    }
    const Encoder &e0 = *strips[0].enc;
    au.clear();
    if (e0.last_headers) write_parameter_sets(layout, au);
    write_slice_nal(layout, e0.last_idr != 0, e0.last_poc, e0.last_qp, sub_len.data(), (int)sub_len.size(), data.data(), data.size(), au);
  }

  // pic == nullptr drains.  au is empty while the strips' pipelines fill.  Every strip is driven by
  // its own host thread for the duration of the call: the per-picture host work (three strided
  // uploads, a dozen launches, one event wait) then overlaps across strips and GPUs instead of
  // adding up on the caller's thread.
  bool step(const uint8_t *pic)
  {
    const int n = (int)strips.size();
    std::vector<int> ok(n, 0), ready(n, 0);
    std::vector<std::string> err(n);
    auto work = [&](int i) {
      Strip &s = strips[i];
      std::vector<uint8_t> got;
      if (cudaSetDevice(s.device) != cudaSuccess) { err[i] = "cannot select the strip's CUDA device"; return; }
      ok[i] = pic ? s.enc->encode_host_strip(pic, width, height, s.x0, s.y0, got) : s.enc->flush(got);
      if (!ok[i]) err[i] = b200_last_error();
      ready[i] = got.empty() ? 0 : 1;
    };
    if (n == 1) {
      int prev = 0;
      cudaGetDevice(&prev);
      work(0);
      cudaSetDevice(prev);
    } else {
      std::vector<std::thread> th;
      th.reserve(n);
      for (int i = 0; i < n; i++) th.emplace_back(work, i);
      for (std::thread &t : th) t.join();
    }
    au.clear();
    int n_ready = 0;
    for (int i = 0; i < n; i++) {
      if (!ok[i]) { set_error("tiled encoder, strip %d: %s", i, err[i].c_str()); return false; }
      n_ready += ready[i];
    }
    if (n_ready == 0) return true;
    if (n_ready != n) { set_error("tiled encoder: strips out of step"); return false; }   // they run in lock step
    compose();
    return true;
  }

  int pending() const { return strips.empty() ? 0 : strips[0].enc->pending(); }
};

}  // namespace b200

using b200::TiledEncoder;

namespace b200 {
const StreamLayout &tiled_layout(void *tiled_encoder) { return ((TiledEncoder *)tiled_encoder)->layout; }
}  // namespace b200

extern "C" {

void *b200_tiled_open(int width, int height, int qp, int intra_period, int search_range, int deblock, int depth,
                      int tile_cols, int wpp, const int *devices, int n_devices)
{
  b200::EncoderConfig c;
  c.width = width; c.height = height; c.qp = qp; c.intra_period = intra_period; c.search_range = search_range;
  c.deblock = deblock; c.depth = depth; c.debug = 0;
  TiledEncoder *t = new TiledEncoder();
  if (!t->open(c, tile_cols, 1, wpp, devices, n_devices)) { delete t; return nullptr; }
  return t;
}

void b200_tiled_params_default(b200_tiled_params *p)
{
  if (!p) return;
  memset(p, 0, sizeof(*p));
  p->struct_size = (int)sizeof(*p);
  p->qp = 32; p->intra_period = 64; p->search_range = 8; p->deblock = 1; p->depth = 1; p->tile_cols = 1; p->tile_rows = 1; p->vps_period = 1;
}

void *b200_tiled_open_params(const b200_tiled_params *up, const int *devices, int n_devices)
{
  if (!up || up->struct_size < (int)(11 * sizeof(int))) { b200::set_error("b200_tiled_open_params: bad arguments"); return nullptr; }
  b200_tiled_params p;
  b200_tiled_params_default(&p);
  memcpy(&p, up, std::min<size_t>((size_t)up->struct_size, sizeof(p)));
  b200::EncoderConfig c;
  c.width = p.width; c.height = p.height; c.qp = p.qp; c.intra_period = p.intra_period; c.search_range = p.search_range;
  c.deblock = p.deblock; c.depth = p.depth; c.debug = 0; c.fps_num = p.fps_num; c.fps_den = p.fps_den; c.sao = p.sao; c.intra_in_p = p.intra_in_p; c.me_coarse = p.me_coarse; c.intra_satd = p.intra_satd; c.subme_satd = p.subme_satd; c.scaling_list = p.scaling_list ? 1 : 0; c.mv_edges = p.mv_edges & 15; c.vps_period = p.vps_period;
  TiledEncoder *t = new TiledEncoder();
  if (!t->open(c, p.tile_cols, p.tile_rows < 1 ? 1 : p.tile_rows, p.wpp, devices, n_devices)) { delete t; return nullptr; }
  return t;
}

void b200_tiled_close(void *h) { delete (TiledEncoder *)h; }

// VUI timing info of the parameter sets written from now on (kvz_api: "input-fps")
void b200_tiled_set_fps(void *h, int fps_num, int fps_den)
{
  TiledEncoder *t = (TiledEncoder *)h;
  if (t) { t->layout.fps_num = fps_num; t->layout.fps_den = fps_den; }
}

static int tiled_finish(TiledEncoder *t, uint8_t *out, int cap)
{
  if (t->au.empty()) return 0;
  if ((size_t)cap < t->au.size()) { b200::set_error("tiled encoder: output buffer too small (%zu needed)", t->au.size()); return -(int)t->au.size(); }
  memcpy(out, t->au.data(), t->au.size());
  return (int)t->au.size();
}

int b200_tiled_encode(void *h, const uint8_t *i420, uint8_t *out, int cap)
{
  TiledEncoder *t = (TiledEncoder *)h;
  if (!t || !i420 || !out) { b200::set_error("b200_tiled_encode: bad arguments"); return B200_ERR_ARG; }
  if (!t->step(i420)) return B200_ERR_CUDA;
  return tiled_finish(t, out, cap);
}

int b200_tiled_flush(void *h, uint8_t *out, int cap)
{
  TiledEncoder *t = (TiledEncoder *)h;
  if (!t || !out) { b200::set_error("b200_tiled_flush: bad arguments"); return B200_ERR_ARG; }
  if (!t->step(nullptr)) return B200_ERR_CUDA;
  return tiled_finish(t, out, cap);
}

int b200_tiled_pending(void *h) { return h ? ((TiledEncoder *)h)->pending() : 0; }

// Reconstruction of the last submitted picture (depth 1 only), packed I420 of the whole picture.
int b200_tiled_recon(void *h, uint8_t *dst, size_t cap)
{
  TiledEncoder *t = (TiledEncoder *)h;
  const size_t ysz = t ? (size_t)t->width * t->height : 0;
  if (!t || !dst || cap < ysz * 3 / 2) { b200::set_error("b200_tiled_recon: bad arguments"); return B200_ERR_ARG; }
  int prev = 0;
  cudaGetDevice(&prev);
  for (TiledEncoder::Strip &s : t->strips) {
    cudaSetDevice(s.device);
    b200::Encoder &e = *s.enc;
    cudaStreamSynchronize(e.stream);
    const uint8_t *rec = e.last_rec();
    const size_t sy = (size_t)s.wd * s.ht, cw = t->width / 2;
    cudaMemcpy2D(dst + (size_t)s.y0 * t->width + s.x0, t->width, rec, s.wd, s.wd, s.ht, cudaMemcpyDeviceToHost);
    cudaMemcpy2D(dst + ysz + (size_t)(s.y0 / 2) * cw + s.x0 / 2, cw, rec + sy, s.wd / 2, s.wd / 2, s.ht / 2, cudaMemcpyDeviceToHost);
    cudaMemcpy2D(dst + ysz + ysz / 4 + (size_t)(s.y0 / 2) * cw + s.x0 / 2, cw, rec + sy + sy / 4, s.wd / 2, s.wd / 2, s.ht / 2, cudaMemcpyDeviceToHost);
  }
  cudaSetDevice(prev);
  return B200_OK;
}

}  // extern "C"
