// kvz_api boundary (include/b200_kvazaar.h): the C ABI KvazaarFilter binds
// (reference src/media/processing/kvazaarfilter.cpp:145-318, 374-484), implemented on the B200
// encoder engine.  Option names follow Kvazaar's config_parse; presets map to the search range.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <deque>
#include <new>

#include "../../include/b200_kvazaar.h"
#include "../../include/b200_hevc.h"
#include "hevc_encoder.h"
#include "runtime.h"

using b200::Encoder;
using b200::EncoderConfig;

struct kvz_encoder {
  Encoder eng;
  void *tiled = nullptr;                // b200_tiled_* handle when cfg.tiles_width_count > 1 (eng stays closed)
  std::vector<uint8_t> tiled_out;
  ~kvz_encoder() { if (tiled) b200_tiled_close(tiled); }
  kvz_config cfg;
  std::deque<int64_t> pts;              // presentation timestamps of the pictures in flight
  // Frame-level rate control in the lambda domain (target_bitrate != 0, rc-algorithm lambda; K9 of
  // SURVEY.md 8a): bits per pixel of a picture <-> lambda through R = (lambda / alpha)^(1 / beta), the
  // model Kvazaar's (and HM's, JCTVC-K0103) lambda-domain control uses, one (alpha, beta) pair per
  // picture type, updated from every coded picture; the target of a picture comes from a sliding
  // window over the bits spent so far.
  struct RcModel { double alpha = 3.2003, beta = -1.367; };
  RcModel rc_model[2];                  // [0] P pictures, [1] IDR pictures
  double bits_per_frame = 0;            // target_bitrate / frame rate
  double rc_spent = 0;                  // bits of the pictures returned so far
  long long rc_returned = 0, rc_submitted = 0;
  std::deque<std::pair<int, bool>> rc_pending;   // (QP, IDR) of the pictures in flight, oldest first
  int rc_last_qp[2] = {-1, -1};
  std::vector<uint8_t> au;
};

namespace {

// What a preset selects.  Restated from the intent of Kvazaar's preset table (cfg.c; not verified against
// 2.3.1, which is absent from this image): SAO is off in ultrafast and on ("full") from superfast up;
// faster presets search less.  Here "search" is the two-level motion search of k_me_ctu: me_coarse =
// range of the coarse level in 4x4-mean samples (16 = +-64 luma samples), me_range = window searched
// around the zero vector and around each 32x32 block's coarse vector.
// intra_satd: the intra mode search of I pictures compares SATD (what Kvazaar's rough search does) from superfast up.
// subme_satd: SATD in the fractional motion refinement as well, from fast up (-0.3 ... -0.5 % BD-rate for about a
// fifth of the encoder's throughput: the presets quoted for speed keep the SAD).
struct Preset { const char *name; int me_range; int sao; int me_coarse; int intra_satd; int subme_satd; };
const Preset kPresets[] = {{"ultrafast", 4, 0, 16, 0, 0}, {"superfast", 4, 3, 16, 1, 0}, {"veryfast", 6, 3, 16, 1, 0}, {"faster", 8, 3, 16, 1, 0},
                           {"fast", 8, 3, 32, 1, 1},      {"medium", 12, 3, 32, 1, 1},   {"slow", 16, 3, 32, 1, 1},    {"slower", 16, 3, 32, 1, 1},
                           {"veryslow", 16, 3, 32, 1, 1}, {"placebo", 16, 3, 32, 1, 1}};

int parse_int(const char *v, int *out)
{
  if (!v || !*v) return 0;
  char *end = nullptr;
  long x = strtol(v, &end, 10);
  if (*end) return 0;
  *out = (int)x;
  return 1;
}

int parse_bool(const char *v, int *out)
{
  if (!v || !*v) { *out = 1; return 1; }
  if (!strcmp(v, "1") || !strcmp(v, "true") || !strcmp(v, "on") || !strcmp(v, "yes")) { *out = 1; return 1; }
  if (!strcmp(v, "0") || !strcmp(v, "false") || !strcmp(v, "off") || !strcmp(v, "no")) { *out = 0; return 1; }
  return 0;
}

kvz_config *config_alloc(void) { return (kvz_config *)calloc(1, sizeof(kvz_config)); }

int config_destroy(kvz_config *cfg) { free(cfg); return 1; }

int config_init(kvz_config *cfg)
{
  if (!cfg) return 0;
  memset(cfg, 0, sizeof(*cfg));
  cfg->framerate_num = 25; cfg->framerate_denom = 1;
  cfg->qp = 22;                    // Kvazaar's default; the reference always overrides it (:219)
  cfg->intra_period = 64; cfg->vps_period = 0;
  cfg->wpp = 1; cfg->owf = 0; cfg->threads = 0;
  cfg->deblock_enable = 1; cfg->sao_type = 3;    // preset veryfast
  cfg->tiles_width_count = 1; cfg->tiles_height_count = 1;
  cfg->gop_lowdelay = 1; cfg->gop_len = 4;
  cfg->me_range = 6; cfg->me_coarse = 16; cfg->intra_satd = 1; cfg->device = -1;     // preset veryfast
  snprintf(cfg->preset, sizeof(cfg->preset), "veryfast");
  return 1;
}

// Options Kvazaar knows that do not change what this encoder does: accepted so that the
// reference's "parameters" array (kvazaarfilter.cpp:351-371) does not log spurious warnings.
const char *const kIgnored[] = {"rd", "rdoq", "rdoq-skip", "signhide", "smp", "amp", "subme", "me", "me-steps",
  "pu-depth-inter", "pu-depth-intra", "tr-depth-intra", "bipred", "ref", "transform-skip", "full-intra-search",
  "cu-split-termination", "me-early-termination", "intra-rdo-et", "early-skip", "fast-residual-cost", "max-merge",
  "cqmfile", "erp-aqp", "level", "force-level", "high-tier", "implicit-rdpcm", "mrl", "tmvp", "open-gop",
  "input-bitdepth", "input-format", "aud", "psnr", "info", "cpuid", "roi", "fastrd-sampling", "fastrd-accuracy-check",
  "intra-qp-offset", "intra-bits", "clip-neighbour", "partial-coding", "zero-coeff-rdo", "combine-intra-cus", NULL};

int config_parse(kvz_config *cfg, const char *name, const char *value)
{
  if (!cfg || !name) return 0;
  int v = 0;
  if (!strncmp(name, "--", 2)) name += 2;
  if (!strcmp(name, "preset")) {
    for (const Preset &p : kPresets)
      if (value && !strcmp(value, p.name)) { cfg->me_range = p.me_range; cfg->sao_type = p.sao; cfg->me_coarse = p.me_coarse; cfg->intra_satd = p.intra_satd; cfg->subme_satd = p.subme_satd; snprintf(cfg->preset, sizeof(cfg->preset), "%s", p.name); return 1; }
    return 0;
  }
  if (!strcmp(name, "input-res")) {
    int w = 0, h = 0;
    if (!value || sscanf(value, "%dx%d", &w, &h) != 2 || w <= 0 || h <= 0) return 0;
    cfg->width = w; cfg->height = h;
    return 1;
  }
  if (!strcmp(name, "input-fps")) {
    int n = 0, d = 1;
    if (!value) return 0;
    if (sscanf(value, "%d/%d", &n, &d) == 2 && n > 0 && d > 0) { cfg->framerate_num = n; cfg->framerate_denom = d; return 1; }
    double f = atof(value);
    if (f <= 0) return 0;
    cfg->framerate_num = (int)(f * 1000 + 0.5); cfg->framerate_denom = 1000;
    return 1;
  }
  if (!strcmp(name, "qp")) { if (!parse_int(value, &v) || v < 0 || v > 51) return 0; cfg->qp = v; return 1; }
  if (!strcmp(name, "period")) { if (!parse_int(value, &v) || v < 0) return 0; cfg->intra_period = v; return 1; }
  if (!strcmp(name, "vps-period")) { if (!parse_int(value, &v) || v < 0) return 0; cfg->vps_period = v; return 1; }
  if (!strcmp(name, "threads")) { if (value && !strcmp(value, "auto")) { cfg->threads = -1; return 1; } if (!parse_int(value, &v) || v < 0) return 0; cfg->threads = v; return 1; }
  if (!strcmp(name, "owf")) { if (value && !strcmp(value, "auto")) { cfg->owf = 3; return 1; } if (!parse_int(value, &v) || v < 0 || v > 127) return 0; cfg->owf = v; return 1; }
  if (!strcmp(name, "wpp")) { if (!parse_bool(value, &v)) return 0; cfg->wpp = v; return 1; }
  if (!strcmp(name, "no-wpp")) { cfg->wpp = 0; return 1; }
  if (!strcmp(name, "tiles")) {
    int c = 0, r = 0;
    if (!value || sscanf(value, "%dx%d", &c, &r) != 2) return 0;
    if (c < 1 || c > 32 || r < 1 || r > 32 || c * r > 64) return 0;   // uniform tile grid, see hevc_tiles.cu
    cfg->tiles_width_count = c; cfg->tiles_height_count = r;
    return 1;
  }
  if (!strcmp(name, "slices")) { return 0; }   // one slice per picture (the reference notes slices break uvgRTP, :204)
  if (!strcmp(name, "bitrate")) { if (!parse_int(value, &v) || v < 0) return 0; cfg->target_bitrate = v; return 1; }
  if (!strcmp(name, "rc-algorithm")) {
    if (!value) return 0;
    if (!strcmp(value, "no-rc")) { cfg->rc_algorithm = KVZ_NO_RC; return 1; }
    if (!strcmp(value, "lambda")) { cfg->rc_algorithm = KVZ_LAMBDA; return 1; }
    // "oba" (Kvazaar's optimal bit allocation) is not built: refused, the reference then logs a warning
    // and the lambda-domain control stays in force (kvazaarfilter.cpp:363-369)
    return 0;
  }
  if (!strcmp(name, "gop")) {
    // low-delay P only: "lp-g<len>d<depth>t<layers>" (the reference hard-codes lp-g4d3t1, :233) or 0
    if (!value) return 0;
    if (!strcmp(value, "0")) { cfg->gop_lowdelay = 1; return 1; }
    if (!strncmp(value, "lp-", 3)) { cfg->gop_lowdelay = 1; int g = 4; sscanf(value, "lp-g%d", &g); cfg->gop_len = g; return 1; }
    return 0;
  }
  if (!strcmp(name, "scaling-list")) {        // off / default (the standard's default lists); custom list files are not read
    if (value && !strcmp(value, "off")) { cfg->scaling_list = 0; return 1; }
    if (value && !strcmp(value, "default")) { cfg->scaling_list = 1; return 1; }
    return 0;
  }
  if (!strcmp(name, "mv-constraint")) {
    if (!value || !*value || !strcmp(value, "none")) { cfg->mv_constraint = KVZ_MV_CONSTRAIN_NONE; return 1; }
    if (!strcmp(value, "frame")) { cfg->mv_constraint = KVZ_MV_CONSTRAIN_FRAME; return 1; }
    if (!strcmp(value, "tile")) { cfg->mv_constraint = KVZ_MV_CONSTRAIN_TILE; return 1; }
    if (!strcmp(value, "frametile")) { cfg->mv_constraint = KVZ_MV_CONSTRAIN_FRAME_AND_TILE; return 1; }
    if (!strcmp(value, "frametilemargin")) { cfg->mv_constraint = KVZ_MV_CONSTRAIN_FRAME_AND_TILE_MARGIN; return 1; }
    return 0;
  }
  if (!strcmp(name, "vaq")) { if (!parse_int(value, &v) || v < 0 || v > 20) return 0; cfg->vaq = v; return 1; }
  if (!strcmp(name, "deblock")) {
    if (!value || !*value) { cfg->deblock_enable = 1; return 1; }
    if (parse_bool(value, &v)) { cfg->deblock_enable = v; return 1; }
    int b = 0, t = 0;
    if (sscanf(value, "%d:%d", &b, &t) == 2 && b == 0 && t == 0) { cfg->deblock_enable = 1; return 1; }
    return 0;
  }
  if (!strcmp(name, "no-deblock")) { cfg->deblock_enable = 0; return 1; }
  if (!strcmp(name, "sao")) {
    // Kvazaar: off / edge / band / full.  Edge and band offsets are always decided together here, so
    // any value but "off" switches both on.
    if (!value || !*value) { cfg->sao_type = 3; return 1; }
    if (!strcmp(value, "off") || !strcmp(value, "0")) { cfg->sao_type = 0; return 1; }
    if (!strcmp(value, "edge")) { cfg->sao_type = 1; return 1; }
    if (!strcmp(value, "band")) { cfg->sao_type = 2; return 1; }
    if (!strcmp(value, "full") || !strcmp(value, "1")) { cfg->sao_type = 3; return 1; }
    return 0;
  }
  if (!strcmp(name, "no-sao")) { cfg->sao_type = 0; return 1; }
  if (!strcmp(name, "lossless")) { if (!parse_bool(value, &v)) return 0; cfg->lossless = v; return 1; }
  if (!strcmp(name, "hash")) {
    if (!value) return 0;
    if (!strcmp(value, "none")) { cfg->hash = KVZ_HASH_NONE; return 1; }
    return 0;
  }
  if (!strcmp(name, "set-qp-in-cu")) { if (!parse_bool(value, &v)) return 0; cfg->set_qp_in_cu = v; return 1; }
  if (!strcmp(name, "b200-me-range")) { if (!parse_int(value, &v) || v < 1 || v > 32) return 0; cfg->me_range = v; return 1; }
  if (!strcmp(name, "b200-subme-satd")) { if (!parse_bool(value, &v)) return 0; cfg->subme_satd = v; return 1; }
  if (!strcmp(name, "b200-intra-satd")) { if (!parse_bool(value, &v)) return 0; cfg->intra_satd = v; return 1; }
  if (!strcmp(name, "b200-me-coarse")) { if (!parse_int(value, &v) || v < 0 || v > 32 || (v & 3)) return 0; cfg->me_coarse = v; return 1; }
  if (!strcmp(name, "b200-recon")) { if (!parse_bool(value, &v)) return 0; cfg->return_recon = v; return 1; }
  if (!strcmp(name, "b200-roi")) { if (!parse_bool(value, &v)) return 0; cfg->roi_enable = v; return 1; }
  if (!strcmp(name, "b200-device")) { if (!parse_int(value, &v)) return 0; cfg->device = v; return 1; }
  for (int i = 0; kIgnored[i]; i++)
    if (!strcmp(name, kIgnored[i])) return 1;
  return 0;
}

kvz_picture *picture_alloc_csp(enum kvz_chroma_format csp, int32_t w, int32_t h)
{
  if (csp != KVZ_CSP_420 || w <= 0 || h <= 0 || (w & 1) || (h & 1)) return NULL;
  kvz_picture *p = (kvz_picture *)calloc(1, sizeof(kvz_picture));
  if (!p) return NULL;
  size_t ysz = (size_t)w * h;
  // page-locked when a CUDA device is present, so that encoder_encode can upload the picture in
  // place (the reference keeps a ring of owf + 1 pictures and does not touch a picture again before
  // its access unit came back, kvazaarfilter.cpp:76-88,299); plain memory otherwise
  p->base_image = NULL;
  if (cudaHostAlloc((void **)&p->fulldata_buf, ysz + ysz / 2, cudaHostAllocDefault) == cudaSuccess) {
    p->base_image = p;                         // marks "page-locked, allocated by us"
  } else {
    cudaGetLastError();
    p->fulldata_buf = (kvz_pixel *)malloc(ysz + ysz / 2);
  }
  if (!p->fulldata_buf) { free(p); return NULL; }
  p->fulldata = p->fulldata_buf;
  p->y = p->data[0] = p->fulldata;
  p->u = p->data[1] = p->fulldata + ysz;
  p->v = p->data[2] = p->fulldata + ysz + ysz / 4;
  p->width = w; p->height = h; p->stride = w;
  p->refcount = 1; p->chroma_format = csp;
  return p;
}

kvz_picture *picture_alloc(int32_t w, int32_t h) { return picture_alloc_csp(KVZ_CSP_420, w, h); }

void picture_free(kvz_picture *p)
{
  if (!p) return;
  if (--p->refcount > 0) return;
  if (p->base_image == p) cudaFreeHost(p->fulldata_buf);
  else free(p->fulldata_buf);
  free(p);
}

void chunk_free(kvz_data_chunk *c)
{
  while (c) { kvz_data_chunk *n = c->next; free(c); c = n; }
}

kvz_data_chunk *to_chunks(const std::vector<uint8_t> &au)
{
  kvz_data_chunk *head = NULL, **tail = &head;
  for (size_t off = 0; off < au.size(); off += KVZ_DATA_CHUNK_SIZE) {
    kvz_data_chunk *c = (kvz_data_chunk *)malloc(sizeof(kvz_data_chunk));
    if (!c) { chunk_free(head); return NULL; }
    c->len = (uint32_t)std::min<size_t>(KVZ_DATA_CHUNK_SIZE, au.size() - off);
    memcpy(c->data, au.data() + off, c->len);
    c->next = NULL;
    *tail = c;
    tail = &c->next;
  }
  return head;
}

kvz_encoder *encoder_open(const kvz_config *cfg)
{
  if (!cfg) { b200::set_error("encoder_open: NULL config"); return NULL; }
  if (cfg->lossless) { b200::set_error("encoder_open: lossless coding is not supported"); return NULL; }
  if (cfg->device >= 0 && cudaSetDevice(cfg->device) != cudaSuccess) { b200::set_error("encoder_open: cannot select CUDA device %d", cfg->device); return NULL; }
  kvz_encoder *e = new (std::nothrow) kvz_encoder();
  if (!e) return NULL;
  e->cfg = *cfg;
  EncoderConfig c;
  // sizes that are not multiples of 8 are coded padded, with a conformance window (what Kvazaar does)
  if (cfg->width <= 0 || cfg->height <= 0 || (cfg->width & 1) || (cfg->height & 1)) {
    b200::set_error("encoder_open: width and height must be positive and even (got %dx%d)", cfg->width, cfg->height);
    delete e;
    return NULL;
  }
  c.width = (cfg->width + 7) & ~7; c.height = (cfg->height + 7) & ~7;
  if (c.width != cfg->width || c.height != cfg->height) { c.src_width = cfg->width; c.src_height = cfg->height; }
  c.qp = cfg->qp; c.intra_period = cfg->intra_period;
  c.search_range = cfg->me_range > 0 ? cfg->me_range : 6;
  c.me_coarse = cfg->me_coarse;
  c.intra_satd = cfg->intra_satd; c.subme_satd = cfg->subme_satd;
  if (c.me_coarse > 0 && c.search_range > 16) c.search_range = 16;
  c.deblock = cfg->deblock_enable; c.debug = 0; c.depth = cfg->owf + 1;
  const bool tiled = cfg->tiles_width_count > 1 || cfg->tiles_height_count > 1;
  c.vaq = tiled ? 0 : cfg->vaq;                   // per-CTU QP is not available together with tiles
  c.scaling_list = cfg->scaling_list ? 1 : 0;
  c.vps_period = cfg->vps_period;
  // mv-constraint frame / frametile / frametilemargin: no motion vector leaves the picture (tiles confine motion to the
  // tile in any case, see below)
  const bool mv_frame = cfg->mv_constraint == KVZ_MV_CONSTRAIN_FRAME || cfg->mv_constraint == KVZ_MV_CONSTRAIN_FRAME_AND_TILE ||
                        cfg->mv_constraint == KVZ_MV_CONSTRAIN_FRAME_AND_TILE_MARGIN;
  if (mv_frame) c.mv_edges = 15;
  c.qp_delta = (cfg->roi_enable || cfg->set_qp_in_cu || c.vaq) ? 1 : 0;
  c.fps_num = cfg->framerate_num; c.fps_den = cfg->framerate_denom;     // VUI timing: the decoder side reports it
  c.sao = cfg->sao_type != 0 ? 2 : 0;             // with sao_merge_left / _up flags
  c.intra_in_p = 1;                               // every Kvazaar preset may code intra CUs in P pictures
  if (tiled && (c.src_width || c.src_height)) {
    b200::set_error("encoder_open: tiles need a picture size that is a multiple of 8 (got %dx%d)", cfg->width, cfg->height);
    delete e;
    return NULL;
  }
  if (tiled) {
    // tiles: independent tile encoders on this GPU; motion is confined to the tile, like
    // Kvazaar's mv-constraint frametilemargin (the reference exposes it, kvazaarfilter.cpp:246-276);
    // constant QP only (no ROI, no rate control)
    b200_tiled_params tp;
    b200_tiled_params_default(&tp);
    tp.width = c.width; tp.height = c.height; tp.qp = c.qp; tp.intra_period = c.intra_period; tp.search_range = c.search_range;
    tp.deblock = c.deblock; tp.depth = c.depth; tp.tile_cols = cfg->tiles_width_count; tp.tile_rows = cfg->tiles_height_count; tp.wpp = cfg->wpp ? 1 : 0;
    tp.fps_num = c.fps_num; tp.fps_den = c.fps_den; tp.sao = c.sao; tp.intra_in_p = c.intra_in_p; tp.me_coarse = c.me_coarse; tp.intra_satd = c.intra_satd; tp.subme_satd = c.subme_satd; tp.scaling_list = c.scaling_list; tp.mv_edges = c.mv_edges; tp.vps_period = c.vps_period;
    e->tiled = b200_tiled_open_params(&tp, nullptr, 0);
    if (!e->tiled) { delete e; return NULL; }
    e->tiled_out.resize((size_t)c.width * c.height * 3 + 65536);
    return e;
  }
  if (!e->eng.open(c)) { delete e; return NULL; }
  if (cfg->target_bitrate > 0 && cfg->framerate_num > 0) {
    e->bits_per_frame = (double)cfg->target_bitrate * cfg->framerate_denom / cfg->framerate_num;
  }
  return e;
}

void encoder_close(kvz_encoder *e) { delete e; }

int encoder_headers(kvz_encoder *e, kvz_data_chunk **data_out, uint32_t *len_out)
{
  // VPS, SPS and PPS as one chunk list.  They also travel in-band: before the first picture and before every
  // vps-period-th IDR picture after it (vps-period 1 in the reference, :221)
  if (!e) { b200::set_error("encoder_headers: NULL encoder"); return 0; }
  std::vector<uint8_t> ps;
  b200::write_parameter_sets(e->tiled ? b200::tiled_layout(e->tiled) : e->eng.layout(), ps);
  if (data_out) {
    *data_out = to_chunks(ps);
    if (!*data_out) { b200::set_error("encoder_headers: out of memory"); return 0; }
  }
  if (len_out) *len_out = (uint32_t)ps.size();
  return 1;
}

// ---- lambda-domain rate control ----------------------------------------------------------------
// QP <-> lambda: QP = 4.2005 ln(lambda) + 13.7122 (JCTVC-K0103, the relation HM and Kvazaar use)
double rc_lambda_of_qp(int qp) { return exp((qp - 13.7122) / 4.2005); }

constexpr int kRcWindow = 40;           // pictures over which a surplus / deficit is worked off
constexpr double kRcIntraShare = 6.0;   // an IDR picture may take this many average pictures' worth of bits

// QP of the picture about to be submitted.
int rate_control_pick_qp(kvz_encoder *e, bool idr)
{
  const double px = (double)e->cfg.width * e->cfg.height;
  // bits spent by pictures still in flight are not known yet: count them at the average
  const double spent = e->rc_spent + (double)(e->rc_submitted - e->rc_returned) * e->bits_per_frame;
  const double n = (double)e->rc_submitted;
  double target = (e->bits_per_frame * (n + kRcWindow) - spent) / kRcWindow;         // sliding-window allocation
  target = std::min(std::max(target, 0.25 * e->bits_per_frame), 4.0 * e->bits_per_frame);
  // an intra period of P pictures holds one IDR worth kRcIntraShare average pictures: the P pictures share the rest
  const double period = (double)e->cfg.intra_period;
  if (idr) target *= kRcIntraShare;
  else if (period > kRcIntraShare + 1) target *= (period - kRcIntraShare) / (period - 1.0);
  if (target < 64) target = 64;
  const kvz_encoder::RcModel &m = e->rc_model[idr ? 1 : 0];
  const double bpp = target / px;
  double lambda = m.alpha * pow(bpp, m.beta);
  lambda = std::min(std::max(lambda, 0.1), 10000.0);
  int qp = (int)floor(4.2005 * log(lambda) + 13.7122 + 0.5);
  const int last = e->rc_last_qp[idr ? 1 : 0];
  if (last >= 0) qp = std::min(std::max(qp, last - 3), last + 3);                      // no jumps between pictures of a kind
  else if (idr && e->rc_last_qp[0] >= 0) qp = std::min(std::max(qp, e->rc_last_qp[0] - 6), e->rc_last_qp[0] + 2);
  qp = std::min(std::max(qp, 10), 51);
  e->rc_last_qp[idr ? 1 : 0] = qp;
  return qp;
}

// A coded picture came back: account its bits and move (alpha, beta) of its kind towards what it showed.
void rate_control_update(kvz_encoder *e, size_t au_bytes, bool idr)
{
  if (e->bits_per_frame <= 0 || e->rc_pending.empty()) return;
  const int qp = e->rc_pending.front().first;
  e->rc_pending.pop_front();
  const double bits = 8.0 * au_bytes, px = (double)e->cfg.width * e->cfg.height;
  e->rc_spent += bits;
  e->rc_returned++;
  kvz_encoder::RcModel &m = e->rc_model[idr ? 1 : 0];
  const double bpp = std::max(bits / px, 1e-5);
  const double ln_real = log(rc_lambda_of_qp(qp)), ln_comp = log(m.alpha) + m.beta * log(bpp);
  // Only alpha follows the content, and fast (half of the log error per picture): with K0103's step
  // sizes (0.1 / 0.05 for alpha / beta) content far from the model's natural-video starting point drives
  // beta into its clip within ten pictures and the model goes flat -- the loop then has no authority
  // over the rate.  The exponent stays at its starting value; the sliding window absorbs what a fixed
  // exponent misses.
  const double err = std::min(std::max(ln_real - ln_comp, -2.0), 2.0);
  m.alpha *= exp(0.5 * err);
  m.alpha = std::min(std::max(m.alpha, 0.001), 200.0);
}

int encoder_encode(kvz_encoder *e, kvz_picture *pic_in, kvz_data_chunk **data_out, uint32_t *len_out,
                   kvz_picture **pic_recon, kvz_picture **pic_src, kvz_frame_info *info_out)
{
  if (data_out) *data_out = NULL;
  if (len_out) *len_out = 0;
  if (pic_recon) *pic_recon = NULL;
  if (pic_src) *pic_src = NULL;
  if (!e) { b200::set_error("encoder_encode: NULL encoder"); return 0; }
  if (e->tiled) {
    int n;
    if (pic_in) {
      const size_t ysz = (size_t)e->cfg.width * e->cfg.height;
      if (pic_in->width != e->cfg.width || pic_in->height != e->cfg.height || !pic_in->y || pic_in->u != pic_in->y + ysz ||
          pic_in->v != pic_in->u + ysz / 4 || pic_in->stride != pic_in->width) {
        b200::set_error("encoder_encode: tiled encoding needs a contiguous picture from picture_alloc of the configured size");
        return 0;
      }
      n = b200_tiled_encode(e->tiled, pic_in->y, e->tiled_out.data(), (int)e->tiled_out.size());
    } else {
      n = b200_tiled_flush(e->tiled, e->tiled_out.data(), (int)e->tiled_out.size());
    }
    if (n < 0) return 0;
    if (n == 0) return 1;
    e->au.assign(e->tiled_out.begin(), e->tiled_out.begin() + n);
    if (data_out) {
      *data_out = to_chunks(e->au);
      if (!*data_out) { b200::set_error("encoder_encode: out of memory"); return 0; }
    }
    if (len_out) *len_out = (uint32_t)n;
    if (info_out) memset(info_out, 0, sizeof(*info_out));
    return 1;
  }
  bool ok;
  if (pic_in) {
    if (pic_in->width != e->cfg.width || pic_in->height != e->cfg.height || !pic_in->y || !pic_in->u || !pic_in->v) {
      b200::set_error("encoder_encode: picture does not match the configured %dx%d", e->cfg.width, e->cfg.height);
      return 0;
    }
    if (e->eng.cfg.qp_delta) {
      // ROI: Kvazaar reads the delta-QP map at the top-left corner of each LCU (the reference's
      // overlay code compensates for exactly that, videodrawhelper.cpp:737-739)
      const int cols = e->eng.fp.ctb_cols, rows = e->eng.fp.ctb_rows;
      if (pic_in->roi.roi_array && pic_in->roi.width > 0 && pic_in->roi.height > 0 && e->cfg.target_bitrate == 0) {
        std::vector<int8_t> dqp((size_t)cols * rows);
        for (int cy = 0; cy < rows; cy++)
          for (int cx = 0; cx < cols; cx++)
            dqp[(size_t)cy * cols + cx] = pic_in->roi.roi_array[(size_t)(cy * pic_in->roi.height / rows) * pic_in->roi.width + cx * pic_in->roi.width / cols];
        e->eng.set_ctu_dqp(dqp.data(), cols * rows);
      } else {
        e->eng.set_ctu_dqp(nullptr, 0);
      }
    }
    if (e->bits_per_frame > 0) {
      const bool idr = e->eng.frame_idx == 0 || (e->cfg.intra_period > 0 && e->eng.frame_idx % e->cfg.intra_period == 0);
      const int qp = rate_control_pick_qp(e, idr);
      e->eng.set_qp(qp);
      e->rc_pending.emplace_back(qp, idr);
      e->rc_submitted++;
    }
    const size_t ysz = (size_t)e->cfg.width * e->cfg.height;
    if (pic_in->u == pic_in->y + ysz && pic_in->v == pic_in->u + ysz / 4 && pic_in->stride == pic_in->width) {
      ok = e->eng.encode_host(pic_in->y, e->au, pic_in->base_image == pic_in);
    } else {
      std::vector<uint8_t> tmp(ysz + ysz / 2);
      for (int r = 0; r < pic_in->height; r++) memcpy(&tmp[(size_t)r * pic_in->width], pic_in->y + (size_t)r * pic_in->stride, pic_in->width);
      for (int r = 0; r < pic_in->height / 2; r++) {
        memcpy(&tmp[ysz + (size_t)r * (pic_in->width / 2)], pic_in->u + (size_t)r * (pic_in->stride / 2), pic_in->width / 2);
        memcpy(&tmp[ysz + ysz / 4 + (size_t)r * (pic_in->width / 2)], pic_in->v + (size_t)r * (pic_in->stride / 2), pic_in->width / 2);
      }
      ok = e->eng.encode_host(tmp.data(), e->au);
    }
    e->pts.push_back(pic_in->pts);
  } else {
    ok = e->eng.flush(e->au);
  }
  if (!ok) return 0;
  if (e->au.empty()) return 1;                      // pipeline still filling / nothing left to drain
  rate_control_update(e, e->au.size(), e->eng.last_idr != 0);
  if (data_out) {
    *data_out = to_chunks(e->au);
    if (!*data_out) { b200::set_error("encoder_encode: out of memory"); return 0; }
  }
  if (len_out) *len_out = (uint32_t)e->au.size();
  if (info_out) {
    memset(info_out, 0, sizeof(*info_out));
    info_out->poc = e->eng.last_poc;
    info_out->qp = (int8_t)e->eng.last_qp;
    info_out->nal_unit_type = e->eng.last_idr ? KVZ_NAL_IDR_W_RADL : KVZ_NAL_TRAIL_R;
    info_out->slice_type = e->eng.last_idr ? KVZ_SLICE_I : KVZ_SLICE_P;
    if (!e->eng.last_idr) { info_out->ref_list_len[0] = 1; info_out->ref_list[0][0] = e->eng.last_poc - 1; }
  }
  if (!e->pts.empty()) e->pts.pop_front();
  if (pic_recon && e->cfg.return_recon && e->eng.cfg.depth == 1) {
    kvz_picture *r = picture_alloc(e->cfg.width, e->cfg.height);
    if (r) {
      cudaStreamSynchronize(e->eng.stream);
      const uint8_t *rec = e->eng.last_rec();              // coded size: crop to the conformance window
      const int cw = e->eng.fp.w, ch = e->eng.fp.h, w = e->cfg.width, h = e->cfg.height;
      cudaMemcpy2D(r->y, w, rec, cw, w, h, cudaMemcpyDeviceToHost);
      cudaMemcpy2D(r->u, w / 2, rec + (size_t)cw * ch, cw / 2, w / 2, h / 2, cudaMemcpyDeviceToHost);
      cudaMemcpy2D(r->v, w / 2, rec + (size_t)cw * ch * 5 / 4, cw / 2, w / 2, h / 2, cudaMemcpyDeviceToHost);
      *pic_recon = r;
    }
  }
  return 1;
}

const kvz_api kApi = {config_alloc, config_destroy, config_init, config_parse, picture_alloc, picture_free,
                      chunk_free, encoder_open, encoder_close, encoder_headers, encoder_encode, picture_alloc_csp};

}  // namespace

// Changes the target bitrate of a running encoder (what an RTCP-driven allocation such as the
// reference's ResourceAllocator::addRTCPReport, resourceallocator.cpp:67-104, would feed back); 0
// switches rate control off.  The (alpha, beta) models are kept.
extern "C" int b200_kvz_set_bitrate(kvz_encoder *e, int bits_per_second)
{
  if (!e || bits_per_second < 0 || e->tiled) { b200::set_error("b200_kvz_set_bitrate: bad arguments"); return B200_ERR_ARG; }
  e->cfg.target_bitrate = bits_per_second;
  e->bits_per_frame = bits_per_second > 0 && e->cfg.framerate_num > 0 ? (double)bits_per_second * e->cfg.framerate_denom / e->cfg.framerate_num : 0;
  // restart the window so that the old rate's surplus / deficit does not leak into the new target
  e->rc_spent = (double)e->rc_returned * e->bits_per_frame;
  return B200_OK;
}

// The reference's reaction to an RTCP receiver report (resourceallocator.cpp:67-90): halve the bitrate
// when more packets were lost, take 10 % off when only the jitter grew, add 10 % otherwise.
extern "C" int b200_rtcp_bitrate_update(int bitrate, int lost_increased, int jitter_increased)
{
  if (lost_increased) return bitrate / 2;
  if (jitter_increased) return (int)(bitrate * 0.9);
  return (int)(bitrate * 1.1);
}

// The engine parameters a preset stands for (what encoder_open derives from "preset"): lets the
// benchmark and the tests run the bare engine in exactly the configuration kvz_api would.
extern "C" int b200_enc_params_from_preset(const char *preset, b200_enc_params *p)
{
  if (!preset || !p) return B200_ERR_ARG;
  for (const Preset &pr : kPresets)
    if (!strcmp(preset, pr.name)) {
      p->search_range = pr.me_range; p->me_coarse = pr.me_coarse; p->sao = pr.sao ? 2 : 0; p->intra_in_p = 1; p->intra_satd = pr.intra_satd; p->subme_satd = pr.subme_satd;
      return B200_OK;
    }
  b200::set_error("unknown preset '%s'", preset);
  return B200_ERR_ARG;
}

extern "C" const kvz_api *kvz_api_get(int bit_depth)
{
  if (bit_depth != 8) { b200::set_error("kvz_api_get: only 8-bit video is supported (asked for %d)", bit_depth); return NULL; }
  return &kApi;
}
