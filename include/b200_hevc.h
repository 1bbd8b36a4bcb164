/*
 * b200_hevc.h -- plain-argument C ABI of the B200 HEVC encoder / decoder engines.
 *
 * These are the engines underneath the reference-shaped boundaries (kvz_api in b200_kvazaar.h,
 * libOpenHevc* in b200_openhevc.h).  They exist so that the benchmark can keep frames resident
 * in HBM and so that the parity tests can read intermediate state (cu map, levels,
 * reconstruction) and compare it with the CPU oracle stage by stage.
 */
#ifndef B200_HEVC_H_
#define B200_HEVC_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Encoder: constant QP, low-delay P (each picture references the previous reconstruction),
 * IDR every intra_period pictures (0 = first only), one slice, WPP substreams.
 * width/height multiples of 8 (the reference's camera filter truncates to that, camerafilter.cpp:329-331); other even
 * sizes through b200_enc_params::src_width / src_height (padding + conformance window).
 * depth = pictures in flight (Kvazaar's owf + 1): with depth > 1 the access unit of picture n is
 * returned by the call that submits picture n + depth - 1 (0 = nothing ready yet) and the tail is
 * drained with b200_enc_flush(), exactly like kvz_api's encoder_encode(pic = NULL).
 * Returns NULL on error (see b200_last_error()). */
void *b200_enc_open(int width, int height, int qp, int intra_period, int search_range, int deblock, int debug,
                    int depth);
/* Same, with cu_qp_delta enabled in the PPS (one quantisation group per CTU): the region-of-interest
 * path behind kvz_picture::roi (kvazaarfilter.cpp:423-431).  Without offsets every CTU codes delta 0. */
void *b200_enc_open_roi(int width, int height, int qp, int intra_period, int search_range, int deblock, int debug,
                        int depth);
/* The same engine with every option by name.  Fill with b200_enc_params_default(), change what is
 * needed, open.  struct_size lets the structure grow: fields beyond it keep their defaults. */
typedef struct b200_enc_params {
  int struct_size;               /* sizeof(b200_enc_params) of the caller */
  int width, height, qp, intra_period, search_range, deblock, debug, depth;
  int qp_delta;                  /* cu_qp_delta_enabled_flag (what b200_enc_open_roi sets) */
  int fps_num, fps_den;          /* both > 0: VUI timing info in the SPS */
  int sao;                       /* sample adaptive offset (edge / band offsets per CTU) after deblocking;
                                    2 = also use sao_merge_left / _up where parameters repeat */
  int intra_in_p;                /* 16x16 intra CUs in P pictures where inter prediction is poor */
  int me_coarse;                 /* two-level motion search: range of the coarse level in 4x4-mean samples (multiple
                                    of 4, <= 32; 16 = +-64 luma samples); search_range (<= 16) is then the window
                                    searched around the zero vector and around each 32x32 block's coarse vector */
  int intra_satd;                /* I pictures: the 35-mode intra search compares the Hadamard SATD of the residual
                                    (8x8 tiles) instead of its SAD */
  int subme_satd;                /* P pictures: the half- / quarter-sample motion refinement compares SATD instead of SAD */
  int vaq;                       /* variance adaptive quantisation strength (Kvazaar --vaq, kvazaarfilter.cpp:280-284), 0 = off,
                                    1..20: every picture each CTU's QP moves by strength * 0.1 * ln(CTU variance / picture
                                    variance), on top of the offsets of b200_enc_set_ctu_dqp; needs qp_delta */
  int scaling_list;              /* 1 = scaling_list_enabled_flag with the default lists of the standard (Kvazaar
                                    --scaling-list default, kvazaarfilter.cpp:236-243): quantisation steps grow with frequency */
  int src_width, src_height;     /* size of the pictures passed to b200_enc_encode[_dev] when it is not a multiple of 8
                                    (even, at most 6 samples short of width / height; 0 = width / height): the margins
                                    are filled by edge repetition on the GPU and the SPS carries a conformance window, so
                                    decoders output src_width x src_height.  b200_enc_debug_read returns coded-size planes */
  int mv_edges;                  /* bit 0 / 1 / 2 / 3: motion vectors must not reach beyond the left / right / top / bottom
                                    picture edge (no sample outside, interpolation taps included) -- 15 = Kvazaar's
                                    mv-constraint "frame" (kvazaarfilter.cpp:246-276) */
  int vps_period;                /* Kvazaar's vps-period (kvazaarfilter.cpp:221): VPS / SPS / PPS before every n-th IDR picture,
                                    the first always; 0 = before the first picture only; default 1 */
} b200_enc_params;
void  b200_enc_params_default(b200_enc_params *p);
/* Host only (no GPU needed): the VPS, SPS and PPS (Annex B, 4-byte start codes) an encoder opened with these
 * parameters writes -- for out-of-band signalling (sprop-vps / -sps / -pps) before an encoder exists, and what
 * kvz_api's encoder_headers returns.  tile_cols / tile_rows / wpp as in b200_tiled_params (1, 1, 1 for
 * b200_enc_open*).  Returns the number of bytes, or -(bytes needed) when cap is too small. */
int   b200_enc_parameter_sets(const b200_enc_params *p, int tile_cols, int tile_rows, int wpp, uint8_t *out, int cap);
void *b200_enc_open_params(const b200_enc_params *p);
/* Fills search_range, me_coarse, sao, intra_in_p, intra_satd and subme_satd with what the kvz_api preset of that name selects
 * ("ultrafast" ... "placebo"); the other fields are left alone.  0 on success. */
int   b200_enc_params_from_preset(const char *preset, b200_enc_params *p);
/* Per-CTU QP offsets (raster, one int8 per 64x64 CTU, n = CTU count) for the pictures submitted
 * from now on; CTU QP = clip(qp + dqp, 0, 51).  NULL clears.  Needs b200_enc_open_roi. */
int   b200_enc_set_ctu_dqp(void *enc, const int8_t *dqp, int n);
void  b200_enc_close(void *enc);
/* Encode one packed I420 picture from host / device memory; writes one Annex-B access unit
 * (VPS+SPS+PPS precede every IDR).  Returns its size, or <0 (-needed when cap is short). */
int   b200_enc_encode(void *enc, const uint8_t *i420, uint8_t *out, int cap);
int   b200_enc_encode_dev(void *enc, const uint8_t *d_i420, uint8_t *out, int cap);
int   b200_enc_flush(void *enc, uint8_t *out, int cap);   /* next pending access unit, 0 when drained */
int   b200_enc_pending(void *enc);
/* Per-kernel device time measured with CUDA events on the launching stream.  Kernel ids:
 * 0 intra, 1 motion search, 2 inter reconstruction, 3 merge/skip modes, 4 deblocking, 5 binarisation,
 * 6 arithmetic coding (both phases), 7 pack, 8 SAO. */
int   b200_enc_set_profile(void *enc, int on);
int   b200_enc_get_profile(void *enc, double *ms, unsigned long long *count, int n);
/* Work counters of the motion search since b200_enc_set_profile(enc, 1): out[0] CTUs, [1] 32x32 quadrants
 * whose second centre set was searched, [2] 16x16 intra mode searches, [3] intra CUs chosen.  n >= 4. */
int   b200_enc_get_me_stats(void *enc, unsigned long long *out, int n);
/* begin/end (ms since open) of each kernel id of the last returned picture: out[2*id], out[2*id+1] */
int   b200_enc_get_timeline(void *enc, float *out, int n);
int   b200_enc_last_was_idr(void *enc);
unsigned long long b200_enc_last_bins(void *enc);

/* Test hooks.  what: 0 reconstruction after deblocking, 1 before deblocking (debug=1),
 * 2 cu map (16 bytes per 8x8 unit), 3 quantised levels (int16, I420-shaped). */
int   b200_enc_debug_read(void *enc, int what, void *dst, size_t bytes);
int   b200_enc_debug_set_reference(void *enc, const uint8_t *i420);

/* ---- tile columns (BASELINE config 3: 4K call with tiles) -----------------------------------
 * One encoder per tile column (uniform spacing), each coding its strip as a picture of its own:
 * motion never crosses an interior tile edge (Kvazaar's mv-constraint frametilemargin) and tiles are
 * not loop-filtered across, so the strips need nothing from each other and may run on different
 * GPUs (`devices`: CUDA ordinals, tile i on devices[i % n_devices]; n_devices = 0 = current device).
 * wpp = 0: one substream per tile (HEVC Main: tiles or WPP, not both -- the mode verified with an
 * independent decoder); wpp = 1: one substream per CTU row of every tile, as Kvazaar emits with
 * --tiles and --wpp.  Every tile must be at least two CTUs (128 samples) wide.  Same call shape as
 * b200_enc_*: access units come back in order, `depth` pictures in flight. */
void *b200_tiled_open(int width, int height, int qp, int intra_period, int search_range, int deblock, int depth,
                      int tile_cols, int wpp, const int *devices, int n_devices);
/* The same with every option by name (see b200_enc_params). */
typedef struct b200_tiled_params {
  int struct_size;
  int width, height, qp, intra_period, search_range, deblock, depth, tile_cols, wpp;
  int fps_num, fps_den;          /* both > 0: VUI timing info in the SPS */
  int sao;                       /* SAO inside every tile (never across tile edges) */
  int intra_in_p;                /* intra CUs in P pictures */
  int me_coarse;                 /* two-level motion search (see b200_enc_params) */
  int intra_satd;                /* SATD-based intra mode search in I pictures (see b200_enc_params) */
  int subme_satd;                /* SATD-based fractional motion refinement (see b200_enc_params) */
  int tile_rows;                 /* uniform tile rows (default 1): tile_cols x tile_rows tiles in raster order; a tile
                                    row may be a single CTU row high */
  int scaling_list;              /* default scaling lists (see b200_enc_params) */
  int mv_edges;                  /* picture edges motion must not cross (see b200_enc_params); tile edges never are */
  int vps_period;                /* parameter sets before every n-th IDR picture (see b200_enc_params) */
} b200_tiled_params;
void  b200_tiled_params_default(b200_tiled_params *p);
void *b200_tiled_open_params(const b200_tiled_params *p, const int *devices, int n_devices);
void  b200_tiled_close(void *enc);
void  b200_tiled_set_fps(void *enc, int fps_num, int fps_den);   /* VUI timing info of the parameter sets (both > 0) */
int   b200_tiled_encode(void *enc, const uint8_t *i420, uint8_t *out, int cap);   /* host picture */
int   b200_tiled_flush(void *enc, uint8_t *out, int cap);
int   b200_tiled_pending(void *enc);
int   b200_tiled_recon(void *enc, uint8_t *dst, size_t cap);     /* last picture, depth 1 only */

/* ---- K2: SATD primitive -------------------------------------------------------
 * SATD of every 8x8 block between two 8-bit planes of width x height (multiples of 8), HM / Kvazaar
 * convention: 8x8 Hadamard of the difference, (sum |h| + 2) >> 2; the SATD of a larger block is the
 * sum of its 8x8 entries.  out: (width/8)*(height/8) values, raster.  _dev: device pointers,
 * asynchronous on `stream` (cudaStream_t as void*). */
int   b200_satd8x8(const uint8_t *a, const uint8_t *b, int width, int height, uint32_t *out);
int   b200_satd8x8_dev(const uint8_t *d_a, const uint8_t *d_b, int width, int height, uint32_t *d_out, void *stream);

#ifdef __cplusplus
}
#endif
#endif
