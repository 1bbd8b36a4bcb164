/*
 * b200_kvazaar.h -- kvz_api-shaped C ABI of the B200 HEVC encoder.
 *
 * SOURCE-compatible stand-in for <kvazaar.h> 2.3.1 for everything the reference touches
 * (src/media/processing/kvazaarfilter.cpp): include this header instead of <kvazaar.h>, link
 * libb200media.so instead of libkvazaar, and KvazaarFilter compiles and runs unchanged.
 * Binary compatibility with libkvazaar.so is NOT claimed: kvazaar.h is not in the reference tree
 * (Kvazaar is fetched at build time, dependencies/kvazaar.cmake:10-14), so struct layouts below
 * are our own; only the names, types and semantics of the members the reference uses match.
 *
 * Reference call sites (file:line in /root/reference/src/media/processing/kvazaarfilter.cpp):
 *   kvz_api_get(8)                         :145
 *   api->config_alloc / config_init        :151, :160
 *   api->config_parse(cfg, name, value)    :172-283, :363   (returns 1 on success)
 *   cfg->width/height/framerate_*          :381-384
 *   cfg->wpp, owf, target_bitrate, lossless, mv_constraint, set_qp_in_cu, hash
 *                                          :207, :299, :223, :244, :259-275, :278, :289
 *   api->encoder_open / encoder_close      :291, :317
 *   api->config_destroy                    :318
 *   api->picture_alloc / picture_free      :69, :54, :476  (picture_free(NULL) is legal)
 *   pic->y/u/v, pts, roi.{width,height,roi_array}   :410-430
 *   api->encoder_encode(enc, pic|NULL, &chunks, &len, &recon, NULL, &info)   :435-449
 *   chunk->data/len/next, api->chunk_free  :469-475
 */
#ifndef B200_KVAZAAR_H_
#define B200_KVAZAAR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KVZ_BIT_DEPTH 8
typedef uint8_t kvz_pixel;

#define KVZ_DATA_CHUNK_SIZE 4096
typedef struct kvz_data_chunk {
  uint8_t data[KVZ_DATA_CHUNK_SIZE];
  uint32_t len;
  struct kvz_data_chunk *next;
} kvz_data_chunk;

enum kvz_hash { KVZ_HASH_NONE = 0, KVZ_HASH_CHECKSUM = 1, KVZ_HASH_MD5 = 2 };
enum kvz_mv_constraint {
  KVZ_MV_CONSTRAIN_NONE = 0, KVZ_MV_CONSTRAIN_FRAME = 1, KVZ_MV_CONSTRAIN_TILE = 2,
  KVZ_MV_CONSTRAIN_FRAME_AND_TILE = 3, KVZ_MV_CONSTRAIN_FRAME_AND_TILE_MARGIN = 4
};
enum kvz_chroma_format { KVZ_CSP_400 = 0, KVZ_CSP_420 = 1, KVZ_CSP_422 = 2, KVZ_CSP_444 = 3 };
enum kvz_slice_type { KVZ_SLICE_B = 0, KVZ_SLICE_P = 1, KVZ_SLICE_I = 2 };
enum kvz_nal_unit_type { KVZ_NAL_TRAIL_R = 1, KVZ_NAL_IDR_W_RADL = 19, KVZ_NAL_VPS_NUT = 32, KVZ_NAL_SPS_NUT = 33, KVZ_NAL_PPS_NUT = 34 };
enum kvz_rc_algorithm { KVZ_NO_RC = 0, KVZ_LAMBDA = 1, KVZ_OBA = 2 };

/* Encoder configuration.  Members the reference reads or writes directly keep Kvazaar's names. */
typedef struct kvz_config {
  int32_t width, height;                 /* "input-res" */
  int32_t framerate_num, framerate_denom;/* "input-fps" */
  int32_t qp;                            /* "qp" 0..51 */
  int32_t intra_period;                  /* "period"; 0 = only the first picture is intra */
  int32_t vps_period;                    /* "vps-period": VPS / SPS / PPS before every n-th IDR picture (the first always), 0 = first only */
  int32_t wpp;                           /* "wpp"; substreams are always one per CTU row */
  int32_t owf;                           /* "owf": pictures in flight - 1 */
  int32_t threads;                       /* "threads": accepted, meaningless on a GPU */
  int32_t target_bitrate;                /* bits/s; 0 = constant QP */
  int32_t rc_algorithm;                  /* "rc-algorithm": no-rc / lambda (frame-level lambda-domain control); oba is refused */
  int32_t lossless;                      /* must be 0 */
  enum kvz_mv_constraint mv_constraint;
  int32_t set_qp_in_cu;                  /* != 0 enables per-CTU QP (cu_qp_delta), see roi_enable */
  enum kvz_hash hash;
  int32_t deblock_enable;                /* "deblock" */
  int32_t sao_type;                      /* "sao": 0 off, != 0 edge + band offsets per CTU (hevc_sao.cu) */
  int32_t tiles_width_count, tiles_height_count;   /* "tiles": "CxR" uniform tile grid (hevc_tiles.cu) */
  int32_t slices;                        /* "slices" */
  int32_t vaq;                           /* "vaq" 0..20: variance adaptive quantisation strength (enables cu_qp_delta; ignored with tiles) */
  int32_t scaling_list;                  /* "scaling-list": off (0) or default (1: the default lists of the standard) */
  int32_t gop_lowdelay, gop_len;         /* "gop lp-g4d3t1": low-delay P is the only structure */
  int32_t me_range;                      /* full-sample search window around each centre, from "preset" or "b200-me-range" */
  int32_t me_coarse;                     /* range of the coarse search level (4x4-mean samples), "preset" or "b200-me-coarse" */
  int32_t subme_satd;                    /* SATD instead of SAD in the fractional motion refinement, from "preset" or "b200-subme-satd" */
  int32_t intra_satd;                    /* SATD instead of SAD in the intra mode search of I pictures, from "preset" or "b200-intra-satd" */
  int32_t return_recon;                  /* "b200-recon": 1 = encoder_encode also returns the reconstruction */
  int32_t device;                        /* "b200-device": CUDA device ordinal, -1 = current */
  int32_t roi_enable;                    /* "b200-roi" (also implied by set_qp_in_cu): cu_qp_delta in the PPS so that
                                            kvz_picture::roi maps take effect */
  char preset[16];
} kvz_config;

typedef struct kvz_picture {
  kvz_pixel *fulldata_buf;               /* allocation */
  kvz_pixel *fulldata;
  kvz_pixel *y, *u, *v;                  /* contiguous planes, stride == width (kvazaarfilter.cpp:410-418) */
  kvz_pixel *data[3];
  int32_t width, height, stride;
  struct kvz_picture *base_image;
  int32_t refcount;
  int64_t pts, dts;
  enum kvz_chroma_format chroma_format;
  /* caller-owned delta-QP map (kvazaarfilter.cpp:423-431).  Like Kvazaar, each 64x64 CTU takes the
   * entry at (ctu_x * width / ctus_wide, ctu_y * height / ctus_high).  Takes effect when the encoder
   * was opened with set_qp_in_cu or "b200-roi" and target_bitrate == 0; ignored otherwise. */
  struct { int width; int height; int8_t *roi_array; } roi;
} kvz_picture;

typedef struct kvz_frame_info {
  int32_t poc;
  int8_t qp;
  enum kvz_nal_unit_type nal_unit_type;
  enum kvz_slice_type slice_type;
  int ref_list[2][16];
  int ref_list_len[2];
} kvz_frame_info;

typedef struct kvz_encoder kvz_encoder;

typedef struct kvz_api {
  kvz_config *(*config_alloc)(void);
  int (*config_destroy)(kvz_config *cfg);
  int (*config_init)(kvz_config *cfg);
  /* 1 = option understood and applied (or knowingly ignored), 0 = unknown option / bad value */
  int (*config_parse)(kvz_config *cfg, const char *name, const char *value);
  kvz_picture *(*picture_alloc)(int32_t width, int32_t height);
  void (*picture_free)(kvz_picture *pic);            /* NULL is accepted */
  void (*chunk_free)(kvz_data_chunk *chunk);         /* frees the whole list; NULL is accepted */
  kvz_encoder *(*encoder_open)(const kvz_config *cfg);   /* NULL on error; see b200_last_error() */
  void (*encoder_close)(kvz_encoder *encoder);
  int (*encoder_headers)(kvz_encoder *encoder, kvz_data_chunk **data_out, uint32_t *len_out);
  /* pic_in may be NULL to drain.  *data_out == NULL means "no access unit ready".  Outputs come
   * back in input order.  pic_in stays caller-owned.  As with Kvazaar (which keeps a reference to the
   * input picture until it is coded), a picture from picture_alloc must stay UNTOUCHED until its own
   * access unit has been returned: it is page-locked and uploaded asynchronously in place.  With
   * owf = n that means a ring of n + 1 pictures, exactly what the reference keeps
   * (kvazaarfilter.cpp:76-88, 299).  A picture whose planes were not allocated by picture_alloc is
   * copied before the call returns and may be reused at once.
   * Returns 1 on success, 0 on failure. */
  int (*encoder_encode)(kvz_encoder *encoder, kvz_picture *pic_in, kvz_data_chunk **data_out, uint32_t *len_out,
                        kvz_picture **pic_recon, kvz_picture **pic_src, kvz_frame_info *info_out);
  kvz_picture *(*picture_alloc_csp)(enum kvz_chroma_format chroma_format, int32_t width, int32_t height);
} kvz_api;

/* B200 extensions for bitrate adaptation during a call: a new target for a running encoder (0 = back to
 * constant QP), and the reference's reaction to an RTCP receiver report (resourceallocator.cpp:67-90) as
 * a function. */
int b200_kvz_set_bitrate(kvz_encoder *encoder, int bits_per_second);
int b200_rtcp_bitrate_update(int bitrate, int lost_increased, int jitter_increased);

/* bit_depth must be 8; any other value returns NULL (kvazaarfilter.cpp:145-150 treats NULL as failure). */
const kvz_api *kvz_api_get(int bit_depth);

#ifdef __cplusplus
}
#endif
#endif /* B200_KVAZAAR_H_ */
