/*
 * b200_openhevc.h -- libOpenHevc-shaped C ABI of the B200 HEVC decoder.
 *
 * SOURCE-compatible stand-in for gpac's openHevcWrapper.h (the header the reference copies next to
 * its OpenHEVC build, /root/reference/dependencies/openhevc.cmake:24-25) for everything
 * src/media/processing/openhevcfilter.cpp touches:
 *   libOpenHevcInit(threads, thread_type)                  :38-46
 *   libOpenHevcStartDecoder                                :49
 *   libOpenHevcSetTemporalLayer_id / SetActiveDecoders / SetViewLayers   :54-56
 *   libOpenHevcVersion                                     :64
 *   libOpenHevcDecode(h, buf, len, pts)                    :145   (<0 error, 0 no picture, >0 picture ready)
 *   libOpenHevcGetOutput(h, got, &frame)                   :195
 *   libOpenHevcGetPictureInfo(h, &frame.frameInfo)         :199
 *   libOpenHevcFlush / libOpenHevcClose                    :81-82
 * Struct members read by the reference: OpenHevc_Frame::{pvY,pvU,pvV,frameInfo},
 * frameInfo.{nWidth,nHeight,nYPitch,nUPitch,frameRate.num,frameRate.den}   :201-233
 *
 * Decoding scope: Main profile 8-bit 4:2:0 low-delay streams as a Kvazaar-family
 * peer sends them -- 64x64 CTUs with CUs of 8..64; inter CUs 2Nx2N (merge / skip / AMVP, several
 * reference pictures from any short-term RPS of earlier pictures, temporal motion vector candidates);
 * intra CUs 2Nx2N and NxN of every size with explicit chroma modes and strong intra smoothing;
 * transform trees down to 4x4 luma blocks (DST-VII for intra); sign data hiding; cu_qp_delta with one
 * quantisation group per CTU; chroma QP and deblocking offsets; scaling lists (the default ones or lists
 * carried in the SPS / PPS); deblocking and SAO; WPP entry points;
 * cabac_init_flag; uniformly spaced tile grids without loop filtering across tiles whose motion stays
 * inside the tile (what b200_tiled_* and kvz_api "tiles" emit), with or without WPP inside the tiles.
 * A conformance window is honoured: nWidth / nHeight are the window, pvY / pvU / pvV point at its origin
 * inside the coded picture and nYPitch / nUPitch / nVPitch are the coded pitches (always even).
 * Not decoded: B slices, AMP / 2NxN / Nx2N inter partitions, PCM, transform skip,
 * transquant bypass, long-term references, several slices per picture, quantisation groups below the
 * CTU.  Those make libOpenHevcDecode return -1 with the reason in
 * b200_last_error() -- never a silently wrong picture.
 */
#ifndef B200_OPENHEVC_H_
#define B200_OPENHEVC_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void *OpenHevc_Handle;

typedef struct OpenHevc_Rational {
  int num;
  int den;
} OpenHevc_Rational;

typedef struct OpenHevc_FrameInfo {
  int nYPitch;
  int nUPitch;
  int nVPitch;
  int nBitDepth;
  int nWidth;
  int nHeight;
  int chromat_format;
  OpenHevc_Rational sample_aspect_ratio;
  OpenHevc_Rational frameRate;
  int display_picture_number;
  int flag;
  int64_t nTimeStamp;
} OpenHevc_FrameInfo;

typedef struct OpenHevc_Frame {
  void *pvY;                 /* decoder-owned, valid until the next libOpenHevcDecode call */
  void *pvU;
  void *pvV;
  OpenHevc_FrameInfo frameInfo;
} OpenHevc_Frame;

/* thread_type: 1 frame, 2 slice, 3 frame+slice (openhevcfilter.cpp:11).  Slice threading needs no
 * host threads here (WPP substreams are parsed in parallel on the GPU) and outputs every picture
 * in the call that completes it.  Frame threading keeps nb_pthreads pictures in flight (their
 * CABAC parses run concurrently on the GPU) and, like OpenHEVC's frame threads, delays output by
 * nb_pthreads - 1 pictures. */
OpenHevc_Handle libOpenHevcInit(int nb_pthreads, int thread_type);
int  libOpenHevcStartDecoder(OpenHevc_Handle h);            /* -1 on failure (no CUDA device) */
/* buff: one or more NAL units, each with a 3- or 4-byte start code (the reference passes one NAL
 * per call, openhevcfilter.cpp:145).  Returns 1 when a picture became available, 0 when not, -1 on
 * error.  buff == NULL / nal_len == 0 hands over the next held-back picture (frame threading),
 * 0 when none is left. */
int  libOpenHevcDecode(OpenHevc_Handle h, const unsigned char *buff, int nal_len, int64_t pts);
int  libOpenHevcGetOutput(OpenHevc_Handle h, int got_picture, OpenHevc_Frame *frame);   /* >0: frame filled */
void libOpenHevcGetPictureInfo(OpenHevc_Handle h, OpenHevc_FrameInfo *info);
void libOpenHevcSetTemporalLayer_id(OpenHevc_Handle h, int id);
void libOpenHevcSetActiveDecoders(OpenHevc_Handle h, int n);
void libOpenHevcSetViewLayers(OpenHevc_Handle h, int n);
void libOpenHevcSetDebugMode(OpenHevc_Handle h, int level);
void libOpenHevcSetCheckMD5(OpenHevc_Handle h, int on);
const char *libOpenHevcVersion(OpenHevc_Handle h);
void libOpenHevcFlush(OpenHevc_Handle h);
void libOpenHevcClose(OpenHevc_Handle h);

/* B200 extension: copy of the last decoded picture as packed I420 (w*h*3/2 bytes). */
int  b200_dec_last_picture(OpenHevc_Handle h, uint8_t *dst, int cap);
/* B200 extensions for an adjacent GPU filter (decode -> I420-to-RGB32 without a
 * host bounce).  b200_dec_output_dev: DEVICE pointer to the last output picture, packed I420,
 * complete when libOpenHevcDecode returned; read-only (it is the next picture's reference) and valid
 * until the next libOpenHevcDecode call.  b200_dec_set_host_output(h, 0) skips the device-to-host
 * copy; libOpenHevcGetOutput planes are then stale. */
const uint8_t *b200_dec_output_dev(OpenHevc_Handle h);
void b200_dec_set_host_output(OpenHevc_Handle h, int on);
/* Number of P pictures decoded although the picture they reference (POC - 1) was not the previously
 * decoded one -- a picture was lost on the way.  Like OpenHEVC the decoder conceals with the last
 * picture it has and the error drifts until the next IDR; an application that watches this counter
 * can ask the sender for a key frame instead. */
int  b200_dec_missing_refs(OpenHevc_Handle h);

/* Host-only look at the parameter sets of a stream (no GPU needed, e.g. when the sprop parameter sets of a
 * peer arrive): parses the VPS / SPS / PPS NAL units of an Annex-B buffer and says whether this decoder
 * takes streams coded with them.  struct_size lets the structure grow. */
typedef struct b200_stream_info {
  int struct_size;
  int width, height;               /* what libOpenHevcGetPictureInfo will report (the conformance window) */
  int coded_width, coded_height;   /* multiples of 8 */
  int crop_left, crop_top;         /* window origin inside the coded picture, luma samples */
  int fps_num, fps_den;            /* VUI timing info, 0 / 0 when absent */
  int tile_cols, tile_rows, wpp;
  int sao, sign_hiding, qp_delta, tmvp, strong_intra, cabac_init_present;
  int scaling_list;                /* 0 off, 1 default lists, 2 lists carried in the SPS, 3 in the PPS */
  int max_tr_depth_inter, max_tr_depth_intra;
  int max_dec_pic_buffering;
  int decodable;                   /* 1 = within the decoding scope stated at the top of this header (parameter sets and
                                      every slice header found in the buffer) */
  char reason[96];                 /* why not, when decodable == 0 (the first reason met) */
  /* slice segment headers found after the parameter sets (none: slices = 0 and the fields below are 0) */
  int slices;                      /* slice NAL units whose header was parsed */
  int slice_type, slice_qp;        /* of the last one: 0 B, 1 P, 2 I; SliceQpY */
  int entry_points;                /* ... its number of entry points (substreams - 1) */
  int num_ref_idx_l0, rps_pictures;/* ... active reference indices, pictures in its reference picture set */
} b200_stream_info;
/* Returns 0 when an SPS and the PPS that refers to it were found and parsed (the last pair in the buffer counts),
 * B200_ERR_ARG otherwise (b200_last_error() says why).  scaling_table: NULL, or 1552 bytes that receive the scaling
 * factors in force (layout: DESIGN.md, the ScalingTable of hevc_headers.h) when scaling_list != 0. */
int  b200_dec_probe(const uint8_t *annexb, size_t n, b200_stream_info *info, uint8_t *scaling_table);

#ifdef __cplusplus
}
#endif
#endif /* B200_OPENHEVC_H_ */
