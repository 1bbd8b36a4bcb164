/*
 * b200media.h -- C ABI of libb200media.so: colour conversion entry points.
 *
 * Drop-in boundary for the display- and camera-conversion stages of the
 * uvgComm media Filter chain.  Every entry point is
 * plain C: pointers, sizes and ints only.  Host-buffer entry points take the
 * same buffers the reference functions take and move them to/from the GPU
 * internally; the *_dev entry points work on device-resident, batched frames
 * and are what adjacent GPU filters (and the benchmark's HBM-resident leg)
 * use.  There is NO CPU fallback: without a CUDA device every compute entry
 * point returns B200_ERR_CUDA and sets b200_last_error().
 *
 * The encoder boundary (kvz_api) is in b200_kvazaar.h, the decoder boundary
 * (libOpenHevc*) in b200_openhevc.h.
 */
#ifndef B200MEDIA_H_
#define B200MEDIA_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_OK            0
#define B200_ERR_ARG      -1   /* bad argument / unsupported fourcc (libyuv convention: -1) */
#define B200_ERR_CUDA     -2   /* no device, launch or copy failure */

/* ---- runtime ---------------------------------------------------------- */
int         b200_device_count(void);          /* 0 when no usable GPU */
int         b200_set_device(int device);      /* per calling thread */
const char *b200_last_error(void);            /* thread-local, never NULL */
const char *b200_version(void);
/* Number of kernels this library has launched in the calling process. */
unsigned long long b200_launch_count(void);
/* Measured peak of one integer instruction on the current device, thread-level instructions per
 * second (register-only microbenchmark, csrc/int_peaks.cu): the roofline denominator of the
 * integer-pipe-bound encoder kernels.  kind: 0 vabsdiff4.add, 1 dp4a, 2 dp2a,
 * 3 mad.lo.s32, 4 shf, 5 add.u32.  < 0 on error. */
double      b200_int_peak(int kind);

/* ---- FOURCC codes (values identical to libyuv's video_common.h) -------- */
#define B200_FOURCC(a, b, c, d) \
  ((uint32_t)(a) | ((uint32_t)(b) << 8) | ((uint32_t)(c) << 16) | ((uint32_t)(d) << 24))
#define B200_FOURCC_I420 B200_FOURCC('I', '4', '2', '0')
#define B200_FOURCC_I422 B200_FOURCC('I', '4', '2', '2')
#define B200_FOURCC_NV12 B200_FOURCC('N', 'V', '1', '2')
#define B200_FOURCC_NV21 B200_FOURCC('N', 'V', '2', '1')
#define B200_FOURCC_YUY2 B200_FOURCC('Y', 'U', 'Y', '2')
#define B200_FOURCC_YUYV B200_FOURCC('Y', 'U', 'Y', 'V')
#define B200_FOURCC_UYVY B200_FOURCC('U', 'Y', 'V', 'Y')
#define B200_FOURCC_ARGB B200_FOURCC('A', 'R', 'G', 'B')
#define B200_FOURCC_BGRA B200_FOURCC('B', 'G', 'R', 'A')
#define B200_FOURCC_ABGR B200_FOURCC('A', 'B', 'G', 'R')
#define B200_FOURCC_RGBA B200_FOURCC('R', 'G', 'B', 'A')
#define B200_FOURCC_24BG B200_FOURCC('2', '4', 'B', 'G')
#define B200_FOURCC_RAW  B200_FOURCC('r', 'a', 'w', ' ')
#define B200_FOURCC_MJPG B200_FOURCC('M', 'J', 'P', 'G')

/* ---- display conversion: I420 -> RGB32 (memory order B,G,R,0) ----------
 * Replaces yuv420_to_rgb_i_avx2_mt / _avx2 / _sse41
 * (reference src/media/processing/yuvconversions.h:9-11, called from
 * src/media/processing/yuvtorgb32.cpp:40-52).  Same buffers: `input` is packed
 * I420 (w*h*3/2 bytes), `output` is caller-allocated 4*w*h bytes.  Returns 1 on
 * success like the reference; <0 on error.  Result is bit-identical to the
 * reference's SIMD variants for every even w,h (the reference's scalar `_c`
 * fallback, taken when w%16!=0, has U/V and R/B swapped -- not reproduced). */
int b200_yuv420_to_rgb32(const uint8_t *input, uint8_t *output, uint16_t width, uint16_t height);

/* Replaces half_rgb (yuvconversions.h:24; caller halfrgbfilter.cpp:27-38). */
int b200_half_rgb(const uint8_t *input, uint8_t *output, uint16_t width, uint16_t height);

/* Replaces flip_rgb (yuvconversions.h:26-27; caller filter.cpp:263-294).
 * With both flags 0 the output buffer is left untouched, like the reference. */
int b200_flip_rgb(const uint8_t *input, uint8_t *output, uint16_t width, uint16_t height,
                  int horizontally, int vertically);

/* Self-view chain fused (Camera -> I420 -> RGB32 -> HalfRGB -> mirrored Display,
 * filtergraph.cpp:247-325): equals b200_flip_rgb(b200_half_rgb(b200_yuv420_to_rgb32(in))) with the
 * stages selected by the flags, written in one pass.  Output: (half ? w/2 x h/2 : w x h) RGB32. */
int b200_selfview(const uint8_t *input, uint8_t *output, uint16_t width, uint16_t height,
                  int half, int horizontally, int vertically);

/* ---- camera conversion: any supported format -> I420 -------------------
 * C twin of libyuv::ConvertToI420 (sole call site
 * src/media/processing/libyuvconverter.cpp:120-127).  Same argument list;
 * crop must be the whole frame and rotation 0 (the only form the reference
 * uses), otherwise B200_ERR_ARG.  Returns 0 on success, -1 for an
 * unsupported fourcc (destination untouched), like libyuv. */
int b200_ConvertToI420(const uint8_t *sample, size_t sample_size,
                       uint8_t *dst_y, int dst_stride_y,
                       uint8_t *dst_u, int dst_stride_u,
                       uint8_t *dst_v, int dst_stride_v,
                       int crop_x, int crop_y, int src_width, int src_height,
                       int crop_width, int crop_height, int rotation, uint32_t fourcc);

/* Bytes of one packed source frame of `fourcc` at w x h (0 if unsupported). */
size_t b200_frame_bytes(uint32_t fourcc, int width, int height);

/* ---- device-resident, batched entry points ------------------------------
 * All pointers are device pointers; frames are packed back to back
 * (frame f of the input at d_in + f*frame_bytes).  `stream` is a cudaStream_t
 * passed as void* (NULL = default stream).  Asynchronous: returns after
 * enqueueing. */
int b200_i420_to_rgb32_dev(const uint8_t *d_i420, uint8_t *d_bgra, int width, int height,
                           int n_frames, void *stream);
int b200_half_rgb_dev(const uint8_t *d_in, uint8_t *d_out, int width, int height,
                      int n_frames, void *stream);
int b200_flip_rgb_dev(const uint8_t *d_in, uint8_t *d_out, int width, int height,
                      int horizontally, int vertically, int n_frames, void *stream);
int b200_selfview_dev(const uint8_t *d_i420, uint8_t *d_out, int width, int height, int half,
                      int horizontally, int vertically, int n_frames, void *stream);
int b200_convert_to_i420_dev(const uint8_t *d_src, uint8_t *d_i420, int width, int height,
                             uint32_t fourcc, int n_frames, void *stream);
/* MJPG (the thirteenth format of LibYUVConverter, libyuvconverter.cpp:94): one baseline JPEG frame in HOST
 * memory -> packed I420 in DEVICE memory (for b200_enc_encode_dev).  The Huffman-coded scan is read on the
 * host (one serial bit stream per frame); dequantisation, libjpeg's accurate integer IDCT and the conversion
 * of 4:2:2 / 4:4:4 / 4:0:0 to 4:2:0 (libyuv's rules) run on the GPU.  8-bit baseline frames with one
 * interleaved scan, restart intervals, frames without DHT (the tables of T.81 Annex K).  The frame must be
 * width x height (libyuv::MJPGToI420 fails otherwise).  One stream per calling thread.  B200_OK or < 0. */
int b200_mjpg_to_i420_dev(const uint8_t *jpeg, size_t jpeg_bytes, uint8_t *d_i420, int width, int height, void *stream);
/* Host only (no GPU needed): B200_OK when the frame is one b200_mjpg_to_i420_dev converts -- headers parsed, the
 * whole scan read -- with its size and subsampling (420, 422, 444 or 400); < 0 with b200_last_error() otherwise. */
int b200_mjpg_probe(const uint8_t *jpeg, size_t jpeg_bytes, int *width, int *height, int *subsampling);

#ifdef __cplusplus
}
#endif
#endif /* B200MEDIA_H_ */
