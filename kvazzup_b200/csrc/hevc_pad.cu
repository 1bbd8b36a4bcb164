// Source pictures whose size is not a multiple of 8 (a 1366x768 screen, say): Kvazaar codes them padded to
// the next multiple of the smallest CU and signals a conformance window (7.4.3.2.1) that the decoder crops.
// The planes of the source arrive in the top-left corner of the coded picture (2-D copies); this kernel
// fills the margins by repeating the last column / row of every plane, at most 7 luma columns and 7 rows:
// one thread per margin sample, a few kB per picture.
#include "hevc_kernels.h"

namespace b200 {
namespace {

__global__ void k_pad_edges(uint8_t *pic, int w, int h, int sw, int sh)
{
  const int right = (w - sw) * sh, bottom = w * (h - sh);            // luma margin samples
  const int per_c = ((w - sw) >> 1) * (sh >> 1) + (w >> 1) * ((h - sh) >> 1);
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  uint8_t *pl = pic;
  int pw = w, psw = sw, psh = sh, nr = right;
  if (i >= right + bottom) {                                          // a chroma plane
    i -= right + bottom;
    const int c = i >= per_c ? 1 : 0;
    if (c) i -= per_c;
    if (i >= per_c) return;
    pl = pic + (size_t)w * h + (size_t)c * (w >> 1) * (h >> 1);
    pw = w >> 1; psw = sw >> 1; psh = sh >> 1; nr = (pw - psw) * psh;
  }
  int x, y;
  if (i < nr) { const int mw = pw - psw; y = i / mw; x = psw + i % mw; }
  else { i -= nr; y = psh + i / pw; x = i % pw; }
  pl[(size_t)y * pw + x] = pl[(size_t)min(y, psh - 1) * pw + min(x, psw - 1)];
}

}  // namespace

cudaError_t launch_pad_edges(uint8_t *pic, int w, int h, int src_w, int src_h, cudaStream_t s)
{
  const int total = (w - src_w) * src_h + w * (h - src_h) +
                    2 * (((w - src_w) >> 1) * (src_h >> 1) + (w >> 1) * ((h - src_h) >> 1));
  if (total <= 0) return cudaSuccess;
  k_pad_edges<<<(total + 255) / 256, 256, 0, s>>>(pic, w, h, src_w, src_h);
  return cudaGetLastError();
}

}  // namespace b200
