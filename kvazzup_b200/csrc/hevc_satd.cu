// K2 of SURVEY.md 8a: sum of absolute Hadamard-transformed differences (SATD), HM xCalcHADs8x8
// convention: 8x8 Hadamard of the residual, (sum |h| + 2) >> 2 per 8x8 block; larger blocks are sums
// of their 8x8 tiles.  Kvazaar uses it for the fractional motion refinement and the rough intra
// search (satd_8x8 ... satd_64x64 in strategies/*/picture-*.c); this encoder's decisions use SAD
// (DESIGN.md section 5), so the kernel is exposed as a primitive: the SATD map of two planes at 8x8
// granularity, e.g. source vs reconstruction for rate control, or source vs previous source for
// scene-cut detection.
//
// Eight lanes per block: lane r holds row r of the residual, does the horizontal 8-point Hadamard in
// registers, the vertical one with three __shfl_xor butterflies, and the eight lanes add up.
#include <string.h>

#include "../../include/b200_hevc.h"
#include "../../include/b200media.h"
#include "hevc_device.cuh"
#include "runtime.h"

namespace b200 {
namespace {

__device__ __forceinline__ void hadamard8(int v[8])
{
#pragma unroll
  for (int s = 1; s < 8; s <<= 1)
#pragma unroll
    for (int i = 0; i < 8; i++)
      if (!(i & s)) { int a = v[i], b = v[i | s]; v[i] = a + b; v[i | s] = a - b; }
}

__global__ void __launch_bounds__(256)
k_satd8x8(const uint8_t *__restrict__ a, const uint8_t *__restrict__ b, int w, int h, uint32_t *__restrict__ out)
{
  const int w8 = w >> 3, blocks = w8 * (h >> 3);
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  const int blk = gid >> 3, r = gid & 7;
  const bool live = blk < blocks;
  int v[8];
  if (live) {
    const int bx = blk % w8, by = blk / w8;
    const size_t off = (size_t)(by * 8 + r) * w + bx * 8;
    const uint2 pa = __ldg((const uint2 *)(a + off)), pb = __ldg((const uint2 *)(b + off));
#pragma unroll
    for (int i = 0; i < 4; i++) {
      v[i] = (int)((pa.x >> (8 * i)) & 255) - (int)((pb.x >> (8 * i)) & 255);
      v[4 + i] = (int)((pa.y >> (8 * i)) & 255) - (int)((pb.y >> (8 * i)) & 255);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = 0;
  }
  hadamard8(v);
#pragma unroll
  for (int s = 1; s < 8; s <<= 1) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const int o = __shfl_xor_sync(0xffffffffu, v[i], s);
      v[i] = (r & s) ? o - v[i] : v[i] + o;
    }
  }
  unsigned sum = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) sum += (unsigned)abs(v[i]);
#pragma unroll
  for (int s = 1; s < 8; s <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, s);
  if (live && r == 0) out[blk] = (sum + 2) >> 2;
}

}  // namespace
}  // namespace b200

extern "C" {

int b200_satd8x8_dev(const uint8_t *d_a, const uint8_t *d_b, int width, int height, uint32_t *d_out, void *stream)
{
  if (!d_a || !d_b || !d_out || width <= 0 || height <= 0 || (width & 7) || (height & 7)) {
    b200::set_error("b200_satd8x8_dev: planes must be non-null with dimensions that are multiples of 8");
    return B200_ERR_ARG;
  }
  const int threads = (width >> 3) * (height >> 3) * 8;
  b200::k_satd8x8<<<(threads + 255) / 256, 256, 0, (cudaStream_t)stream>>>(d_a, d_b, width, height, d_out);
  b200::count_launch(1);
  B200_CHECK(cudaGetLastError(), "k_satd8x8 launch");
  return B200_OK;
}

int b200_satd8x8(const uint8_t *a, const uint8_t *b, int width, int height, uint32_t *out)
{
  if (!a || !b || !out || width <= 0 || height <= 0 || (width & 7) || (height & 7)) {
    b200::set_error("b200_satd8x8: planes must be non-null with dimensions that are multiples of 8");
    return B200_ERR_ARG;
  }
  if (b200_device_count() <= 0) { b200::set_error("no CUDA device: b200_satd8x8 has no CPU fallback"); return B200_ERR_CUDA; }
  const size_t n = (size_t)width * height, nb = (size_t)(width >> 3) * (height >> 3) * sizeof(uint32_t);
  b200::Scratch &s = b200::scratch();
  if (!s.ensure(2 * n, nb)) return B200_ERR_CUDA;
  memcpy(s.h_in, a, n);
  memcpy(s.h_in + n, b, n);
  B200_CHECK(cudaMemcpyAsync(s.d_in, s.h_in, 2 * n, cudaMemcpyHostToDevice, s.stream), "H2D planes");
  int rc = b200_satd8x8_dev(s.d_in, s.d_in + n, width, height, (uint32_t *)s.d_out, s.stream);
  if (rc != B200_OK) return rc;
  B200_CHECK(cudaMemcpyAsync(s.h_out, s.d_out, nb, cudaMemcpyDeviceToHost, s.stream), "D2H satd");
  B200_CHECK(cudaStreamSynchronize(s.stream), "sync");
  memcpy(out, s.h_out, nb);
  return B200_OK;
}

}  // extern "C"
