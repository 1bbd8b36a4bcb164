// Register-only microbenchmarks of the integer instructions the encoder kernels are built on
// ("peak INT32 / __vsadu4 / __dp4a rate is not published -> measure a register-only
// microbenchmark per instruction on the box and use that as the denominator").
//
// Every thread runs kChains independent dependency chains of one instruction, kIters times, with no
// memory traffic; 8 CTAs of 256 threads per SM.  The figure returned is thread-level instructions
// per second over the whole GPU (one VABSDIFF4 = one instruction = four byte differences).
#include <cuda_runtime.h>

#include "../../include/b200media.h"
#include "runtime.h"

namespace b200 {

namespace {

constexpr int kChains = 8, kIters = 4096, kPeakThreads = 256;

template <int kKind>
__device__ __forceinline__ unsigned op(unsigned a, unsigned b, unsigned c)
{
  unsigned d;
  if (kKind == 0) asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  else if (kKind == 1) asm volatile("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  else if (kKind == 2) asm volatile("dp2a.lo.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  else if (kKind == 3) asm volatile("mad.lo.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  else if (kKind == 4) asm volatile("shf.r.wrap.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(c), "r"(a), "r"(b & 31u));
  else asm volatile("add.u32 %0, %1, %2;" : "=r"(d) : "r"(c), "r"(a));
  return d;
}

template <int kKind>
__global__ void __launch_bounds__(kPeakThreads)
k_int_peak(unsigned *out, unsigned seed)
{
  unsigned acc[kChains];
  const unsigned a = seed * 2654435761u + threadIdx.x, b = (seed ^ 0x9e3779b9u) + blockIdx.x;
#pragma unroll
  for (int i = 0; i < kChains; i++) acc[i] = a + i;
  for (int it = 0; it < kIters; it++) {
#pragma unroll
    for (int i = 0; i < kChains; i++) acc[i] = op<kKind>(a, b, acc[i]);
  }
  unsigned s = 0;
#pragma unroll
  for (int i = 0; i < kChains; i++) s ^= acc[i];
  if (s == 0x12345678u) out[0] = s;                 // keeps the chains alive; practically never taken
}

template <int kKind>
double run_peak(int sms, unsigned *d_out)
{
  const int grid = sms * 8;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_int_peak<kKind><<<grid, kPeakThreads>>>(d_out, 1u);            // warm-up
  float best = 1e30f;
  for (int rep = 0; rep < 5; rep++) {
    cudaEventRecord(e0);
    k_int_peak<kKind><<<grid, kPeakThreads>>>(d_out, 2u + rep);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  count_launch(6);
  const double ops = (double)grid * kPeakThreads * kChains * kIters;
  return ops / (best * 1e-3);
}

}  // namespace
}  // namespace b200

extern "C" {

// kind: 0 vabsdiff4.add, 1 dp4a, 2 dp2a, 3 mad.lo.s32, 4 shf (funnel shift), 5 add.u32.
// Returns thread-level instructions per second on the current device, < 0 on error.
double b200_int_peak(int kind)
{
  if (b200_device_count() <= 0) { b200::set_error("no CUDA device: b200_int_peak measures the GPU"); return -1.0; }
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  unsigned *d_out = nullptr;
  if (cudaMalloc((void **)&d_out, 64) != cudaSuccess) { b200::set_error("b200_int_peak: cudaMalloc failed"); return -1.0; }
  double r = -1.0;
  switch (kind) {
  case 0: r = b200::run_peak<0>(sms, d_out); break;
  case 1: r = b200::run_peak<1>(sms, d_out); break;
  case 2: r = b200::run_peak<2>(sms, d_out); break;
  case 3: r = b200::run_peak<3>(sms, d_out); break;
  case 4: r = b200::run_peak<4>(sms, d_out); break;
  case 5: r = b200::run_peak<5>(sms, d_out); break;
  default: b200::set_error("b200_int_peak: unknown kind %d", kind);
  }
  cudaFree(d_out);
  return r;
}

}  // extern "C"
