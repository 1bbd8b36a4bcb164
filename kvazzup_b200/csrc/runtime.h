// Shared host-side runtime helpers for libb200media.so (internal header).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <atomic>

#include "../../include/b200media.h"

namespace b200 {

void set_error(const char *fmt, ...);
bool cuda_ok(cudaError_t e, const char *what);

// Flags for an event a host thread waits on.  A pipelined stream (pictures in flight) wants the lowest
// wake-up latency and spins; a call-shaped stream (one picture at a time, dozens of streams per GPU, one
// host thread each) must not burn a core while its picture is on the GPU: the thread sleeps on the
// event instead (cudaEventBlockingSync).  B200_SYNC=spin|block overrides the choice.
unsigned wait_event_flags(bool pipelined);

extern std::atomic<unsigned long long> g_launches;
inline void count_launch(unsigned n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// Grow-only per-thread device + pinned staging buffers for the host-buffer
// entry points (one media stream = one host thread in the Filter model).
struct Scratch {
  uint8_t *d_in = nullptr;  size_t d_in_cap = 0;
  uint8_t *d_out = nullptr; size_t d_out_cap = 0;
  uint8_t *h_in = nullptr;  size_t h_in_cap = 0;   // pinned
  uint8_t *h_out = nullptr; size_t h_out_cap = 0;  // pinned
  cudaStream_t stream = nullptr;
  bool ensure(size_t in_bytes, size_t out_bytes);
  ~Scratch();
};
Scratch &scratch();

#define B200_CHECK(expr, what)                     \
  do {                                             \
    if (!::b200::cuda_ok((expr), (what))) return B200_ERR_CUDA; \
  } while (0)

}  // namespace b200
