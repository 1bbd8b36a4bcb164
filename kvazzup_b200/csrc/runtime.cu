// Runtime plumbing of libb200media.so: error reporting, device selection,
// launch accounting and the per-thread staging buffers used by the
// host-buffer (reference-shaped) entry points.
#include "runtime.h"

#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <stdlib.h>

namespace b200 {

// The encoder keeps one CUDA stream per picture in flight next to its main stream.  With the
// default of 8 hardware work queues, unrelated streams share a queue and a 3 ms entropy-coding
// kernel then blocks the next picture's motion search behind it.  Ask for the maximum (32)
// unless the application chose a value; this only has an effect when the library is loaded
// before the CUDA context is created (bench.py also sets it before importing torch).
namespace {
struct ConnectionsInit {
  ConnectionsInit() { setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0); }
} g_connections_init;
}  // namespace

static thread_local char t_err[512] = "";
std::atomic<unsigned long long> g_launches{0};

unsigned wait_event_flags(bool pipelined)
{
  bool block = !pipelined;
  if (const char *ev = getenv("B200_SYNC")) {
    if (!strcmp(ev, "spin")) block = false;
    else if (!strcmp(ev, "block")) block = true;
  }
  return cudaEventDisableTiming | (block ? cudaEventBlockingSync : 0u);
}

void set_error(const char *fmt, ...)
{
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_err, sizeof(t_err), fmt, ap);
  va_end(ap);
}

bool cuda_ok(cudaError_t e, const char *what)
{
  if (e == cudaSuccess) return true;
  set_error("%s: %s", what, cudaGetErrorString(e));
  return false;
}

bool Scratch::ensure(size_t in_bytes, size_t out_bytes)
{
  if (!stream && !cuda_ok(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking), "cudaStreamCreate"))
    return false;
  auto grow_dev = [](uint8_t *&p, size_t &cap, size_t need) {
    if (need <= cap) return true;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    size_t n = need + need / 4 + 256;
    if (!cuda_ok(cudaMalloc((void **)&p, n), "cudaMalloc")) return false;
    cap = n;
    return true;
  };
  auto grow_host = [](uint8_t *&p, size_t &cap, size_t need) {
    if (need <= cap) return true;
    if (p) cudaFreeHost(p);
    p = nullptr; cap = 0;
    size_t n = need + need / 4 + 256;
    if (!cuda_ok(cudaMallocHost((void **)&p, n), "cudaMallocHost")) return false;
    cap = n;
    return true;
  };
  return grow_dev(d_in, d_in_cap, in_bytes) && grow_dev(d_out, d_out_cap, out_bytes) &&
         grow_host(h_in, h_in_cap, in_bytes) && grow_host(h_out, h_out_cap, out_bytes);
}

Scratch::~Scratch()
{
  // Process teardown may already have destroyed the context; ignore errors.
  if (d_in) cudaFree(d_in);
  if (d_out) cudaFree(d_out);
  if (h_in) cudaFreeHost(h_in);
  if (h_out) cudaFreeHost(h_out);
  if (stream) cudaStreamDestroy(stream);
}

Scratch &scratch()
{
  static thread_local Scratch s;
  return s;
}

}  // namespace b200

extern "C" {

int b200_device_count(void)
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int b200_set_device(int device)
{
  B200_CHECK(cudaSetDevice(device), "cudaSetDevice");
  return B200_OK;
}

const char *b200_last_error(void) { return b200::t_err; }

const char *b200_version(void) { return "b200media 0.1 (sm_100a)"; }

unsigned long long b200_launch_count(void) { return b200::g_launches.load(); }

}  // extern "C"
