// Host-side plumbing shared by the libgr_cuda translation units: error reporting across the C
// ABI (no exceptions), launch accounting, per-device lookup tables, pinned staging.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/gr_cuda.h"

namespace grb {

extern thread_local std::string g_last_error;
extern thread_local int g_last_error_code;
extern std::atomic<unsigned long long> g_launches;

inline int set_error(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_last_error = buf;
  g_last_error_code = code;
  return code;
}

#define GRB_CUDA(expr)                                                                           \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess)                                                                       \
      return grb::set_error(GRCUDA_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                            __FILE__, __LINE__);                                                 \
  } while (0)

#define GRB_LAUNCH_CHECK()                                                                       \
  do {                                                                                           \
    grb::g_launches.fetch_add(1, std::memory_order_relaxed);                                     \
    cudaError_t _e = cudaGetLastError();                                                         \
    if (_e != cudaSuccess)                                                                       \
      return grb::set_error(GRCUDA_ECUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), \
                            __FILE__, __LINE__);                                                 \
  } while (0)

// device copies of the two reference lookup tables (generated header, see tools/gen_tables.py)
struct DeviceTables {
  float* atan = nullptr;      // [257]
  float* mmse_eff = nullptr;  // [129][8], coefficient applied to in[ii + i]
};
int get_tables(DeviceTables* out);  // for the current device
int sm_count();

// growable device buffer
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes) {
    if (bytes <= cap) return GRCUDA_OK;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) {
      p = nullptr;
      return set_error(GRCUDA_ENOMEM, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    }
    cap = bytes;
    return GRCUDA_OK;
  }
  template <class T> T* as() { return reinterpret_cast<T*>(p); }
  ~DevBuf() { if (p) cudaFree(p); }
};

// Pinned, double-buffered staging between pageable host memory and the device.  Pointers that
// are already page-locked (cudaHostAlloc / cudaHostRegister, e.g. torch pinned tensors) are
// DMA'd directly.
struct Stager {
  static const size_t kChunk = 8u << 20;
  void* pin[2] = {nullptr, nullptr};
  cudaEvent_t ev[2] = {nullptr, nullptr};
  int turn = 0;
  ~Stager();
  int init();
  static bool is_pinned(const void* p);
  int h2d(void* dst, const void* src, size_t bytes, cudaStream_t s);
  int d2h(void* dst, const void* src, size_t bytes, cudaStream_t s);  // synchronises s before returning
};

}  // namespace grb
