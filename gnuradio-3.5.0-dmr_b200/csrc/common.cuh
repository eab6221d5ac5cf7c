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
// cudaFuncAttributeMaxDynamicSharedMemorySize belongs to the KERNEL (per device), not to a plan: plans that share a
// kernel (every FIR block, every generic FFT) must only ever RAISE it, or a plan created later with a smaller tile
// makes the launches of an earlier one fail.  Thread safe.
cudaError_t raise_dynamic_smem(const void* kernel, size_t bytes);

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

// Per-stage device timing with CUDA events recorded on the launch stream (bench.py's roofline
// numbers come from here: the events bracket exactly the kernels of one stage).
struct EventProfiler {
  static const int kStages = 8;
  bool on = false;
  struct Span { int stage; cudaEvent_t a, b; };
  std::vector<Span> spans;
  std::vector<cudaEvent_t> pool;
  float ms[kStages] = {0};
  int launches[kStages] = {0};
  cudaEvent_t get() {
    if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
  }
  void begin(int stage, cudaStream_t s) {
    if (!on) return;
    Span sp; sp.stage = stage; sp.a = get(); sp.b = nullptr;
    cudaEventRecord(sp.a, s);
    spans.push_back(sp);
  }
  void end(cudaStream_t s, int nlaunch = 1) {
    if (!on || spans.empty()) return;
    Span& sp = spans.back();
    sp.b = get();
    cudaEventRecord(sp.b, s);
    launches[sp.stage] += nlaunch;
  }
  // synchronises the recorded events, accumulates, returns totals and resets
  void read(float* out_ms, int* out_launches) {
    for (Span& sp : spans) {
      if (sp.b) {
        cudaEventSynchronize(sp.b);
        float t = 0.f;
        if (cudaEventElapsedTime(&t, sp.a, sp.b) == cudaSuccess) ms[sp.stage] += t;
        pool.push_back(sp.b);
      }
      pool.push_back(sp.a);
    }
    spans.clear();
    for (int i = 0; i < kStages; i++) {
      if (out_ms) out_ms[i] = ms[i];
      if (out_launches) out_launches[i] = launches[i];
      ms[i] = 0.f;
      launches[i] = 0;
    }
  }
  ~EventProfiler() {
    for (Span& sp : spans) { cudaEventDestroy(sp.a); if (sp.b) cudaEventDestroy(sp.b); }
    for (cudaEvent_t e : pool) cudaEventDestroy(e);
  }
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
