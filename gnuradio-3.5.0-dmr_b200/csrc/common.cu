#include "common.cuh"

#include <map>
#include <mutex>
#include <utility>

#include "gr_tables.h"  // build/generated (tools/gen_tables.py)

namespace grb {

thread_local std::string g_last_error;
thread_local int g_last_error_code = 0;
std::atomic<unsigned long long> g_launches{0};

static std::mutex g_tab_mu;
static std::map<int, DeviceTables> g_tabs;
static std::map<int, int> g_sms;

int get_tables(DeviceTables* out) {
  int dev = 0;
  GRB_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(g_tab_mu);
  auto it = g_tabs.find(dev);
  if (it == g_tabs.end()) {
    DeviceTables t;
    float atan_h[257], raw[129 * 8], eff[129 * 8];
    memcpy(atan_h, GR_FAST_ATAN_TABLE_BITS, sizeof atan_h);
    memcpy(raw, GR_MMSE_TAPS_BITS, sizeof raw);
    // gri_mmse_fir_interpolator.cc:38-41 hands taps[imu] to gr_fir_fff, which stores them
    // reversed (gr_fir_XXX.h.t:65): the coefficient applied to input[i] is taps[imu][7-i].
    for (int s = 0; s < 129; s++)
      for (int i = 0; i < 8; i++) eff[s * 8 + i] = raw[s * 8 + (7 - i)];
    GRB_CUDA(cudaMalloc(&t.atan, sizeof atan_h));
    GRB_CUDA(cudaMalloc(&t.mmse_eff, sizeof eff));
    GRB_CUDA(cudaMemcpy(t.atan, atan_h, sizeof atan_h, cudaMemcpyHostToDevice));
    GRB_CUDA(cudaMemcpy(t.mmse_eff, eff, sizeof eff, cudaMemcpyHostToDevice));
    it = g_tabs.emplace(dev, t).first;
  }
  *out = it->second;
  return GRCUDA_OK;
}

cudaError_t raise_dynamic_smem(const void* kernel, size_t bytes) {
  static std::mutex mu;
  static std::map<std::pair<int, const void*>, size_t> cur;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lk(mu);
  size_t& have = cur[std::make_pair(dev, kernel)];
  if (bytes <= have) return cudaSuccess;
  e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e == cudaSuccess) have = bytes;
  return e;
}

int sm_count() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  std::lock_guard<std::mutex> lk(g_tab_mu);
  auto it = g_sms.find(dev);
  if (it != g_sms.end()) return it->second;
  int n = 148;
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  g_sms[dev] = n;
  return n;
}

Stager::~Stager() {
  for (int i = 0; i < 2; i++) {
    if (pin[i]) cudaFreeHost(pin[i]);
    if (ev[i]) cudaEventDestroy(ev[i]);
  }
}

int Stager::init() {
  if (pin[0]) return GRCUDA_OK;
  for (int i = 0; i < 2; i++) {
    GRB_CUDA(cudaHostAlloc(&pin[i], kChunk, cudaHostAllocDefault));
    GRB_CUDA(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming));
  }
  return GRCUDA_OK;
}

bool Stager::is_pinned(const void* p) {
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return attr.type == cudaMemoryTypeHost;
}

int Stager::h2d(void* dst, const void* src, size_t bytes, cudaStream_t s) {
  if (bytes == 0) return GRCUDA_OK;
  if (is_pinned(src)) {
    GRB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, s));
    return GRCUDA_OK;
  }
  int rc = init();
  if (rc) return rc;
  size_t off = 0;
  while (off < bytes) {  // CPU copy of chunk i+1 overlaps the DMA of chunk i
    const size_t n = std::min(kChunk, bytes - off);
    GRB_CUDA(cudaEventSynchronize(ev[turn]));
    memcpy(pin[turn], (const char*)src + off, n);
    GRB_CUDA(cudaMemcpyAsync((char*)dst + off, pin[turn], n, cudaMemcpyHostToDevice, s));
    GRB_CUDA(cudaEventRecord(ev[turn], s));
    turn ^= 1;
    off += n;
  }
  return GRCUDA_OK;
}

int Stager::d2h(void* dst, const void* src, size_t bytes, cudaStream_t s) {
  if (bytes == 0) {
    GRB_CUDA(cudaStreamSynchronize(s));
    return GRCUDA_OK;
  }
  if (is_pinned(dst)) {
    GRB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, s));
    GRB_CUDA(cudaStreamSynchronize(s));
    return GRCUDA_OK;
  }
  int rc = init();
  if (rc) return rc;
  size_t off = 0;
  size_t pend_off[2] = {0, 0}, pend_n[2] = {0, 0};
  while (off < bytes || pend_n[0] || pend_n[1]) {
    if (pend_n[turn]) {  // drain the buffer we are about to reuse
      GRB_CUDA(cudaEventSynchronize(ev[turn]));
      memcpy((char*)dst + pend_off[turn], pin[turn], pend_n[turn]);
      pend_n[turn] = 0;
    }
    if (off < bytes) {
      const size_t n = std::min(kChunk, bytes - off);
      GRB_CUDA(cudaMemcpyAsync(pin[turn], (const char*)src + off, n, cudaMemcpyDeviceToHost, s));
      GRB_CUDA(cudaEventRecord(ev[turn], s));
      pend_off[turn] = off;
      pend_n[turn] = n;
      off += n;
    }
    turn ^= 1;
  }
  GRB_CUDA(cudaStreamSynchronize(s));
  return GRCUDA_OK;
}

}  // namespace grb
