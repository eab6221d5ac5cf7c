// libgr_cuda: implementation of the C ABI declared in include/gr_cuda.h.
// Host logic of the plans (tap bookkeeping, history/"updated" contracts of the reference blocks,
// staging) + kernel launches.  No CPU compute path exists here: every work call runs CUDA
// kernels or fails with GRCUDA_ECUDA.
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdlib>

#include "common.cuh"
#include "fft_plan.h"
#include "internal.h"
#include "kernels_demod.cuh"
#include "kernel_mm.cuh"
#include "kernel_mm_quad.cuh"
#include "kernels_fir.cuh"
#include "kernel_fft_filter.cuh"

using namespace grb;

namespace {

struct PinBuf {
  void* p = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes) {
    if (bytes <= cap) return GRCUDA_OK;
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
    cudaError_t e = cudaHostAlloc(&p, bytes, cudaHostAllocDefault);
    if (e != cudaSuccess) return set_error(GRCUDA_ENOMEM, "cudaHostAlloc(%zu): %s", bytes, cudaGetErrorString(e));
    cap = bytes;
    return GRCUDA_OK;
  }
  ~PinBuf() { if (p) cudaFreeHost(p); }
};

struct PlanBase {
  cudaStream_t stream = nullptr;
  Stager stager;
  DevBuf d_in, d_out;
  std::mutex mu;  // guards setter-visible state (reference setters are unlocked; ours are not)
  int base_init() {
    GRB_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    return GRCUDA_OK;
  }
  cudaStream_t pick(void* s) const { return s ? (cudaStream_t)s : stream; }
  virtual ~PlanBase() { if (stream) cudaStreamDestroy(stream); }
};

bool device_ok() {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    set_error(GRCUDA_ECUDA, "no CUDA device available (%s); libgr_cuda has no CPU fallback",
              e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    cudaGetLastError();
    return false;
  }
  return true;
}

int grid_for(long items, int threads, int per_sm = 8) {
  long g = (items + threads - 1) / threads;
  long cap = (long)sm_count() * per_sm;
  return (int)std::max<long>(1, std::min(g, cap));
}

}  // namespace

// =============================================================================================
// library / device
// =============================================================================================
extern "C" {

const char* grcuda_version(void) { return "gr-b200 0.1 (sm_100a)"; }
const char* grcuda_last_error(void) { return g_last_error.c_str(); }
int grcuda_last_error_code(void) { return g_last_error_code; }
int grcuda_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}
int grcuda_set_device(int device) { GRB_CUDA(cudaSetDevice(device)); return GRCUDA_OK; }
int grcuda_device_synchronize(void) { GRB_CUDA(cudaDeviceSynchronize()); return GRCUDA_OK; }
void* grcuda_malloc_device(size_t bytes) {
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes ? bytes : 1);
  if (e != cudaSuccess) { set_error(GRCUDA_ENOMEM, "cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e)); return nullptr; }
  return p;
}
void grcuda_free_device(void* p) { if (p) cudaFree(p); }
void* grcuda_malloc_pinned(size_t bytes) {
  void* p = nullptr;
  cudaError_t e = cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault);
  if (e != cudaSuccess) { set_error(GRCUDA_ENOMEM, "cudaHostAlloc(%zu): %s", bytes, cudaGetErrorString(e)); return nullptr; }
  return p;
}
void grcuda_free_pinned(void* p) { if (p) cudaFreeHost(p); }
int grcuda_memcpy_h2d(void* dst, const void* src, size_t bytes, void* stream) {
  GRB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
  if (!stream) GRB_CUDA(cudaStreamSynchronize(nullptr));
  return GRCUDA_OK;
}
int grcuda_memcpy_d2h(void* dst, const void* src, size_t bytes, void* stream) {
  GRB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  if (!stream) GRB_CUDA(cudaStreamSynchronize(nullptr));
  return GRCUDA_OK;
}
int grcuda_stream_synchronize(void* stream) { GRB_CUDA(cudaStreamSynchronize((cudaStream_t)stream)); return GRCUDA_OK; }
unsigned long long grcuda_kernel_launch_count(void) { return g_launches.load(); }

}  // extern "C"

// =============================================================================================
// a1 / a3: decimating FIR with real or complex taps
// =============================================================================================
struct FirCore {
  int decim = 1, ntaps = 0, J = 8;
  bool ctaps = false;
  DevBuf d_rtp;
  int threads = 128, tile_out = 1024, pitch = 0;
  size_t smem = 0;
  int max_ctas = 0;

  // rt: reversed taps (float or complex<float>), polyphase-split and zero padded on upload
  int upload(const float* rt, int n, bool complex_taps) {
    ntaps = n;
    ctaps = complex_taps;
    const int D = decim;
    const int jraw = std::max(1, (n + D - 1) / D);
    J = ((jraw + FIR_R - 1) / FIR_R) * FIR_R;
    const int w = complex_taps ? 2 : 1;
    std::vector<float> rtp((size_t)D * J * w, 0.f);
    for (int i = 0; i < n; i++) {
      const int q = i / D, p = i % D;
      for (int c = 0; c < w; c++) rtp[((size_t)p * J + q) * w + c] = rt[(size_t)i * w + c];
    }
    int rc = d_rtp.reserve(rtp.size() * sizeof(float));
    if (rc) return rc;
    GRB_CUDA(cudaMemcpy(d_rtp.p, rtp.data(), rtp.size() * sizeof(float), cudaMemcpyHostToDevice));
    for (threads = 128; threads >= 32; threads /= 2) {
      tile_out = threads * FIR_R;
      const int rows_n = tile_out + J + FIR_R;
      pitch = rows_n + (rows_n >> 3) + 1;
      pitch += (4 - (pitch % 16) + 16) % 16;  // pitch = 4 (mod 16): conflict-free phase-split fill
      smem = (size_t)D * pitch * sizeof(float2) + rtp.size() * sizeof(float);
      if (smem <= 200 * 1024) break;
    }
    if (smem > 200 * 1024)
      return set_error(GRCUDA_EUNSUPPORTED, "FIR with %d taps x decimation %d exceeds the shared-memory tile", n, D);
    const void* k = complex_taps ? (const void*)fir_decim_kernel<true> : (const void*)fir_decim_kernel<false>;
    GRB_CUDA(raise_dynamic_smem(k, smem));
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, threads, smem);
    max_ctas = std::max(1, per_sm) * sm_count();
    return GRCUDA_OK;
  }

  int launch(const float2* d_in, float2* d_out, long nout, bool rotate, double theta, long out_index0,
             cudaStream_t s) {
    if (nout <= 0) return GRCUDA_OK;
    if (ntaps == 0) {  // gr_fir_ccf_simd::filter returns 0 for an empty filter (gr_fir_ccf_simd.cc:103-104)
      GRB_CUDA(cudaMemsetAsync(d_out, 0, nout * sizeof(float2), s));
      return GRCUDA_OK;
    }
    FirArgs a;
    a.in = d_in; a.out = d_out; a.nout = nout; a.decim = decim; a.ntaps = ntaps; a.J = J;
    a.rtp = d_rtp.as<float>(); a.tile_out = tile_out; a.pitch = pitch;
    a.rotate = rotate ? 1 : 0; a.theta = theta; a.out_index0 = out_index0;
    for (int r = 0; r < FIR_R; r++) a.rot_step[r] = make_float2((float)cos(theta * r), (float)sin(theta * r));
    const long ntiles = (nout + tile_out - 1) / tile_out;
    const int grid = (int)std::min<long>(ntiles, max_ctas);
    if (ctaps) fir_decim_kernel<true><<<grid, threads, smem, s>>>(a);
    else fir_decim_kernel<false><<<grid, threads, smem, s>>>(a);
    GRB_LAUNCH_CHECK();
    return GRCUDA_OK;
  }
};

struct grcuda_fir_ccf : PlanBase {
  FirCore core;
  std::vector<float> new_taps;
  bool updated = false;
  unsigned history = 1;
  int set_now(const std::vector<float>& taps) {
    cudaDeviceSynchronize();  // no kernel may still be reading the tap store we are replacing
    std::vector<float> rt(taps.rbegin(), taps.rend());  // gr_fir_XXX.h.t:65 d_taps = reverse(taps)
    history = (unsigned)taps.size();                     // gr_fir_filter_XXX.cc.t:51 set_history(ntaps)
    return core.upload(rt.data(), (int)rt.size(), false);
  }
};

extern "C" {

grcuda_fir_ccf* grcuda_fir_filter_ccf_create(int decimation, const float* taps, int ntaps) {
  if (decimation < 1 || ntaps < 0) { set_error(GRCUDA_EINVAL, "fir_filter_ccf: bad decimation/ntaps"); return nullptr; }
  if (!device_ok()) return nullptr;
  grcuda_fir_ccf* h = new grcuda_fir_ccf;
  h->core.decim = decimation;
  if (h->base_init() || h->set_now(std::vector<float>(taps, taps + ntaps))) { delete h; return nullptr; }
  return h;
}
void grcuda_fir_filter_ccf_destroy(grcuda_fir_ccf* h) { delete h; }
int grcuda_fir_filter_ccf_set_taps(grcuda_fir_ccf* h, const float* taps, int ntaps) {
  std::lock_guard<std::mutex> lk(h->mu);  // gr_fir_filter_XXX.cc.t:59-64: deferred to the next work()
  h->new_taps.assign(taps, taps + ntaps);
  h->updated = true;
  return GRCUDA_OK;
}
unsigned grcuda_fir_filter_ccf_history(grcuda_fir_ccf* h) { return h->history; }
int grcuda_fir_filter_ccf_decimation(grcuda_fir_ccf* h) { return h->core.decim; }
int grcuda_fir_filter_ccf_work_device(grcuda_fir_ccf* h, long nout, const grcuda_complex* d_in, grcuda_complex* d_out,
                                      void* stream) {
  {  // a pending set_taps takes effect here too (set_now synchronises the device before it replaces the tap store)
    std::lock_guard<std::mutex> lk(h->mu);
    if (h->updated) {
      h->updated = false;
      int rc = h->set_now(h->new_taps);
      if (rc) return rc;
    }
  }
  return h->core.launch((const float2*)d_in, (float2*)d_out, nout, false, 0.0, 0, h->pick(stream));
}
int grcuda_fir_filter_ccf_work(grcuda_fir_ccf* h, int nout, const grcuda_complex* in, grcuda_complex* out) {
  {
    std::lock_guard<std::mutex> lk(h->mu);
    if (h->updated) {  // gr_fir_filter_XXX.cc.t:74-79
      h->updated = false;
      int rc = h->set_now(h->new_taps);
      return rc ? rc : 0;
    }
  }
  if (nout <= 0) return 0;
  const size_t nin = (size_t)(nout - 1) * h->core.decim + std::max(h->core.ntaps, 1);
  int rc;
  if ((rc = h->d_in.reserve(nin * sizeof(float2))) || (rc = h->d_out.reserve((size_t)nout * sizeof(float2)))) return rc;
  if ((rc = h->stager.h2d(h->d_in.p, in, nin * sizeof(float2), h->stream))) return rc;
  if ((rc = h->core.launch(h->d_in.as<float2>(), h->d_out.as<float2>(), nout, false, 0.0, 0, h->stream))) return rc;
  if ((rc = h->stager.d2h(out, h->d_out.p, (size_t)nout * sizeof(float2), h->stream))) return rc;
  return nout;
}

}  // extern "C"

// ---- a3 gr_freq_xlating_fir_filter_ccf ------------------------------------------------------------
struct grcuda_fxlat : PlanBase {
  FirCore core;
  std::vector<float> proto;
  double center_freq = 0, sampling_freq = 1;
  bool updated = false;
  unsigned history = 1;
  double theta = 0;      // angle of the normalised float phase increment (gr_rotator.h:38)
  long out_count = 0;    // d_counter of the rotator = outputs produced so far

  int build() {  // build_composite_fir (gr_freq_xlating_fir_filter_XXX.cc.t:72-83), same float ops
    cudaDeviceSynchronize();
    const int n = (int)proto.size();
    std::vector<std::complex<float> > ctaps(n);
    const float fwT0 = 2 * M_PI * center_freq / sampling_freq;
    for (int i = 0; i < n; i++) ctaps[i] = proto[i] * std::exp(std::complex<float>(0, i * fwT0));
    // set_taps(gr_reverse(ctaps)) and the FIR stores reverse(taps): filter() uses ctaps in proto order
    std::vector<float> rt((size_t)n * 2);
    for (int i = 0; i < n; i++) { rt[2 * i] = ctaps[i].real(); rt[2 * i + 1] = ctaps[i].imag(); }
    std::complex<float> incr = std::exp(std::complex<float>(0, fwT0 * core.decim));
    incr = incr / std::abs(incr);  // set_phase_incr
    theta = std::atan2((double)incr.imag(), (double)incr.real());
    history = (unsigned)n;
    return core.upload(rt.data(), n, true);
  }
};

extern "C" {

grcuda_fxlat* grcuda_freq_xlating_fir_filter_ccf_create(int decimation, const float* taps, int ntaps,
                                                       double center_freq, double sampling_freq) {
  if (decimation < 1 || ntaps < 0) { set_error(GRCUDA_EINVAL, "freq_xlating_fir_filter_ccf: bad arguments"); return nullptr; }
  if (!device_ok()) return nullptr;
  grcuda_fxlat* h = new grcuda_fxlat;
  h->core.decim = decimation;
  h->proto.assign(taps, taps + ntaps);
  h->center_freq = center_freq;
  h->sampling_freq = sampling_freq;
  if (h->base_init() || h->build()) { delete h; return nullptr; }
  return h;
}
void grcuda_freq_xlating_fir_filter_ccf_destroy(grcuda_fxlat* h) { delete h; }
int grcuda_freq_xlating_fir_filter_ccf_set_taps(grcuda_fxlat* h, const float* taps, int ntaps) {
  std::lock_guard<std::mutex> lk(h->mu);
  h->proto.assign(taps, taps + ntaps);
  h->updated = true;
  return GRCUDA_OK;
}
int grcuda_freq_xlating_fir_filter_ccf_set_center_freq(grcuda_fxlat* h, double f) {
  std::lock_guard<std::mutex> lk(h->mu);
  h->center_freq = f;
  h->updated = true;
  return GRCUDA_OK;
}
unsigned grcuda_freq_xlating_fir_filter_ccf_history(grcuda_fxlat* h) { return h->history; }
int grcuda_freq_xlating_fir_filter_ccf_work_device(grcuda_fxlat* h, long nout, const grcuda_complex* d_in,
                                                   grcuda_complex* d_out, void* stream) {
  {
    std::lock_guard<std::mutex> lk(h->mu);
    if (h->updated) {  // pending set_taps / set_center_freq
      h->updated = false;
      int rc0 = h->build();
      if (rc0) return rc0;
    }
  }
  int rc = h->core.launch((const float2*)d_in, (float2*)d_out, nout, true, h->theta, h->out_count, h->pick(stream));
  if (rc == GRCUDA_OK) h->out_count += nout;
  return rc;
}
int grcuda_freq_xlating_fir_filter_ccf_work(grcuda_fxlat* h, int nout, const grcuda_complex* in, grcuda_complex* out) {
  {
    std::lock_guard<std::mutex> lk(h->mu);
    if (h->updated) {  // :107-112 (the rotator keeps its phase across a rebuild: only incr changes)
      h->updated = false;
      int rc = h->build();
      return rc ? rc : 0;
    }
  }
  if (nout <= 0) return 0;
  const size_t nin = (size_t)(nout - 1) * h->core.decim + std::max(h->core.ntaps, 1);
  int rc;
  if ((rc = h->d_in.reserve(nin * sizeof(float2))) || (rc = h->d_out.reserve((size_t)nout * sizeof(float2)))) return rc;
  if ((rc = h->stager.h2d(h->d_in.p, in, nin * sizeof(float2), h->stream))) return rc;
  if ((rc = grcuda_freq_xlating_fir_filter_ccf_work_device(h, nout, (const grcuda_complex*)h->d_in.p,
                                                           (grcuda_complex*)h->d_out.p, h->stream)))
    return rc;
  if ((rc = h->stager.d2h(out, h->d_out.p, (size_t)nout * sizeof(float2), h->stream))) return rc;
  return nout;
}

}  // extern "C"

// =============================================================================================
// a2: gr_fir_filter_fff (single stream and batched [time][channel])
// =============================================================================================
namespace grb {
__global__ void fir_fff_stream_kernel(const float* __restrict__ in, float* __restrict__ out, long nout, int decim,
                                      const float* __restrict__ rt_g, int ntaps, int order, long abs0) {
  extern __shared__ float rt[];
  for (int i = threadIdx.x; i < ntaps; i += blockDim.x) rt[i] = rt_g[i];
  __syncthreads();
  for (long o = blockIdx.x * (long)blockDim.x + threadIdx.x; o < nout; o += (long)gridDim.x * blockDim.x) {
    const float* p = in + o * decim;
    out[o] = (order == GR_ORDER_SSE) ? dot_sse(rt, ntaps, p, 1, mod4(abs0 + o * decim)) : dot_generic(rt, ntaps, p, 1);
  }
}
}  // namespace grb

struct grcuda_fir_fff : PlanBase {
  int decim = 1, ntaps = 0, order = GRCUDA_ORDER_SSE;
  DevBuf d_rt, d_front_tp;
  std::vector<float> h_front_tp;   // host image of d_front_tp
  bool has_front_tp = false;
  std::vector<float> new_taps, rt_host;
  bool updated = false;
  unsigned history = 1;
  int set_now(const std::vector<float>& taps) {
    cudaDeviceSynchronize();
    std::vector<float> rt(taps.rbegin(), taps.rend());
    rt_host = rt;
    ntaps = (int)rt.size();
    history = (unsigned)ntaps;
    if ((size_t)ntaps * sizeof(float) > 96 * 1024)
      return set_error(GRCUDA_EUNSUPPORTED, "fir_filter_fff: %d taps exceed the shared-memory tap store", ntaps);
    int rc = d_rt.reserve(std::max<size_t>(4, rt.size() * sizeof(float)));
    if (rc) return rc;
    if (ntaps) GRB_CUDA(cudaMemcpy(d_rt.p, rt.data(), rt.size() * sizeof(float), cudaMemcpyHostToDevice));
    has_front_tp = false;
    if (ntaps >= 1 && ntaps <= demod_front_max_taps()) {  // tap store of the fused discriminator + FIR kernel
      const std::vector<float> tp = demod_front_tap_table(rt.data(), ntaps);
      if ((rc = d_front_tp.reserve(tp.size() * sizeof(float)))) return rc;
      GRB_CUDA(cudaMemcpy(d_front_tp.p, tp.data(), tp.size() * sizeof(float), cudaMemcpyHostToDevice));
      h_front_tp = tp;
      has_front_tp = true;
    }
    if ((size_t)ntaps * sizeof(float) > 48 * 1024) {
      GRB_CUDA(raise_dynamic_smem((const void*)fir_fff_kernel, 96 * 1024));
      GRB_CUDA(raise_dynamic_smem((const void*)fir_fff_stream_kernel, 96 * 1024));
    }
    return GRCUDA_OK;
  }
  int launch(const float* d_in, float* d_out, long nout, int nchan, long abs0, cudaStream_t s) {
    if (nout <= 0 || nchan <= 0) return GRCUDA_OK;
    const size_t smem = std::max<size_t>(4, (size_t)ntaps * sizeof(float));
    if (nchan == 1) {
      fir_fff_stream_kernel<<<grid_for(nout, 256), 256, smem, s>>>(d_in, d_out, nout, decim, d_rt.as<float>(), ntaps,
                                                                    order, abs0);
    } else {
      const int gx = (nchan + 127) / 128;
      // enough row tiles to fill the machine ~4x, at least 16 rows each
      long tiles = std::max<long>(1, (long)sm_count() * 16 / gx);
      int rpt = (int)std::max<long>(16, (nout + tiles - 1) / tiles);
      dim3 grid(gx, (unsigned)((nout + rpt - 1) / rpt));
      fir_fff_kernel<<<grid, 128, smem, s>>>(d_in, d_out, nout, nchan, decim, d_rt.as<float>(), ntaps, order, abs0, rpt);
    }
    GRB_LAUNCH_CHECK();
    return GRCUDA_OK;
  }
};

extern "C" {

grcuda_fir_fff* grcuda_fir_filter_fff_create(int decimation, const float* taps, int ntaps, int order) {
  if (decimation < 1 || ntaps < 0 || (order != GRCUDA_ORDER_GENERIC && order != GRCUDA_ORDER_SSE)) {
    set_error(GRCUDA_EINVAL, "fir_filter_fff: bad arguments");
    return nullptr;
  }
  if (!device_ok()) return nullptr;
  grcuda_fir_fff* h = new grcuda_fir_fff;
  h->decim = decimation;
  h->order = order;
  if (h->base_init() || h->set_now(std::vector<float>(taps, taps + ntaps))) { delete h; return nullptr; }
  return h;
}
void grcuda_fir_filter_fff_destroy(grcuda_fir_fff* h) { delete h; }
int grcuda_fir_filter_fff_set_taps(grcuda_fir_fff* h, const float* taps, int ntaps) {
  std::lock_guard<std::mutex> lk(h->mu);
  h->new_taps.assign(taps, taps + ntaps);
  h->updated = true;
  return GRCUDA_OK;
}
unsigned grcuda_fir_filter_fff_history(grcuda_fir_fff* h) { return h->history; }
int grcuda_fir_filter_fff_work_device(grcuda_fir_fff* h, long nout, int nchan, const float* d_in, float* d_out,
                                      long abs_index0, void* stream) {
  {
    std::lock_guard<std::mutex> lk(h->mu);
    if (h->updated) {
      h->updated = false;
      int rc = h->set_now(h->new_taps);
      if (rc) return rc;
    }
  }
  return h->launch(d_in, d_out, nout, nchan, abs_index0, h->pick(stream));
}
int grcuda_fir_filter_fff_work(grcuda_fir_fff* h, int nout, const float* in, float* out, long abs_index0) {
  {
    std::lock_guard<std::mutex> lk(h->mu);
    if (h->updated) {
      h->updated = false;
      int rc = h->set_now(h->new_taps);
      return rc ? rc : 0;
    }
  }
  if (nout <= 0) return 0;
  const size_t nin = (size_t)(nout - 1) * h->decim + std::max(h->ntaps, 1);
  int rc;
  if ((rc = h->d_in.reserve(nin * sizeof(float))) || (rc = h->d_out.reserve((size_t)nout * sizeof(float)))) return rc;
  if ((rc = h->stager.h2d(h->d_in.p, in, nin * sizeof(float), h->stream))) return rc;
  if ((rc = h->launch(h->d_in.as<float>(), h->d_out.as<float>(), nout, 1, abs_index0, h->stream))) return rc;
  if ((rc = h->stager.d2h(out, h->d_out.p, (size_t)nout * sizeof(float), h->stream))) return rc;
  return nout;
}

}  // extern "C"

// =============================================================================================
// a4 / a5 / a14: gr_pfb_channelizer_ccf
// =============================================================================================
struct grcuda_pfb : PlanBase {
  unsigned M = 0;
  float os = 1.f;
  int rr = 0, output_multiple = 1, T = 0, TT = 0, ntaps = 0;
  double relative_rate = 1.0;
  unsigned history = 1;
  bool updated = false;
  std::vector<float> pending;
  DevBuf d_taps_t, d_taps_plain, d_u, d_stage, d_rows;
  PinBuf pin_in;
  FftPlan* fft = nullptr;
  long chunk_rows = 0;
  EventProfiler prof;  // stage 0 = branch FIR kernel, stage 1 = FFT kernel

  ~grcuda_pfb() { fft_plan_destroy(fft); }

  bool dirty = false;
  // host-visible part of set_taps (:104-139): geometry changes immediately, like set_history (:136)
  void set_geometry(const std::vector<float>& taps) {
    pending = taps;
    ntaps = (int)taps.size();
    T = (int)std::ceil((double)ntaps / (double)M);
    if (T < 1) T = 1;  // keeps the layout valid for an empty prototype (all-zero branches)
    TT = T <= 4 ? 4 : (T <= 8 ? 8 : (T <= 16 ? 16 : (T <= 32 ? 32 : 0)));
    history = (unsigned)T + 1;  // :136
    dirty = true;
  }
  // device part, applied at the next work boundary so that no in-flight kernel sees a torn tap set
  int upload_if_dirty() {
    std::lock_guard<std::mutex> lk(mu);
    if (!dirty) return GRCUDA_OK;
    cudaDeviceSynchronize();
    dirty = false;
    const std::vector<float>& taps = pending;
    std::vector<float> plain((size_t)T * M, 0.f);
    for (int i = 0; i < ntaps; i++) plain[i] = taps[i];  // h[k + t*M] at plain[t*M + k]
    int rc = d_taps_plain.reserve(plain.size() * sizeof(float));
    if (rc) return rc;
    GRB_CUDA(cudaMemcpy(d_taps_plain.p, plain.data(), plain.size() * sizeof(float), cudaMemcpyHostToDevice));
    if (TT) {
      std::vector<float> tt((size_t)TT * M, 0.f);
      for (int t = 0; t < T; t++)
        for (unsigned j = 0; j < M; j++) tt[(size_t)t * M + j] = plain[(size_t)t * M + (M - 1 - j)];
      if ((rc = d_taps_t.reserve(tt.size() * sizeof(float)))) return rc;
      GRB_CUDA(cudaMemcpy(d_taps_t.p, tt.data(), tt.size() * sizeof(float), cudaMemcpyHostToDevice));
    }
    return GRCUDA_OK;
  }

  // rows in: [T + nin][M]; out: [nout][M].  nout output vectors; os == 1 -> nin == nout.
  // demod != nullptr: the FFT kernel applies gr_quadrature_demod_cf and writes D rows (float) instead of the
  // channelizer output (kernel_fft_demod.cuh); one pass over the whole call (no row chunking).
  struct DemodOut { float* d; float gain; const float* atan; const float2* prev_y; float2* last_y; bool coresident; float2* y_out; };
  int run(const float2* d_rows_in, float2* d_out, long nout, cudaStream_t s, const DemodOut* demod = nullptr) {
    int rc0 = upload_if_dirty();
    if (rc0) return rc0;
    if (nout <= 0) return GRCUDA_OK;
    const bool fast = (rr == (int)M) && TT;
    // bulk copies need 16-byte aligned row segments: even M (8-byte samples), 16-byte aligned base
    const bool tma_ok = fast && TT <= 16 && (M % 2 == 0) && (((uintptr_t)d_rows_in & 15) == 0);
    if (tma_ok) {
      GRB_CUDA(raise_dynamic_smem((const void*)pfb_fir_tma_kernel<4>, (size_t)pfb_fir_tma_smem()));
      GRB_CUDA(raise_dynamic_smem((const void*)pfb_fir_tma_kernel<8>, (size_t)pfb_fir_tma_smem()));
      GRB_CUDA(raise_dynamic_smem((const void*)pfb_fir_tma_kernel<16>, (size_t)pfb_fir_tma_smem()));
    }
    // The branch-filter output u is only an intermediate: process in row chunks small enough
    // to stay resident in the 126 MB L2 between the FIR kernel and the FFT kernel.
    long crow = chunk_rows;
    if (crow <= 0) {
      size_t mb = 1u << 20;  // default: no chunking (measured faster than L2-resident chunks, profiles/r1_pfb_chunking.txt)
      if (const char* e = getenv("GRCUDA_PFB_CHUNK_MB")) mb = (size_t)std::max(1, atoi(e));
      crow = std::max<long>(1, (long)(mb << 20) / (long)(M * sizeof(float2)));
    }
    if (!fast || demod) crow = nout;  // the oversampled path indexes rows through the output number
    crow = std::min(crow, nout);
    int rc = d_u.reserve((size_t)crow * M * sizeof(float2));
    if (rc) return rc;
    for (long r0 = 0; r0 < nout; r0 += crow) {
      const long n = std::min(crow, nout - r0);
      prof.begin(0, s);
      if (fast && tma_ok) {
        // TMA-staged branch filter: 256-column tiles, grid sized to whole waves of one CTA per SM
        PfbFirArgs a;
        a.x = d_rows_in + r0 * (long)M; a.u = d_u.as<float2>(); a.taps_t = d_taps_t.as<float>();
        a.M = (int)M; a.T = T; a.nrows = n;
        const int gx = ((int)M + PFT_COLS - 1) / PFT_COLS;
        long gy = std::max<long>(1, (n + 383) / 384);
        const long waves = (gx * gy + sm_count() - 1) / sm_count();
        gy = std::max<long>(1, std::min<long>(n, waves * sm_count() / gx));
        a.rows_per_thread = (int)((n + gy - 1) / gy);
        dim3 grid(gx, (unsigned)((n + a.rows_per_thread - 1) / a.rows_per_thread));
        const size_t smem = pfb_fir_tma_smem();
        switch (TT) {
          case 4: pfb_fir_tma_kernel<4><<<grid, PFT_COLS, smem, s>>>(a); break;
          case 8: pfb_fir_tma_kernel<8><<<grid, PFT_COLS, smem, s>>>(a); break;
          default: pfb_fir_tma_kernel<16><<<grid, PFT_COLS, smem, s>>>(a); break;
        }
      } else if (fast) {
        PfbFirArgs a;
        a.x = d_rows_in + r0 * (long)M; a.u = d_u.as<float2>(); a.taps_t = d_taps_t.as<float>();
        a.M = (int)M; a.T = T; a.nrows = n;
        const int gx = ((int)M + 127) / 128;
        long tiles = std::max<long>(1, (long)sm_count() * 12 / gx);
        int rpt = (int)std::max<long>(4 * TT, (n + tiles - 1) / tiles);
        rpt = ((rpt + TT - 1) / TT) * TT;
        a.rows_per_thread = rpt;
        dim3 grid(gx, (unsigned)((n + rpt - 1) / rpt));
        switch (TT) {
          case 4: pfb_fir_kernel<4><<<grid, 128, 0, s>>>(a); break;
          case 8: pfb_fir_kernel<8><<<grid, 128, 0, s>>>(a); break;
          case 16: pfb_fir_kernel<16><<<grid, 128, 0, s>>>(a); break;
          default: pfb_fir_kernel<32><<<grid, 128, 0, s>>>(a); break;
        }
      } else {
        PfbFirGenArgs a;
        a.x = d_rows_in; a.u = d_u.as<float2>(); a.taps = d_taps_plain.as<float>();
        a.M = (int)M; a.T = T; a.rr = rr; a.nout = n;
        pfb_fir_generic_kernel<<<grid_for(n * (long)M, 256), 256, 0, s>>>(a);
      }
      GRB_LAUNCH_CHECK();
      prof.end(s);
      prof.begin(1, s);
      if (demod) rc = fft_plan_exec_demod(fft, d_u.as<float2>(), demod->d, n, demod->gain, demod->atan, demod->prev_y, demod->last_y,
                                          demod->coresident, s, demod->y_out);
      else rc = fft_plan_exec(fft, d_u.as<float2>(), d_out + r0 * (long)M, n, nullptr, 0, 0, s);
      if (rc) return rc;
      prof.end(s);
    }
    return GRCUDA_OK;
  }
};

extern "C" {

grcuda_pfb* grcuda_pfb_channelizer_ccf_create(unsigned numchans, const float* taps, int ntaps, float oversample_rate) {
  // argument errors first: they are reported identically with or without a device
  if (numchans < 1 || ntaps < 0 || !(oversample_rate > 0)) { set_error(GRCUDA_EINVAL, "pfb_channelizer_ccf: bad arguments"); return nullptr; }
  double intp = 0;
  const double fltp = modf(numchans / oversample_rate, &intp);  // :56-60
  if (fltp != 0.0) {
    set_error(GRCUDA_EINVAL, "gr_pfb_channelizer: oversample rate must be N/i for i in [1, N]");
    return nullptr;
  }
  if (!device_ok()) return nullptr;
  grcuda_pfb* h = new grcuda_pfb;
  h->M = numchans;
  h->os = oversample_rate;
  h->relative_rate = 1.0 / intp;                                  // :62
  h->rr = (int)rintf(numchans / oversample_rate);                 // :81
  h->output_multiple = 1;
  while ((h->output_multiple * h->rr) % (int)numchans != 0) h->output_multiple++;  // :88-91
  h->fft = fft_plan_create((int)numchans, +1);                   // gri_fft_complex(numchans, false) (:76)
  if (!h->fft || h->base_init()) { delete h; return nullptr; }
  h->set_geometry(std::vector<float>(taps, taps + ntaps));
  if (h->upload_if_dirty()) { delete h; return nullptr; }
  h->updated = true;  // the constructor calls set_taps (:74), so the first general_work returns 0
  return h;
}
void grcuda_pfb_channelizer_ccf_destroy(grcuda_pfb* h) { delete h; }
int grcuda_pfb_channelizer_ccf_set_taps(grcuda_pfb* h, const float* taps, int ntaps) {
  std::lock_guard<std::mutex> lk(h->mu);
  h->set_geometry(std::vector<float>(taps, taps + ntaps));
  h->updated = true;
  return GRCUDA_OK;
}
unsigned grcuda_pfb_channelizer_ccf_history(grcuda_pfb* h) { return h->history; }
int grcuda_pfb_channelizer_ccf_output_multiple(grcuda_pfb* h) { return h->output_multiple; }
double grcuda_pfb_channelizer_ccf_relative_rate(grcuda_pfb* h) { return h->relative_rate; }
int grcuda_pfb_channelizer_ccf_taps_per_filter(grcuda_pfb* h) { return h->T; }

int grcuda_pfb_channelizer_ccf_set_profiling(grcuda_pfb* h, int on) { h->prof.on = on != 0; return GRCUDA_OK; }
int grcuda_pfb_channelizer_ccf_profile_read(grcuda_pfb* h, float* ms2, int* launches2) {
  float ms[EventProfiler::kStages];
  int ln[EventProfiler::kStages];
  h->prof.read(ms, ln);
  for (int i = 0; i < 2; i++) { if (ms2) ms2[i] = ms[i]; if (launches2) launches2[i] = ln[i]; }
  return GRCUDA_OK;
}

int grcuda_pfb_channelizer_ccf_work_device(grcuda_pfb* h, long nout, const grcuda_complex* d_in_rows,
                                           grcuda_complex* d_out, void* stream) {
  return h->run((const float2*)d_in_rows, (float2*)d_out, nout, h->pick(stream));
}

static int pfb_gate(grcuda_pfb* h) {
  std::lock_guard<std::mutex> lk(h->mu);
  if (h->updated) { h->updated = false; return 1; }  // :164-167
  return 0;
}

int grcuda_pfb_channelizer_ccf_work_interleaved(grcuda_pfb* h, int nout, const grcuda_complex* in, grcuda_complex* out,
                                                int* consumed) {
  if (consumed) *consumed = 0;
  if (pfb_gate(h)) return 0;
  if (nout <= 0) return 0;
  const int toconsume = (int)rintf(nout / h->os);  // :170
  const size_t rows_in = (size_t)toconsume + h->T;
  const size_t M = h->M;
  int rc;
  if ((rc = h->d_in.reserve(rows_in * M * sizeof(float2))) || (rc = h->d_out.reserve((size_t)nout * M * sizeof(float2)))) return rc;
  if ((rc = h->stager.h2d(h->d_in.p, in, rows_in * M * sizeof(float2), h->stream))) return rc;
  if ((rc = h->run(h->d_in.as<float2>(), h->d_out.as<float2>(), nout, h->stream))) return rc;
  if ((rc = h->stager.d2h(out, h->d_out.p, (size_t)nout * M * sizeof(float2), h->stream))) return rc;
  if (consumed) *consumed = toconsume;
  return nout;
}

int grcuda_pfb_channelizer_ccf_work(grcuda_pfb* h, int nout, const grcuda_complex* const* in, grcuda_complex* out,
                                    int* consumed) {
  if (consumed) *consumed = 0;
  if (pfb_gate(h)) return 0;
  if (nout <= 0) return 0;
  const int toconsume = (int)rintf(nout / h->os);
  const size_t len = (size_t)toconsume + h->T;  // items the reference may touch per stream
  const size_t M = h->M;
  int rc;
  if ((rc = h->pin_in.reserve(M * len * sizeof(float2))) || (rc = h->d_stage.reserve(M * len * sizeof(float2))) ||
      (rc = h->d_in.reserve(M * len * sizeof(float2))) || (rc = h->d_out.reserve((size_t)nout * M * sizeof(float2))))
    return rc;
  // gr_stream_to_streams in reverse: gather the M host streams (general/gr_stream_to_streams.cc:57-63)
  for (size_t j = 0; j < M; j++) memcpy((char*)h->pin_in.p + j * len * sizeof(float2), in[j], len * sizeof(float2));
  GRB_CUDA(cudaMemcpyAsync(h->d_stage.p, h->pin_in.p, M * len * sizeof(float2), cudaMemcpyHostToDevice, h->stream));
  dim3 tb(32, 8), tg((unsigned)((len + 31) / 32), (unsigned)((M + 31) / 32));
  transpose_streams_kernel<<<tg, tb, 0, h->stream>>>(h->d_stage.as<float2>(), h->d_in.as<float2>(), (int)M, (int)len);
  GRB_LAUNCH_CHECK();
  if ((rc = h->run(h->d_in.as<float2>(), h->d_out.as<float2>(), nout, h->stream))) return rc;
  if ((rc = h->stager.d2h(out, h->d_out.p, (size_t)nout * M * sizeof(float2), h->stream))) return rc;
  if (consumed) *consumed = toconsume;
  return nout;
}

}  // extern "C"

// =============================================================================================
// 8f rank 3: gr_pfb_decimator_ccf = decim branch filters + one bin of the decim-point backward DFT
// =============================================================================================
// out[i] = sum_j e^{+j 2 pi j chan / M} * sum_t taps[j + t M] * in_{M-1-j}[i + T-1-t]   (gr_pfb_decimator_ccf.cc:127-175).
// On the interleaved stream X[m M + s] = in_s[m] this is ONE decimate-by-M FIR with the composite complex taps
// h[k] = e^{+j 2 pi (k mod M) chan / M} * taps[k], k = j + t M:  out[i] = sum_k h[k] X[(i + T) M - 1 - k],
// i.e. exactly what fir_decim_kernel<complex taps> computes -- the FFT the reference "abuses" for the
// de-spinning (:151-153) is not needed for a single bin.
struct grcuda_pfb_decim : PlanBase {
  unsigned M = 1, chan = 0, T = 0;
  bool updated = true, tiled = false;
  std::vector<float> taps, new_taps;
  bool taps_pending = false;
  FirCore core;
  DevBuf d_g, d_stage;   // composite taps in window order (warp kernel); [stream][len] staging of the host form
  PinBuf pin_in;
  int build(const std::vector<float>& t) {
    cudaDeviceSynchronize();
    taps = t;
    T = (unsigned)ceil((double)t.size() / (double)M);                      // :80
    const size_t L = (size_t)M * T;
    std::vector<float> h(2 * L, 0.f);                                      // forward order, complex
    for (size_t k = 0; k < L; k++) {
      const float tv = k < t.size() ? t[k] : 0.f;
      const double ph = 2.0 * M_PI * (double)(((unsigned long long)(k % M) * chan) % M) / (double)M;
      h[2 * k] = (float)(tv * cos(ph));
      h[2 * k + 1] = (float)(tv * sin(ph));
    }
    std::vector<float> rev(2 * L);                                         // window order g[n] = h[L-1-n] = reversed taps
    for (size_t n = 0; n < L; n++) { rev[2 * n] = h[2 * (L - 1 - n)]; rev[2 * n + 1] = h[2 * (L - 1 - n) + 1]; }
    core.decim = (int)M;
    tiled = L <= 4096 && core.upload(rev.data(), (int)L, true) == GRCUDA_OK;  // else: too large for the tile
    int rc = d_g.reserve(rev.size() * sizeof(float));
    if (rc) return rc;
    GRB_CUDA(cudaMemcpy(d_g.p, rev.data(), rev.size() * sizeof(float), cudaMemcpyHostToDevice));
    return GRCUDA_OK;
  }
  bool gate() {  // "return 0 once after set_taps" (:136-139); a pending set_taps is applied here
    std::lock_guard<std::mutex> lk(mu);
    if (taps_pending) {
      taps_pending = false;
      build(new_taps);
      updated = true;
    }
    if (updated) { updated = false; return true; }
    return false;
  }
  int run(const float2* d_rows, float2* d_out, long nout, cudaStream_t s) {
    if (nout <= 0) return GRCUDA_OK;
    if (tiled) return core.launch(d_rows, d_out, nout, false, 0.0, 0, s);
    const long warps = std::min<long>(nout, (long)sm_count() * 64);
    pfb_decim_warp_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, s>>>(d_rows, d_out, nout, (int)M, d_g.as<float2>(),
                                                                           (int)(M * T));
    GRB_LAUNCH_CHECK();
    return GRCUDA_OK;
  }
};

extern "C" {

grcuda_pfb_decim* grcuda_pfb_decimator_ccf_create(unsigned decim, const float* taps, int ntaps, unsigned channel) {
  if (decim < 1 || ntaps < 1 || !taps) { set_error(GRCUDA_EINVAL, "pfb_decimator_ccf: bad decimation / taps"); return nullptr; }
  if (!device_ok()) return nullptr;
  grcuda_pfb_decim* h = new grcuda_pfb_decim;
  h->M = decim;
  h->chan = channel;
  if (h->base_init() || h->build(std::vector<float>(taps, taps + ntaps))) { delete h; return nullptr; }
  return h;
}
void grcuda_pfb_decimator_ccf_destroy(grcuda_pfb_decim* h) { delete h; }
int grcuda_pfb_decimator_ccf_set_taps(grcuda_pfb_decim* h, const float* taps, int ntaps) {
  if (ntaps < 1 || !taps) return set_error(GRCUDA_EINVAL, "pfb_decimator_ccf: no taps");
  std::lock_guard<std::mutex> lk(h->mu);
  h->new_taps.assign(taps, taps + ntaps);
  h->taps_pending = true;
  return GRCUDA_OK;
}
unsigned grcuda_pfb_decimator_ccf_history(grcuda_pfb_decim* h) {  // :107 set_history(taps_per_filter)
  std::lock_guard<std::mutex> lk(h->mu);
  return h->taps_pending ? (unsigned)ceil((double)h->new_taps.size() / (double)h->M) : h->T;
}
int grcuda_pfb_decimator_ccf_taps_per_filter(grcuda_pfb_decim* h) { return (int)grcuda_pfb_decimator_ccf_history(h); }
int grcuda_pfb_decimator_ccf_decimation(grcuda_pfb_decim* h) { return (int)h->M; }
// d_in_rows: [history-1 + noutput][decim] interleaved (row m = the m-th item of every stream); one output per row
int grcuda_pfb_decimator_ccf_work_device(grcuda_pfb_decim* h, long noutput_items, const grcuda_complex* d_in_rows,
                                         grcuda_complex* d_out, void* stream) {
  if (h->gate()) return 0;
  int rc = h->run((const float2*)d_in_rows, (float2*)d_out, noutput_items, h->pick(stream));
  return rc ? rc : (int)std::min<long>(noutput_items, 0x7fffffffL);
}
int grcuda_pfb_decimator_ccf_work_interleaved(grcuda_pfb_decim* h, int noutput_items, const grcuda_complex* in_rows,
                                              grcuda_complex* out) {
  if (h->gate()) return 0;
  if (noutput_items <= 0) return 0;
  const size_t rows = (size_t)noutput_items + h->T - 1, M = h->M;
  int rc;
  if ((rc = h->d_in.reserve(rows * M * sizeof(float2))) || (rc = h->d_out.reserve((size_t)noutput_items * sizeof(float2)))) return rc;
  if ((rc = h->stager.h2d(h->d_in.p, in_rows, rows * M * sizeof(float2), h->stream))) return rc;
  if ((rc = h->run(h->d_in.as<float2>(), h->d_out.as<float2>(), noutput_items, h->stream))) return rc;
  if ((rc = h->stager.d2h(out, h->d_out.p, (size_t)noutput_items * sizeof(float2), h->stream))) return rc;
  return noutput_items;
}
// the reference's own form: decim separate streams, each from its first history item (gr_stream_to_streams output)
int grcuda_pfb_decimator_ccf_work(grcuda_pfb_decim* h, int noutput_items, const grcuda_complex* const* in, grcuda_complex* out) {
  if (h->gate()) return 0;
  if (noutput_items <= 0) return 0;
  const size_t len = (size_t)noutput_items + h->T - 1, M = h->M;
  int rc;
  if ((rc = h->pin_in.reserve(M * len * sizeof(float2))) || (rc = h->d_stage.reserve(M * len * sizeof(float2))) ||
      (rc = h->d_in.reserve(M * len * sizeof(float2))) || (rc = h->d_out.reserve((size_t)noutput_items * sizeof(float2))))
    return rc;
  for (size_t j = 0; j < M; j++) memcpy((char*)h->pin_in.p + j * len * sizeof(float2), in[j], len * sizeof(float2));
  GRB_CUDA(cudaMemcpyAsync(h->d_stage.p, h->pin_in.p, M * len * sizeof(float2), cudaMemcpyHostToDevice, h->stream));
  dim3 tb(32, 8), tg((unsigned)((len + 31) / 32), (unsigned)((M + 31) / 32));
  transpose_streams_kernel<<<tg, tb, 0, h->stream>>>(h->d_stage.as<float2>(), h->d_in.as<float2>(), (int)M, (int)len);
  GRB_LAUNCH_CHECK();
  if ((rc = h->run(h->d_in.as<float2>(), h->d_out.as<float2>(), noutput_items, h->stream))) return rc;
  if ((rc = h->stager.d2h(out, h->d_out.p, (size_t)noutput_items * sizeof(float2), h->stream))) return rc;
  return noutput_items;
}

}  // extern "C"

// =============================================================================================
// 8f rank 4: gr_fft_filter_ccc (complex taps, history 1, output in blocks of nsamples)
// =============================================================================================
// What the block computes is y[n] = sum_k taps[k] x[n - k] (x[<0] = 0), decimated; the reference does it by
// overlap-add in the frequency domain (gri_fft_filter_ccc_generic.cc:120-165) and its own QA checks it against
// gr_fir_filter_ccc (qa_fft_filter.py).  Three device paths behind the one block contract -- history 1 (the plan
// carries the last ntaps-1 input items itself, where the reference carries its overlap-add tail), output_multiple =
// nsamples = fftsize - ntaps + 1 with fftsize = 2 * 2^ceil(log2 ntaps) (:103-118), set_taps deferred to the next
// work(), which returns 0 and clears the carried state (:62-69):
//   * ntaps <= kDirectMax: the direct-form decimating FIR (fir_decim_kernel) -- fewer operations than two FFTs;
//   * fftsize <= kFusedMax: overlap-save with both FFTs and the product inside one CTA (kernel_fft_filter.cuh);
//   * longer filters: uniformly partitioned overlap-save -- the same kernel once per 4096-tap partition, reading
//     further back in the carried input and accumulating into the output; no FFT longer than 8192 points.
struct grcuda_fft_filter : PlanBase {
  static const int kDirectMax = 32, kFusedMax = 8192, kPartTaps = 4096;
  int decim = 1, ntaps = 0, nsamples = 1, fftsize = 2;
  int path = 0;                 // 0 direct, 1 fused overlap-save, 2 partitioned fused overlap-save
  int force_path = -1;          // tests: pin the path (-1 = automatic)
  int kfft = 2, ptaps = 1, hop = 1, nparts = 1;   // kernel geometry of paths 1 / 2
  size_t carry_len = 0;         // input items kept in front of the new ones
  bool updated = false;
  std::vector<float> new_taps;  // interleaved re, im
  std::vector<float> cur_taps;
  FirCore core;
  DevBuf d_buf, d_carry;        // [carry | new input] contiguous
  DevBuf d_H, d_tw;
  FftFiltArgs fa;
  static int fftsize_for(int nt) { return (int)(2 * pow(2.0, ceil(log((double)nt) / log(2.0)))); }  // :106
  // FFT(taps[0..nt) zero padded to n) / n in double (:80-94), as float2
  static void transform_taps(const float* t_ri, int nt, int n, float2* out) {
    std::vector<std::complex<double>> H(n, 0.0);
    for (int k = 0; k < nt; k++) H[k] = std::complex<double>(t_ri[2 * k], t_ri[2 * k + 1]) / (double)n;
    for (int i = 1, j = 0; i < n; i++) {  // bit reversal
      int bit = n >> 1;
      for (; j & bit; bit >>= 1) j ^= bit;
      j ^= bit;
      if (i < j) std::swap(H[i], H[j]);
    }
    for (int len = 2; len <= n; len <<= 1) {
      for (int i = 0; i < n; i += len)
        for (int k = 0; k < len / 2; k++) {
          const double ang = -2.0 * M_PI * k / len;
          const std::complex<double> w(cos(ang), sin(ang)), u = H[i + k], v = H[i + k + len / 2] * w;
          H[i + k] = u + v;
          H[i + k + len / 2] = u - v;
        }
    }
    for (int i = 0; i < n; i++) out[i] = make_float2((float)H[i].real(), (float)H[i].imag());
  }
  int build(const std::vector<float>& t_ri) {
    cudaDeviceSynchronize();
    cur_taps = t_ri;
    ntaps = (int)t_ri.size() / 2;
    fftsize = fftsize_for(ntaps);
    nsamples = fftsize - ntaps + 1;
    path = force_path >= 0 ? force_path : (ntaps <= kDirectMax ? 0 : (fftsize <= kFusedMax ? 1 : 2));
    if (path == 1 && fftsize > kFusedMax) path = 2;
    if (path == 1) { kfft = fftsize; ptaps = ntaps; nparts = 1; }
    else if (path == 2) { ptaps = std::min(ntaps, (int)kPartTaps); kfft = fftsize_for(ptaps); nparts = (ntaps + ptaps - 1) / ptaps; }
    hop = kfft - ptaps + 1;
    carry_len = path == 0 ? (size_t)(ntaps - 1) : (size_t)nparts * ptaps - 1;
    int rc;
    const size_t cb = std::max<size_t>(carry_len, 1) * sizeof(float2);
    if ((rc = d_carry.reserve(cb))) return rc;
    GRB_CUDA(cudaMemset(d_carry.p, 0, cb));  // the tail is cleared by set_taps (:67-69)
    if (path == 0) {
      std::vector<float> rev(t_ri.size());
      for (int k = 0; k < ntaps; k++) { rev[2 * k] = t_ri[2 * (ntaps - 1 - k)]; rev[2 * k + 1] = t_ri[2 * (ntaps - 1 - k) + 1]; }
      core.decim = decim;
      return core.upload(rev.data(), ntaps, true);
    }
    const int n = kfft;
    std::vector<float2> Hf((size_t)n * nparts);
    for (int p = 0; p < nparts; p++)
      transform_taps(t_ri.data() + 2 * (size_t)p * ptaps, std::min(ptaps, ntaps - p * ptaps), n, Hf.data() + (size_t)p * n);
    if ((rc = d_H.reserve(Hf.size() * sizeof(float2)))) return rc;
    GRB_CUDA(cudaMemcpy(d_H.p, Hf.data(), Hf.size() * sizeof(float2), cudaMemcpyHostToDevice));
    memset(&fa, 0, sizeof fa);
    int m = 0;
    while ((1 << m) < n) m++;
    std::vector<float2> tw;
    int Ns = 1;
    fa.npass = 0;
    while (m > 0) {
      const int lr = m >= 4 ? 4 : m, R = 1 << lr;
      fa.radix[fa.npass] = R;
      fa.tw_off[fa.npass] = (int)tw.size();
      for (int k = 0; k < Ns; k++) {
        const double ang = -2.0 * M_PI * k / ((double)Ns * R);
        tw.push_back(make_float2((float)cos(ang), (float)sin(ang)));
      }
      Ns *= R;
      m -= lr;
      fa.npass++;
    }
    if ((rc = d_tw.reserve(tw.size() * sizeof(float2)))) return rc;
    GRB_CUDA(cudaMemcpy(d_tw.p, tw.data(), tw.size() * sizeof(float2), cudaMemcpyHostToDevice));
    fa.tw = d_tw.as<float2>();
    fa.n = n; fa.ntaps = ptaps; fa.nsamples = hop; fa.decim = decim;
    const size_t smem = (size_t)(n + n / 16 + 1) * sizeof(float2);
    if (cudaError_t e = raise_dynamic_smem((const void*)fft_filter_ols_kernel<512>, smem))
      return set_error(GRCUDA_ECUDA, "fft_filter_ccc: %zu B of shared memory: %s", smem, cudaGetErrorString(e));
    return GRCUDA_OK;
  }
  // buf = [carry_len carried | nin new] -> out (nin / decim items)
  int run_freq(const float2* buf, float2* out, long nin, cudaStream_t s) {
    const int threads = std::max(32, kfft / FFTF_ELEMS);
    const size_t smem = (size_t)(kfft + kfft / 16 + 1) * sizeof(float2);
    const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(32, (200u << 10) / std::max<size_t>(smem, 1)));
    const long nblk = (nin + hop - 1) / hop;
    const int grid = (int)std::min<long>(nblk, (long)sm_count() * per_sm);
    for (int p = 0; p < nparts; p++) {
      FftFiltArgs a = fa;
      const long start = (long)carry_len - (ptaps - 1) - (long)p * ptaps;   // >= 0 by the choice of carry_len
      a.x = buf + start;
      a.x_limit = (long)carry_len + nin - start;
      a.out = out; a.nblk = nblk; a.total_items = nin; a.accumulate = p > 0;
      a.H = d_H.as<float2>() + (size_t)p * kfft;
      fft_filter_ols_kernel<512><<<grid, threads, smem, s>>>(a);
      GRB_LAUNCH_CHECK();
    }
    return GRCUDA_OK;
  }
};

extern "C" {

grcuda_fft_filter* grcuda_fft_filter_ccc_create(int decimation, const grcuda_complex* taps, int ntaps) {
  if (decimation < 1 || ntaps < 1 || !taps) { set_error(GRCUDA_EINVAL, "fft_filter_ccc: bad decimation / taps"); return nullptr; }
  if (!device_ok()) return nullptr;
  grcuda_fft_filter* h = new grcuda_fft_filter;
  h->decim = decimation;
  if (h->base_init() || h->build(std::vector<float>((const float*)taps, (const float*)taps + 2 * (size_t)ntaps))) { delete h; return nullptr; }
  return h;
}
void grcuda_fft_filter_ccc_destroy(grcuda_fft_filter* h) { delete h; }
int grcuda_fft_filter_ccc_set_taps(grcuda_fft_filter* h, const grcuda_complex* taps, int ntaps) {
  if (ntaps < 1 || !taps) return set_error(GRCUDA_EINVAL, "fft_filter_ccc: no taps");
  std::lock_guard<std::mutex> lk(h->mu);  // gr_fft_filter_ccc.cc:75-79: deferred to the next work()
  h->new_taps.assign((const float*)taps, (const float*)taps + 2 * (size_t)ntaps);
  h->updated = true;
  return GRCUDA_OK;
}
int grcuda_fft_filter_ccc_output_multiple(grcuda_fft_filter* h) { return h->nsamples; }  // :66
int grcuda_fft_filter_ccc_path(grcuda_fft_filter* h) { return h->path; }
int grcuda_fft_filter_ccc_set_path(grcuda_fft_filter* h, int path) {
  if (path < -1 || path > 2) return set_error(GRCUDA_EINVAL, "fft_filter_ccc: unknown path %d", path);
  std::lock_guard<std::mutex> lk(h->mu);
  h->force_path = path;
  if (!h->updated) { h->new_taps = h->cur_taps; h->updated = true; }   // rebuilt at the next work(), like set_taps
  return GRCUDA_OK;
}
int grcuda_fft_filter_ccc_decimation(grcuda_fft_filter* h) { return h->decim; }
unsigned grcuda_fft_filter_ccc_history(grcuda_fft_filter* h) { (void)h; return 1; }        // :58
// d_in: noutput_items * decimation NEW items (history 1); noutput_items must be a multiple of output_multiple()
int grcuda_fft_filter_ccc_work_device(grcuda_fft_filter* h, int noutput_items, const grcuda_complex* d_in,
                                      grcuda_complex* d_out, void* stream) {
  {
    std::lock_guard<std::mutex> lk(h->mu);
    if (h->updated) {  // :85-90: "output multiple may have changed"
      h->updated = false;
      int rc = h->build(h->new_taps);
      return rc ? rc : 0;
    }
  }
  if (noutput_items <= 0) return 0;
  if (noutput_items % h->nsamples) return set_error(GRCUDA_EINVAL, "fft_filter_ccc: noutput_items %d is not a multiple of %d (:92)", noutput_items, h->nsamples);
  cudaStream_t s = h->pick(stream);
  const size_t nin = (size_t)noutput_items * h->decim, nc = h->carry_len;
  int rc;
  if ((rc = h->d_buf.reserve((nc + nin) * sizeof(float2)))) return rc;
  float2* buf = h->d_buf.as<float2>();
  if (nc) GRB_CUDA(cudaMemcpyAsync(buf, h->d_carry.p, nc * sizeof(float2), cudaMemcpyDeviceToDevice, s));
  GRB_CUDA(cudaMemcpyAsync(buf + nc, d_in, nin * sizeof(float2), cudaMemcpyDeviceToDevice, s));
  if (h->path == 0) rc = h->core.launch(buf, (float2*)d_out, noutput_items, false, 0.0, 0, s);
  else rc = h->run_freq(buf, (float2*)d_out, (long)nin, s);
  if (rc) return rc;
  if (nc) GRB_CUDA(cudaMemcpyAsync(h->d_carry.p, buf + nin, nc * sizeof(float2), cudaMemcpyDeviceToDevice, s));
  return noutput_items;
}
int grcuda_fft_filter_ccc_work(grcuda_fft_filter* h, int noutput_items, const grcuda_complex* in, grcuda_complex* out) {
  {
    std::lock_guard<std::mutex> lk(h->mu);
    if (h->updated) {
      h->updated = false;
      int rc = h->build(h->new_taps);
      return rc ? rc : 0;
    }
  }
  if (noutput_items <= 0) return 0;
  const size_t nin = (size_t)noutput_items * h->decim;
  int rc;
  if ((rc = h->d_in.reserve(nin * sizeof(float2))) || (rc = h->d_out.reserve((size_t)noutput_items * sizeof(float2)))) return rc;
  if ((rc = h->stager.h2d(h->d_in.p, in, nin * sizeof(float2), h->stream))) return rc;
  const int r = grcuda_fft_filter_ccc_work_device(h, noutput_items, (const grcuda_complex*)h->d_in.p, (grcuda_complex*)h->d_out.p, h->stream);
  if (r <= 0) return r;
  if ((rc = h->stager.d2h(out, h->d_out.p, (size_t)r * sizeof(float2), h->stream))) return rc;
  return r;
}

}  // extern "C"

// =============================================================================================
// a5 / a6: gr_fft_vcc
// =============================================================================================
struct grcuda_fft : PlanBase {
  int n = 0;
  bool forward = true, shift = false;
  FftPlan* plan = nullptr;
  DevBuf d_window;
  int nwindow = 0;
  ~grcuda_fft() { fft_plan_destroy(plan); }
  int set_window(const float* w, int nw) {  // gr_fft_vcc.cc:55-64
    if (!(nw == 0 || nw == n)) return 0;
    std::lock_guard<std::mutex> lk(mu);
    if (nw) {
      if (d_window.reserve((size_t)nw * sizeof(float))) return GRCUDA_ENOMEM;
      if (cudaMemcpy(d_window.p, w, (size_t)nw * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) return GRCUDA_ECUDA;
    }
    nwindow = nw;
    return 1;
  }
  int run(const float2* d_in, float2* d_out, long nvec, cudaStream_t s) {
    int in_rot = 0, out_rot = 0;
    const float* win = nullptr;
    {
      std::lock_guard<std::mutex> lk(mu);
      if (nwindow) win = d_window.as<float>();                        // gr_fft_vcc_fftw.cc:68-72
      else if (!forward && shift) in_rot = (int)floor(n / 2.0);       // :74-79
      if (forward && shift) out_rot = n - (int)ceil(n / 2.0);         // :89-93
    }
    return fft_plan_exec(plan, d_in, d_out, nvec, win, in_rot % n, out_rot % n, s);
  }
};

extern "C" {

grcuda_fft* grcuda_fft_vcc_create(int fft_size, int forward, const float* window, int nwindow, int shift) {
  if (fft_size <= 0) { set_error(GRCUDA_ERANGE, "gri_fftw: invalid fft_size"); return nullptr; }
  if (!device_ok()) return nullptr;
  grcuda_fft* h = new grcuda_fft;
  h->n = fft_size;
  h->forward = forward != 0;
  h->shift = shift != 0;
  h->plan = fft_plan_create(fft_size, forward ? -1 : +1);
  if (!h->plan || h->base_init()) { delete h; return nullptr; }
  h->set_window(window, nwindow);  // a wrong-sized window is ignored, like the reference constructor
  return h;
}
void grcuda_fft_vcc_destroy(grcuda_fft* h) { delete h; }
int grcuda_fft_vcc_set_window(grcuda_fft* h, const float* window, int nwindow) { return h->set_window(window, nwindow); }
int grcuda_fft_vcc_work_device(grcuda_fft* h, long nvec, const grcuda_complex* d_in, grcuda_complex* d_out, void* stream) {
  return h->run((const float2*)d_in, (float2*)d_out, nvec, h->pick(stream));
}
int grcuda_fft_vcc_work(grcuda_fft* h, int nvec, const grcuda_complex* in, grcuda_complex* out) {
  if (nvec <= 0) return 0;
  const size_t bytes = (size_t)nvec * h->n * sizeof(float2);
  int rc;
  if ((rc = h->d_in.reserve(bytes)) || (rc = h->d_out.reserve(bytes))) return rc;
  if ((rc = h->stager.h2d(h->d_in.p, in, bytes, h->stream))) return rc;
  if ((rc = h->run(h->d_in.as<float2>(), h->d_out.as<float2>(), nvec, h->stream))) return rc;
  if ((rc = h->stager.d2h(out, h->d_out.p, bytes, h->stream))) return rc;
  return nvec;
}

}  // extern "C"

// =============================================================================================
// a7 / a8: gr_quadrature_demod_cf
// =============================================================================================
struct grcuda_quad : PlanBase {
  float gain = 1.f;
  DeviceTables tabs;
  int launch(const float2* d_in, float* d_out, long nrows, int nchan, cudaStream_t s) {
    if (nrows <= 0 || nchan <= 0) return GRCUDA_OK;
    float g;
    { std::lock_guard<std::mutex> lk(mu); g = gain; }
    quad_demod_kernel<<<grid_for(nrows * (long)nchan, 256, 16), 256, 0, s>>>(d_in, d_out, nrows, nchan, g, tabs.atan);
    GRB_LAUNCH_CHECK();
    return GRCUDA_OK;
  }
};

extern "C" {

grcuda_quad* grcuda_quadrature_demod_cf_create(float gain) {
  if (!device_ok()) return nullptr;
  grcuda_quad* h = new grcuda_quad;
  h->gain = gain;
  if (h->base_init() || get_tables(&h->tabs)) { delete h; return nullptr; }
  return h;
}
void grcuda_quadrature_demod_cf_destroy(grcuda_quad* h) { delete h; }
int grcuda_quadrature_demod_cf_set_gain(grcuda_quad* h, float gain) {
  std::lock_guard<std::mutex> lk(h->mu);
  h->gain = gain;
  return GRCUDA_OK;
}
float grcuda_quadrature_demod_cf_gain(grcuda_quad* h) { return h->gain; }
int grcuda_quadrature_demod_cf_work_device(grcuda_quad* h, long nrows, int nchan, const grcuda_complex* d_in,
                                           float* d_out, void* stream) {
  return h->launch((const float2*)d_in, d_out, nrows, nchan, h->pick(stream));
}
int grcuda_quadrature_demod_cf_work(grcuda_quad* h, int nout, const grcuda_complex* in, float* out) {
  if (nout <= 0) return 0;
  int rc;
  if ((rc = h->d_in.reserve((size_t)(nout + 1) * sizeof(float2))) || (rc = h->d_out.reserve((size_t)nout * sizeof(float)))) return rc;
  if ((rc = h->stager.h2d(h->d_in.p, in, (size_t)(nout + 1) * sizeof(float2), h->stream))) return rc;
  if ((rc = h->launch(h->d_in.as<float2>(), h->d_out.as<float>(), nout, 1, h->stream))) return rc;
  if ((rc = h->stager.d2h(out, h->d_out.p, (size_t)nout * sizeof(float), h->stream))) return rc;
  return nout;
}
int grcuda_fast_atan2f_device(const float* d_y, const float* d_x, float* d_out, long n, void* stream) {
  if (!device_ok()) return GRCUDA_ECUDA;
  DeviceTables t;
  int rc = get_tables(&t);
  if (rc) return rc;
  if (n <= 0) return GRCUDA_OK;
  fast_atan2f_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(d_y, d_x, d_out, n, t.atan);
  GRB_LAUNCH_CHECK();
  return GRCUDA_OK;
}

}  // extern "C"

// =============================================================================================
// a9 / a10 / a11: digital_clock_recovery_mm_ff (+ slicers)
// =============================================================================================
struct grcuda_mm : PlanBase {
  int nchan = 1, order = GRCUDA_ORDER_SSE;
  float omega0 = 2, gain_omega = 0, gain_mu = 0, mu0 = 0, limit = 0.001f;
  float min_omega = 0, max_omega = 0, omega_mid = 0;
  int slicer_levels = 0;
  float slicer_alpha = 0, slicer_beta = 1;
  int variant = -1;  // which build of the kernel runs (grcuda_clock_recovery_mm_ff_set_kernel_variant); -1 = automatic
  DeviceTables tabs;
  DevBuf d_state, d_counts, d_slice;
  void calc_omega(float omega) {  // set_omega (digital_clock_recovery_mm_ff.h:75-80)
    min_omega = (float)(omega * (1.0 - limit));
    max_omega = (float)(omega * (1.0 + limit));
    omega_mid = (float)(0.5 * (min_omega + max_omega));
  }
  int reset_state(bool only_mu, bool only_omega, float v) {
    std::vector<MMChanState> st(nchan);
    // every plan / chain stream is non-blocking: a read-modify-write of the loop state through the legacy stream is
    // only ordered against a running clock-recovery kernel by a device-wide synchronisation
    GRB_CUDA(cudaDeviceSynchronize());
    if (only_mu || only_omega) {
      GRB_CUDA(cudaMemcpy(st.data(), d_state.p, st.size() * sizeof(MMChanState), cudaMemcpyDeviceToHost));
      for (auto& s : st) { if (only_mu) s.mu = v; else s.omega = v; }
    } else {
      for (auto& s : st) { s.mu = mu0; s.omega = omega0; s.last_sample = 0.f; s.slicer_avg = 0.f; s.next_abs = 0; s.clamped = 0; s.overflow = 0; }
    }
    GRB_CUDA(cudaMemcpy(d_state.p, st.data(), st.size() * sizeof(MMChanState), cudaMemcpyHostToDevice));
    return GRCUDA_OK;
  }
  int launch(const float* d_in, long ninput, long abs_row0, float* d_out, unsigned char* d_sl, int max_out, int* d_cnt,
             cudaStream_t s, const MMCorrFuse* fuse = nullptr, const void* d_state_in = nullptr, void* d_state_out2 = nullptr) {
    MMArgs a;
    memset(&a.corr, 0, sizeof a.corr);
    if (fuse) a.corr = *fuse;
    a.in = d_in; a.ninput = ninput; a.abs_row0 = abs_row0; a.nchan = nchan; a.out = d_out; a.sliced = d_sl;
    a.max_out = max_out; a.counts = d_cnt; a.state = d_state.as<MMChanState>();
    a.state_in = (const MMChanState*)d_state_in; a.state_out2 = (MMChanState*)d_state_out2;
    {
      std::lock_guard<std::mutex> lk(mu);
      a.p.gain_omega = gain_omega; a.p.gain_mu = gain_mu; a.p.omega_mid = omega_mid; a.p.omega_relative_limit = limit;
      a.slicer_levels = slicer_levels; a.slicer_alpha = slicer_alpha; a.slicer_beta = slicer_beta;
    }
    a.order = order; a.mmse_eff = tabs.mmse_eff; a.one = 1.0f;
    // look-ahead ring depth in rows: ~120 rows is >= 24 symbols up to 5 samples/symbol (several HBM round
    // trips at the loop's pace); slower symbol rates (the 10 samples/symbol single-channel config) go deeper
    const int grid = (nchan + MMW_CH - 1) / MMW_CH;
    typedef void (*mm_kernel_t)(const MMArgs);
    const bool deep = max_omega > 5.0f;
    const bool sse = order == GRCUDA_ORDER_SSE;
    mm_kernel_t k = nullptr;
    int tabrep = 1;
    const int ringrows = deep ? 512 : 128;
#define MMK(RINGV, NREG, TRV, COREV, LDV) \
  (tabrep = TRV, sse ? mm_ws_kernel<RINGV, GR_ORDER_SSE, NREG, TRV, COREV, LDV> : mm_ws_kernel<RINGV, GR_ORDER_GENERIC, NREG, TRV, COREV, LDV>)
    // the bulk-copy loader moves 16-byte multiples from 16-byte aligned addresses
    const bool tma_ok = nchan % 4 == 0 && ((uintptr_t)d_in & 15) == 0;
    int v = variant;
    // automatic: the quad-ring kernel when its 16-byte staging copies apply, else the per-lane loader with the same core
    if (v < 0) v = tma_ok ? 21 : 11;
    if (v >= 16 && !tma_ok) v = 11;
    size_t smem = 0;
    if (v >= 20 && !deep) {  // quad ring + TMA staging (kernel_mm_quad.cuh)
      k = v == 20 ? (sse ? mm_quad_kernel<GR_ORDER_SSE, 80> : mm_quad_kernel<GR_ORDER_GENERIC, 80>)
        : v == 21 ? (sse ? mm_quad_kernel<GR_ORDER_SSE, 64> : mm_quad_kernel<GR_ORDER_GENERIC, 64>)
                  : (sse ? mm_quad_kernel<GR_ORDER_SSE, 96> : mm_quad_kernel<GR_ORDER_GENERIC, 96>);
      smem = mm_quad_smem_bytes();
    } else
    if (deep) k = v == 0 ? MMK(512, 48, 1, 1, 0) : MMK(512, 64, 1, 3, 0);
    else switch (v) {
      case 0: k = MMK(128, 48, 1, 1, 0); break;   // round-1 kernel, co-resident with the front kernels (47 KB)
      case 1: k = MMK(128, 64, 1, 1, 0); break;
      case 2: k = MMK(128, 64, 8, 1, 0); break;
      case 3: k = MMK(128, 48, 1, 2, 0); break;
      case 4: k = MMK(128, 64, 1, 2, 0); break;
      case 5: k = MMK(128, 64, 8, 2, 0); break;
      case 6: k = MMK(128, 80, 8, 2, 0); break;
      case 7: k = MMK(128, 96, 8, 2, 0); break;
      case 8: k = MMK(128, 128, 8, 2, 0); break;
      case 9: k = MMK(128, 80, 1, 2, 0); break;
      case 10: k = MMK(128, 48, 1, 3, 0); break;
      case 11: k = MMK(128, 64, 1, 3, 0); break;
      case 12: k = MMK(128, 64, 8, 3, 0); break;
      case 13: k = MMK(128, 80, 8, 3, 0); break;
      case 14: k = MMK(128, 96, 8, 3, 0); break;
      case 15: k = MMK(128, 128, 8, 3, 0); break;
      case 16: k = MMK(128, 80, 8, 3, 1); break;
      case 17: k = MMK(128, 64, 8, 3, 1); break;
      case 18: k = MMK(128, 64, 1, 3, 1); break;
      case 19: k = MMK(128, 48, 1, 3, 1); break;
      default: return set_error(GRCUDA_EINVAL, "clock_recovery_mm_ff: unknown kernel variant %d", variant);
    }
#undef MMK
    if (!smem) smem = mm_ws_smem_bytes(ringrows, tabrep);
    GRB_CUDA(raise_dynamic_smem((const void*)k, (size_t)smem));  // per device
    k<<<grid, MMW_THREADS, smem, s>>>(a);
    GRB_LAUNCH_CHECK();
    return GRCUDA_OK;
  }
};

extern "C" {

grcuda_mm* grcuda_clock_recovery_mm_ff_create(int nchan, float omega, float gain_omega, float mu, float gain_mu,
                                             float omega_relative_limit, int order) {
  if (omega < 1) { set_error(GRCUDA_ERANGE, "clock rate must be > 0"); return nullptr; }                      // :58-59
  if (gain_mu < 0 || gain_omega < 0) { set_error(GRCUDA_ERANGE, "Gains must be non-negative"); return nullptr; }  // :60-61
  if (nchan < 1) { set_error(GRCUDA_EINVAL, "clock_recovery_mm_ff: nchan < 1"); return nullptr; }
  if (!device_ok()) return nullptr;
  grcuda_mm* h = new grcuda_mm;
  h->nchan = nchan; h->order = order; h->omega0 = omega; h->gain_omega = gain_omega; h->mu0 = mu; h->gain_mu = gain_mu;
  h->limit = omega_relative_limit;
  h->calc_omega(omega);
  if (h->base_init() || get_tables(&h->tabs) || h->d_state.reserve((size_t)nchan * sizeof(MMChanState)) ||
      h->d_counts.reserve((size_t)nchan * sizeof(int)) || h->reset_state(false, false, 0.f)) {
    delete h;
    return nullptr;
  }
  return h;
}
void grcuda_clock_recovery_mm_ff_destroy(grcuda_mm* h) { delete h; }
int grcuda_clock_recovery_mm_ff_forecast(grcuda_mm* h, int noutput_items) {  // :80-87 (uses the nominal omega)
  float omega = h->omega0;
  if (h->nchan == 1) {
    MMChanState st;
    if (cudaMemcpy(&st, h->d_state.p, sizeof st, cudaMemcpyDeviceToHost) == cudaSuccess) omega = st.omega;
  }
  return (int)ceil((noutput_items * omega) + 8);
}
int grcuda_clock_recovery_mm_ff_get_state(grcuda_mm* h, int chan, float* mu, float* omega, float* last_sample) {
  if (chan < 0 || chan >= h->nchan) return set_error(GRCUDA_EINVAL, "channel out of range");
  MMChanState st;
  GRB_CUDA(cudaMemcpy(&st, h->d_state.as<MMChanState>() + chan, sizeof st, cudaMemcpyDeviceToHost));
  if (mu) *mu = st.mu;
  if (omega) *omega = st.omega;
  if (last_sample) *last_sample = st.last_sample;
  return GRCUDA_OK;
}
int grcuda_clock_recovery_mm_ff_set_mu(grcuda_mm* h, float mu) { return h->reset_state(true, false, mu); }
int grcuda_clock_recovery_mm_ff_set_omega(grcuda_mm* h, float omega) {
  { std::lock_guard<std::mutex> lk(h->mu); h->calc_omega(omega); }
  return h->reset_state(false, true, omega);
}
int grcuda_clock_recovery_mm_ff_set_gain_mu(grcuda_mm* h, float g) { std::lock_guard<std::mutex> lk(h->mu); h->gain_mu = g; return GRCUDA_OK; }
int grcuda_clock_recovery_mm_ff_set_gain_omega(grcuda_mm* h, float g) { std::lock_guard<std::mutex> lk(h->mu); h->gain_omega = g; return GRCUDA_OK; }
int grcuda_clock_recovery_mm_ff_set_slicer(grcuda_mm* h, int levels, float alpha) {
  if (levels != 0 && levels != 2 && levels != 4) return set_error(GRCUDA_EINVAL, "slicer levels must be 0, 2 or 4");
  std::lock_guard<std::mutex> lk(h->mu);
  h->slicer_levels = levels;
  h->slicer_alpha = alpha;
  h->slicer_beta = (float)(1.0 - alpha);  // pager_slicer_fb.cc:40
  return GRCUDA_OK;
}
#ifdef MMW_STATS
// lab build only (not declared in gr_cuda.h): reads and clears the core-warp counters of kernel_mm.cuh
__attribute__((visibility("default"))) int grcuda_lab_mm_stats(unsigned long long* out12) {
  GRB_CUDA(cudaDeviceSynchronize());
  GRB_CUDA(cudaMemcpyFromSymbol(out12, grb::mmw_stats, 12 * sizeof(unsigned long long)));
  unsigned long long z[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  GRB_CUDA(cudaMemcpyToSymbol(grb::mmw_stats, z, sizeof z));
  return GRCUDA_OK;
}
#endif
int grcuda_clock_recovery_mm_ff_counters(grcuda_mm* h, long long* clamped, long long* overflow) {
  return mm_counters(h, clamped, overflow);
}
int grcuda_clock_recovery_mm_ff_set_kernel_variant(grcuda_mm* h, int variant) {
  if (variant < -1 || variant >= GRCUDA_MM_VARIANTS) return set_error(GRCUDA_EINVAL, "clock_recovery_mm_ff: unknown kernel variant %d", variant);
  std::lock_guard<std::mutex> lk(h->mu);
  h->variant = variant;
  return GRCUDA_OK;
}
int grcuda_clock_recovery_mm_ff_work_device(grcuda_mm* h, long ninput_rows, long abs_row0, const float* d_in,
                                            float* d_out, unsigned char* d_slice_out, int max_out, int* d_counts,
                                            void* stream) {
  return h->launch(d_in, ninput_rows, abs_row0, d_out, d_slice_out, max_out, d_counts, h->pick(stream));
}
int grcuda_clock_recovery_mm_ff_work(grcuda_mm* h, int noutput_items, int ninput_items, const float* in, float* out,
                                     int* consumed, long abs_index0) {
  if (consumed) *consumed = 0;
  if (h->nchan != 1) return set_error(GRCUDA_EINVAL, "host work() is the single-stream form (nchan == 1)");
  if (noutput_items <= 0 || ninput_items <= 0) return 0;
  int rc;
  if ((rc = h->d_in.reserve((size_t)ninput_items * sizeof(float))) || (rc = h->d_out.reserve((size_t)noutput_items * sizeof(float)))) return rc;
  if ((rc = h->stager.h2d(h->d_in.p, in, (size_t)ninput_items * sizeof(float), h->stream))) return rc;
  // the runtime re-presents unconsumed items at in[0]: the carried position is abs_index0
  GRB_CUDA(cudaMemcpyAsync((char*)h->d_state.p + offsetof(MMChanState, next_abs), &abs_index0, sizeof(long long),
                           cudaMemcpyHostToDevice, h->stream));
  if ((rc = h->launch(h->d_in.as<float>(), ninput_items, abs_index0, h->d_out.as<float>(), nullptr, noutput_items,
                      h->d_counts.as<int>(), h->stream)))
    return rc;
  int produced = 0;
  MMChanState st;
  GRB_CUDA(cudaMemcpyAsync(&produced, h->d_counts.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  GRB_CUDA(cudaMemcpyAsync(&st, h->d_state.p, sizeof st, cudaMemcpyDeviceToHost, h->stream));
  GRB_CUDA(cudaStreamSynchronize(h->stream));
  if ((rc = h->stager.d2h(out, h->d_out.p, (size_t)produced * sizeof(float), h->stream))) return rc;
  if (consumed) *consumed = (int)(st.next_abs - abs_index0);
  return produced;
}

}  // extern "C"

struct grcuda_slicer : PlanBase {
  int levels = 4;
  float alpha = 0, beta = 1;
  DevBuf d_avg;
};

extern "C" {

static grcuda_slicer* slicer_new(int levels, float alpha) {
  if (!device_ok()) return nullptr;
  grcuda_slicer* h = new grcuda_slicer;
  h->levels = levels;
  h->alpha = alpha;
  h->beta = (float)(1.0 - alpha);
  if (h->base_init() || h->d_avg.reserve(sizeof(float)) || cudaMemset(h->d_avg.p, 0, sizeof(float)) != cudaSuccess) {
    delete h;
    return nullptr;
  }
  return h;
}
grcuda_slicer* grcuda_pager_slicer_fb_create(float alpha) { return slicer_new(4, alpha); }
grcuda_slicer* grcuda_binary_slicer_fb_create(void) { return slicer_new(2, 0.f); }
void grcuda_slicer_destroy(grcuda_slicer* h) { delete h; }
float grcuda_pager_slicer_fb_dc_offset(grcuda_slicer* h) {
  float v = 0.f;
  cudaMemcpy(&v, h->d_avg.p, sizeof v, cudaMemcpyDeviceToHost);
  return v;
}
int grcuda_slicer_work(grcuda_slicer* h, int n, const float* in, unsigned char* out) {
  if (n <= 0) return 0;
  int rc;
  if ((rc = h->d_in.reserve((size_t)n * sizeof(float))) || (rc = h->d_out.reserve((size_t)n))) return rc;
  if ((rc = h->stager.h2d(h->d_in.p, in, (size_t)n * sizeof(float), h->stream))) return rc;
  slicer_kernel<<<grid_for(n, 256), 256, 0, h->stream>>>(h->d_in.as<float>(), h->d_out.as<unsigned char>(), n, h->levels,
                                                         h->alpha, h->beta, h->d_avg.as<float>());
  GRB_LAUNCH_CHECK();
  if ((rc = h->stager.d2h(out, h->d_out.p, (size_t)n, h->stream))) return rc;
  return n;
}

}  // extern "C"

// =============================================================================================
// a12 / a13: map_bb + unpack_k_bits_bb + digital_correlate_access_code_bb
// =============================================================================================
struct grcuda_corr : PlanBase {
  int nchan = 1;
  CorrParams p;
  DevBuf d_state, d_hits, d_nhits;
  int set_code(const char* code) {  // set_access_code (:64-85)
    const size_t len = strlen(code);
    if (len > 64) return set_error(GRCUDA_ERANGE, "access_code is > 64 bits");
    std::lock_guard<std::mutex> lk(mu);
    p.mask = len ? ((~0ULL) >> (64 - len)) << (64 - len) : 0ULL;
    p.flag_bit = len ? 1ULL << (64 - len) : 0ULL;
    p.access_code = 0;
    for (unsigned i = 0; i < 64; i++) {
      p.access_code <<= 1;
      if (i < len) p.access_code |= (unsigned long long)(code[i] & 1);
    }
    return GRCUDA_OK;
  }
  int launch(const unsigned char* d_sym, const int* d_counts, int fixed_count, const int* map, int nmap, int k,
             unsigned char* d_out, CorrHit* d_hits_, int max_hits, int* d_nhits_, cudaStream_t s) {
    CorrArgs a;
    a.symbols = d_sym; a.counts = d_counts; a.fixed_count = fixed_count; a.nchan = nchan;
    for (int i = 0; i < 256; i++) a.map[i] = (unsigned char)i;  // gr_map_bb.cc:40-46
    for (int i = 0; i < std::min(nmap, 256); i++) a.map[i] = (unsigned char)map[i];
    a.bits_per_symbol = k; a.out = d_out; a.state = d_state.as<CorrChanState>();
    { std::lock_guard<std::mutex> lk(mu); a.p = p; }
    a.hits = d_hits_; a.max_hits = max_hits; a.nhits = d_nhits_;
    const int threads = 64;
    corr_kernel<<<(nchan + threads - 1) / threads, threads, 0, s>>>(a);
    GRB_LAUNCH_CHECK();
    return GRCUDA_OK;
  }
  // time-parallel form (kernels_demod.cuh: corr_par_kernel): 2 bits per symbol, code length >= 16, no byte output
  bool par_ok(int k, const unsigned char* d_out) {
    std::lock_guard<std::mutex> lk(mu);
    const int len = p.flag_bit ? 64 - (__builtin_ffsll((long long)p.flag_bit) - 1) : 0;
    return k == 2 && d_out == nullptr && len >= 16;
  }
  int launch_par(const unsigned char* d_sym, const int* d_counts, int sym_rows, const int* map, int nmap, CorrHit* d_hits_,
                 int max_hits, int* d_nhits_, cudaStream_t s, const void* ext_in = nullptr, void* ext_out = nullptr) {
    int rc;
    if ((rc = d_state_next.reserve((size_t)nchan * sizeof(CorrChanState)))) return rc;
    if (ext_in) GRB_CUDA(cudaMemcpyAsync(d_state.p, ext_in, (size_t)nchan * sizeof(CorrChanState), cudaMemcpyDeviceToDevice, s));
    CorrParArgs a;
    a.symbols = d_sym; a.counts = d_counts; a.nchan = nchan;
    for (int i = 0; i < 256; i++) a.map[i] = (unsigned char)i;
    for (int i = 0; i < std::min(nmap, 256); i++) a.map[i] = (unsigned char)map[i];
    a.state_in = d_state.as<CorrChanState>(); a.state_out = d_state_next.as<CorrChanState>();
    { std::lock_guard<std::mutex> lk(mu); a.p = p; }
    a.hits = d_hits_; a.max_hits = max_hits; a.nhits = d_nhits_;
    const dim3 grid((nchan + 127) / 128, std::max(1, (sym_rows + CORR_CHUNK - 1) / CORR_CHUNK));
    corr_par_kernel<<<grid, 128, 0, s>>>(a);
    GRB_LAUNCH_CHECK();
    GRB_CUDA(cudaMemcpyAsync(d_state.p, d_state_next.p, (size_t)nchan * sizeof(CorrChanState), cudaMemcpyDeviceToDevice, s));
    if (ext_out) GRB_CUDA(cudaMemcpyAsync(ext_out, d_state_next.p, (size_t)nchan * sizeof(CorrChanState), cudaMemcpyDeviceToDevice, s));
    return GRCUDA_OK;
  }
  DevBuf d_state_next;
};

extern "C" {

grcuda_corr* grcuda_correlate_access_code_bb_create(int nchan, const char* access_code, int threshold) {
  if (nchan < 1 || !access_code) { set_error(GRCUDA_EINVAL, "correlate_access_code_bb: bad arguments"); return nullptr; }
  if (strlen(access_code) > 64) { set_error(GRCUDA_ERANGE, "access_code is > 64 bits"); return nullptr; }  // :54-57
  if (!device_ok()) return nullptr;
  grcuda_corr* h = new grcuda_corr;
  h->nchan = nchan;
  h->p.threshold = (unsigned)threshold;
  if (h->base_init() || h->set_code(access_code) || h->d_state.reserve((size_t)nchan * sizeof(CorrChanState)) ||
      h->d_nhits.reserve(sizeof(int)) || h->d_hits.reserve(1024 * sizeof(CorrHit)) ||
      cudaMemset(h->d_state.p, 0, (size_t)nchan * sizeof(CorrChanState)) != cudaSuccess) {
    delete h;
    return nullptr;
  }
  return h;
}
void grcuda_correlate_access_code_bb_destroy(grcuda_corr* h) { delete h; }
int grcuda_correlate_access_code_bb_set_access_code(grcuda_corr* h, const char* code) { return h->set_code(code); }
int grcuda_correlate_access_code_bb_work(grcuda_corr* h, int n, const unsigned char* in, unsigned char* out) {
  if (h->nchan != 1) return set_error(GRCUDA_EINVAL, "host work() is the single-stream form (nchan == 1)");
  if (n <= 0) return 0;
  int rc;
  if ((rc = h->d_in.reserve((size_t)n)) || (rc = h->d_out.reserve((size_t)n))) return rc;
  if ((rc = h->stager.h2d(h->d_in.p, in, (size_t)n, h->stream))) return rc;
  GRB_CUDA(cudaMemsetAsync(h->d_nhits.p, 0, sizeof(int), h->stream));
  if ((rc = h->launch(h->d_in.as<unsigned char>(), nullptr, n, nullptr, 0, 0, h->d_out.as<unsigned char>(),
                      h->d_hits.as<CorrHit>(), 1024, h->d_nhits.as<int>(), h->stream)))
    return rc;
  if ((rc = h->stager.d2h(out, h->d_out.p, (size_t)n, h->stream))) return rc;
  return n;
}
int grcuda_correlate_access_code_bb_work_symbols_device(grcuda_corr* h, const unsigned char* d_symbols, int sym_rows,
                                                        const int* d_counts, const int* map, int nmap,
                                                        int bits_per_symbol, unsigned char* d_out, int out_rows,
                                                        grcuda_hit* d_hits, int max_hits, int* d_nhits, void* stream) {
  (void)out_rows;
  static_assert(sizeof(grcuda_hit) == sizeof(CorrHit), "hit record layout");
  return h->launch(d_symbols, d_counts, sym_rows, map, nmap, bits_per_symbol, d_out, (CorrHit*)d_hits, max_hits,
                   d_nhits, h->pick(stream));
}

}  // extern "C"

// ---- internal accessors (chain.cu) -------------------------------------------------------------
#include "internal.h"
namespace grb {
// M&M + slicer with the map/unpack/correlator fused into the same kernel (chain.cu's tail stage)
int mm_corr_launch(grcuda_mm* mm, grcuda_corr* corr, const int* map, int nmap, int bits_per_symbol, const float* d_in,
                   long ninput, long abs_row0, float* d_soft, unsigned char* d_sym, int max_out, int* d_counts,
                   unsigned char* d_bytes, grcuda_hit* d_hits, int max_hits, int* d_nhits, cudaStream_t s) {
  if (mm->nchan != corr->nchan) return set_error(GRCUDA_EINVAL, "mm/corr channel counts differ");
  MMCorrFuse f;
  memset(&f, 0, sizeof f);
  f.on = 1;
  f.bits_per_symbol = bits_per_symbol;
  for (int i = 0; i < 256; i++) f.map[i] = (unsigned char)i;
  for (int i = 0; i < std::min(nmap, 256); i++) f.map[i] = (unsigned char)map[i];
  f.out = d_bytes;
  f.state = corr->d_state.as<CorrChanState>();
  { std::lock_guard<std::mutex> lk(corr->mu); f.p = corr->p; }
  f.hits = (CorrHit*)d_hits;
  f.max_hits = max_hits;
  f.nhits = d_nhits;
  return mm->launch(d_in, ninput, abs_row0, d_soft, d_sym, max_out, d_counts, s, &f);
}
// the same stage as two kernels: clock recovery + slicer (soft symbols and decisions to HBM), then the correlator
// parallel over channels AND time.  The correlator leaves the kernel whose per-symbol latency bounds the tail (its
// post warps were the bottleneck of the fused form: 0.73 ms against 0.56 ms without it) and the serial chain of a
// time-sharded run.  Returns GRCUDA_EUNSUPPORTED when the parallel correlator does not apply.
int mm_then_corr_launch(grcuda_mm* mm, grcuda_corr* corr, const int* map, int nmap, int bits_per_symbol, const float* d_in,
                        long ninput, long abs_row0, float* d_soft, unsigned char* d_sym, int max_out, int* d_counts,
                        grcuda_hit* d_hits, int max_hits, int* d_nhits, cudaStream_t s) {
  if (mm->nchan != corr->nchan) return set_error(GRCUDA_EINVAL, "mm/corr channel counts differ");
  if (!corr->par_ok(bits_per_symbol, nullptr) || !d_sym) return GRCUDA_EUNSUPPORTED;
  int rc = mm->launch(d_in, ninput, abs_row0, d_soft, d_sym, max_out, d_counts, s, nullptr);
  if (rc) return rc;
  return corr->launch_par(d_sym, d_counts, max_out, map, nmap, (CorrHit*)d_hits, max_hits, d_nhits, s);
}
// the two halves separately, for a time shard: the clock-recovery kernel reads the loop state where the left
// neighbour's was received and writes its final state where the right neighbour's send starts from
int mm_only_launch(grcuda_mm* mm, const float* d_in, long ninput, long abs_row0, float* d_soft, unsigned char* d_sym, int max_out,
                   int* d_counts, const void* d_state_in, void* d_state_out, cudaStream_t s) {
  return mm->launch(d_in, ninput, abs_row0, d_soft, d_sym, max_out, d_counts, s, nullptr, d_state_in, d_state_out);
}
int corr_par_launch(grcuda_corr* corr, const int* map, int nmap, int bits_per_symbol, const unsigned char* d_sym, const int* d_counts,
                    int max_out, grcuda_hit* d_hits, int max_hits, int* d_nhits, const void* d_state_in, void* d_state_out,
                    cudaStream_t s) {
  if (!corr->par_ok(bits_per_symbol, nullptr)) return GRCUDA_EUNSUPPORTED;
  return corr->launch_par(d_sym, d_counts, max_out, map, nmap, (CorrHit*)d_hits, max_hits, d_nhits, s, d_state_in, d_state_out);
}
const float* fir_fff_reversed_taps(grcuda_fir_fff* h, int* ntaps, int* order) {
  if (ntaps) *ntaps = h->ntaps;
  if (order) *order = h->order;
  return h->rt_host.data();
}
// sizes the channelizer's intermediate once (a later growth would cudaFree = device-wide synchronisation in the
// middle of a stream of blocks, which can deadlock against NCCL operations a peer is waiting to pair up)
// the chain overlaps the channelizer with the previous block's clock recovery: use the FFT build that co-resides
int pfb_prefer_coresident_fft(grcuda_pfb* h) {
  FftPlan* p = fft_plan_create((int)h->M, +1, true);
  if (!p) return grcuda_last_error_code();
  fft_plan_destroy(h->fft);
  h->fft = p;
  return GRCUDA_OK;
}
int pfb_demod_supported(grcuda_pfb* h) { return h->rr == (int)h->M && h->TT && fft_plan_demod_supported(h->fft) ? 1 : 0; }
int pfb_work_device_demod(grcuda_pfb* h, long nrows, const float2* d_in_rows, float* d_D, float gain, const float2* prev_y,
                          float2* last_y, bool coresident, cudaStream_t s, float2* y_out) {
  if (!pfb_demod_supported(h)) return set_error(GRCUDA_EUNSUPPORTED, "pfb_channelizer_ccf: no fused discriminator path for this plan");
  DeviceTables tabs;
  int rc = get_tables(&tabs);
  if (rc) return rc;
  grcuda_pfb::DemodOut o = {d_D, gain, tabs.atan, prev_y, last_y, coresident, y_out};
  return h->run(d_in_rows, nullptr, nrows, s, &o);
}
int pfb_reserve_rows(grcuda_pfb* h, long rows) { return h->d_u.reserve((size_t)rows * h->M * sizeof(float2)); }
const float* fir_fff_front_taps(grcuda_fir_fff* h) { return h->has_front_tp ? h->d_front_tp.as<float>() : nullptr; }
const float* fir_fff_front_taps_host(grcuda_fir_fff* h) { return h->has_front_tp ? h->h_front_tp.data() : nullptr; }
float quad_gain(grcuda_quad* h) { std::lock_guard<std::mutex> lk(h->mu); return h->gain; }
void* mm_state_ptr(grcuda_mm* h) { return h->d_state.p; }
// sums over channels of the two per-channel counters of the loop: steps clamped at the first buffered row, and calls
// that stopped at the output capacity before the input ran out (both are where the output leaves the reference's)
int mm_counters(grcuda_mm* h, long long* clamped, long long* overflow) {
  std::vector<MMChanState> st(h->nchan);
  GRB_CUDA(cudaDeviceSynchronize());
  GRB_CUDA(cudaMemcpy(st.data(), h->d_state.p, st.size() * sizeof(MMChanState), cudaMemcpyDeviceToHost));
  long long c = 0, o = 0;
  for (auto& s : st) { c += s.clamped; o += s.overflow; }
  if (clamped) *clamped = c;
  if (overflow) *overflow = o;
  return GRCUDA_OK;
}
size_t mm_state_bytes(grcuda_mm* h) { return (size_t)h->nchan * sizeof(MMChanState); }
void* corr_state_ptr(grcuda_corr* h) { return h->d_state.p; }
size_t corr_state_bytes(grcuda_corr* h) { return (size_t)h->nchan * sizeof(CorrChanState); }
}  // namespace grb
