// gr_pfb_arb_resampler_ccf on the GPU (reference:
// gnuradio-core/src/lib/filter/gr_pfb_arb_resampler_ccf.cc:42-205, .h:159-163).
//
// The reference walks a bank of `filter_size` polyphase filters with a float accumulator: output i uses filter
// j_i on the input window starting at count_i and adds acc_i times the derivative filter's output.  The
// recurrence (count, j, acc) does not look at the data, only at the rate: it is the block's SCHEDULE.  The plan
// runs that recurrence on the host exactly as written (float accumulator, fmodf, the float round trip of
// `count += ss`) -- it is control flow, a few nanoseconds per output row, shared by every channel -- and ships
// the three small arrays to the device; all sample arithmetic is the kernel below.
//
// Data layout: [time][channel] like the rest of the demod tail (nchan = 1 is the reference's single stream).
// A thread makes one output item: two dot products of taps_per_filter taps over the same window, in the
// reference's generic summation order (gr_fir_XXX_generic.cc.t:28-55, two accumulators, no FMA), so the result
// is bit identical to gr_fir_ccf_generic and within ~2e-7 of the SSE class x86-64 GNU Radio selects.  With
// nchan >= 32 a warp is 32 neighbouring channels of one output row: loads and stores are full lines, the filter
// index is warp uniform (tap loads broadcast) and consecutive output rows re-read their overlapping windows
// from L1/L2, so HBM sees the algorithmic bytes (8 B in per input item, 8 B out per output item).
#include <algorithm>
#include <cmath>
#include <vector>

#include "common.cuh"
#include "internal.h"

using namespace grb;

namespace {

struct ArbArgs {
  const float2* in;            // row 0 = first history row
  float2* out;
  const int* cnt;              // [nout] input row of the window start
  const unsigned short* flt;   // [nout] filter index
  const float* acc;            // [nout] interpolation weight
  const float* rt;             // [int_rate][T] reversed taps (rt[k] pairs with in[count + k])
  const float* rdt;            // [int_rate][T] reversed derivative taps
  int T, M;
  long nitems;                 // nout * M
};

__device__ __forceinline__ float2 arb_dot(const float* __restrict__ t, int T, const float2* __restrict__ x, size_t M) {
  float a0r = 0.f, a0i = 0.f, a1r = 0.f, a1i = 0.f;
  int k = 0;
  for (; k + 2 <= T; k += 2) {
    const float2 v0 = __ldg(x + (size_t)k * M), v1 = __ldg(x + (size_t)(k + 1) * M);
    const float t0 = __ldg(t + k), t1 = __ldg(t + k + 1);
    a0r = __fadd_rn(a0r, __fmul_rn(t0, v0.x)); a0i = __fadd_rn(a0i, __fmul_rn(t0, v0.y));
    a1r = __fadd_rn(a1r, __fmul_rn(t1, v1.x)); a1i = __fadd_rn(a1i, __fmul_rn(t1, v1.y));
  }
  for (; k < T; k++) {
    const float2 v0 = __ldg(x + (size_t)k * M);
    const float t0 = __ldg(t + k);
    a0r = __fadd_rn(a0r, __fmul_rn(t0, v0.x)); a0i = __fadd_rn(a0i, __fmul_rn(t0, v0.y));
  }
  return make_float2(__fadd_rn(a0r, a1r), __fadd_rn(a0i, a1i));
}

__global__ void __launch_bounds__(256) pfb_arb_kernel(const ArbArgs a) {
  const size_t M = (size_t)a.M;
  for (long id = (long)blockIdx.x * blockDim.x + threadIdx.x; id < a.nitems; id += (long)gridDim.x * blockDim.x) {
    const long i = id / a.M;
    const int c = (int)(id - i * a.M);
    const int j = a.flt[i];
    const float2* x = a.in + (size_t)a.cnt[i] * M + c;
    const float2 o0 = arb_dot(a.rt + (size_t)j * a.T, a.T, x, M);    // :178
    const float2 o1 = arb_dot(a.rdt + (size_t)j * a.T, a.T, x, M);   // :179
    const float w = a.acc[i];
    a.out[id] = make_float2(__fadd_rn(o0.x, __fmul_rn(o1.x, w)), __fadd_rn(o0.y, __fmul_rn(o1.y, w)));  // :181
  }
}

// ---- tiled kernel for the batched form (nchan >= 32) ------------------------------------------------------
// A CTA owns a strip of ARB_CH neighbouring channels and a chunk of consecutive output rows.  The input rows that
// chunk touches (its windows overlap almost completely) are staged in shared memory once, with the polyphase taps
// next to them, and a thread then makes the outputs of TWO channels (lane and lane + 64) of one output row: per tap
// two LDS.64 (the complex samples, each used by both filters), two broadcast LDS.32 (tap and derivative tap, shared
// by the two channels) and 4 FMUL2 + 4 FFMA2.  The kernel is bound by shared-memory wavefronts (a broadcast LDS.128
// of pre-duplicated taps costs four of them: the first version ran at half this speed), not by HBM.
// The FFMA2 is acc + p issued as fma(p, 1, acc) with the 1 from a kernel argument: packed, one rounding, and ptxas
// cannot contract the multiply into it -- the reference's separate multiply and add (as in kernel_demod_front.cuh).
#define ARB_CH 128
#define ARB_THREADS 256
#define ARB_MAXROWS 56      // staged input rows per CTA: 56 x 128 channels x 8 B = 56 KB

typedef unsigned long long arb_u64;
__device__ __forceinline__ arb_u64 arb_pack(float a, float b) {
  arb_u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ arb_u64 arb_mul2(arb_u64 a, arb_u64 b) {
  arb_u64 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ arb_u64 arb_add2(arb_u64 p, arb_u64 ones, arb_u64 acc) {  // acc + p, two lanes
  arb_u64 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(p), "l"(ones), "l"(acc));
  return r;
}

struct ArbTileArgs {
  ArbArgs a;
  int nout;        // output rows
  int chunk;       // output rows per CTA
  int ntaps_all;   // int_rate * T
  float one;       // 1.0f, opaque to the compiler
};

// one tap of both filters on the samples of the two channels; S = accumulator set (0: even taps, 1: odd taps)
#define ARB_TAP(S, k_, V0, V1)                                   \
  {                                                              \
    const float t_ = tp[k_], d_ = dp[k_];                        \
    const arb_u64 tt_ = arb_pack(t_, t_), dd_ = arb_pack(d_, d_); \
    a##S = arb_add2(arb_mul2(tt_, V0), ones, a##S);              \
    b##S = arb_add2(arb_mul2(dd_, V0), ones, b##S);              \
    e##S = arb_add2(arb_mul2(tt_, V1), ones, e##S);              \
    f##S = arb_add2(arb_mul2(dd_, V1), ones, f##S);              \
  }

__global__ void __launch_bounds__(ARB_THREADS) pfb_arb_tile_kernel(const ArbTileArgs ta) {
  extern __shared__ __align__(16) unsigned char arb_smem[];
  const ArbArgs& a = ta.a;
  const int T = a.T, M = a.M;
  arb_u64* xs = reinterpret_cast<arb_u64*>(arb_smem);                   // [rows][ARB_CH] complex samples
  float* ts = reinterpret_cast<float*>(xs + (size_t)ARB_MAXROWS * ARB_CH);  // [int_rate * T] reversed taps
  float* ds = ts + ta.ntaps_all;                                        // [int_rate * T] reversed derivative taps
  const int c0 = blockIdx.x * ARB_CH;
  const int i0 = blockIdx.y * ta.chunk, i1 = min(ta.nout, i0 + ta.chunk);
  const int row0 = a.cnt[i0];
  const int nrows = a.cnt[i1 - 1] + T - row0;                           // cnt is non-decreasing
  const bool staged = nrows <= ARB_MAXROWS;                             // CTA uniform
  for (int k = threadIdx.x; k < ta.ntaps_all; k += ARB_THREADS) {
    ts[k] = a.rt[k];
    ds[k] = a.rdt[k];
  }
  if (staged) {
    const arb_u64* g = reinterpret_cast<const arb_u64*>(a.in) + (size_t)row0 * M + c0;
    for (int e = threadIdx.x; e < nrows * ARB_CH; e += ARB_THREADS) {
      const int r = e / ARB_CH, c = e % ARB_CH;
      xs[e] = (c0 + c < M) ? __ldg(g + (size_t)r * M + c) : 0ull;
    }
  }
  __syncthreads();
  constexpr int HALF = ARB_CH / 2;
  const int cl = threadIdx.x % HALF, sub = threadIdx.x / HALF;
  const int c = c0 + cl;
  if (c >= M) return;
  const bool two = c + HALF < M;                                        // the second channel of this thread exists
  const arb_u64 ones = arb_pack(ta.one, ta.one);
  for (int i = i0 + sub; i < i1; i += ARB_THREADS / HALF) {
    const int cnt = a.cnt[i];
    const float* tp = ts + (size_t)a.flt[i] * T;
    const float* dp = ds + (size_t)a.flt[i] * T;
    // a/b: taps and derivative taps on channel c; e/f: on channel c + HALF; 0 / 1: even / odd taps
    // (gr_fir_XXX_generic.cc.t:33-44)
    arb_u64 a0 = 0ull, a1 = 0ull, b0 = 0ull, b1 = 0ull, e0 = 0ull, e1 = 0ull, f0 = 0ull, f1 = 0ull;
    if (staged) {
      const arb_u64* x = xs + (size_t)(cnt - row0) * ARB_CH + cl;
      int k = 0;
      for (; k + 2 <= T; k += 2) {
        const arb_u64 v0 = x[(size_t)k * ARB_CH], v1 = x[(size_t)(k + 1) * ARB_CH];
        const arb_u64 u0 = x[(size_t)k * ARB_CH + HALF], u1 = x[(size_t)(k + 1) * ARB_CH + HALF];
        ARB_TAP(0, k, v0, u0)
        ARB_TAP(1, k + 1, v1, u1)
      }
      if (k < T) {
        const arb_u64 v0 = x[(size_t)k * ARB_CH], u0 = x[(size_t)k * ARB_CH + HALF];
        ARB_TAP(0, k, v0, u0)
      }
    } else {  // a chunk whose windows are far apart (strong decimation): straight from global memory
      const arb_u64* x = reinterpret_cast<const arb_u64*>(a.in) + (size_t)cnt * M + c;
      const size_t h2 = two ? HALF : 0;
      int k = 0;
      for (; k + 2 <= T; k += 2) {
        const arb_u64 v0 = __ldg(x + (size_t)k * M), v1 = __ldg(x + (size_t)(k + 1) * M);
        const arb_u64 u0 = __ldg(x + (size_t)k * M + h2), u1 = __ldg(x + (size_t)(k + 1) * M + h2);
        ARB_TAP(0, k, v0, u0)
        ARB_TAP(1, k + 1, v1, u1)
      }
      if (k < T) {
        const arb_u64 v0 = __ldg(x + (size_t)k * M), u0 = __ldg(x + (size_t)k * M + h2);
        ARB_TAP(0, k, v0, u0)
      }
    }
    const float w = a.acc[i];
    const arb_u64 ww = arb_pack(w, w);
    arb_u64* o = reinterpret_cast<arb_u64*>(a.out) + (size_t)i * M + c;
    // acc0 + acc1 (:46), then o0 + o1 * d_acc (gr_pfb_arb_resampler_ccf.cc:181)
    o[0] = arb_add2(arb_mul2(arb_add2(b1, ones, b0), ww), ones, arb_add2(a1, ones, a0));
    if (two) o[HALF] = arb_add2(arb_mul2(arb_add2(f1, ones, f0), ww), ones, arb_add2(e1, ones, e0));
  }
}
#undef ARB_TAP

struct HostPin {
  void* p = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes) {
    if (bytes <= cap) return GRCUDA_OK;
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
    const cudaError_t e = cudaHostAlloc(&p, bytes, cudaHostAllocDefault);
    if (e != cudaSuccess) return set_error(GRCUDA_ENOMEM, "cudaHostAlloc(%zu): %s", bytes, cudaGetErrorString(e));
    cap = bytes;
    return GRCUDA_OK;
  }
  ~HostPin() { if (p) cudaFreeHost(p); }
};

}  // namespace

struct grcuda_pfb_arb {
  int nchan = 1;
  unsigned int_rate = 32, dec_rate = 0, last_filter = 0, T = 0;
  float flt_rate = 0, acc = 0, rate = 1;
  int start_index = 0;
  bool updated = true;
  std::vector<float> fwd, dfwd;  // [int_rate][T] forward-order taps of each filter (print_taps / get_taps)
  cudaStream_t stream = nullptr;
  cudaEvent_t ev_sched = nullptr;  // the schedule of the previous call has left the pinned buffer
  bool sched_pending = false;
  Stager stager;
  DevBuf d_in, d_out, d_rt, d_rdt, d_sched;
  HostPin h_sched;
  std::mutex mu;
  ~grcuda_pfb_arb() {
    if (ev_sched) cudaEventDestroy(ev_sched);
    if (stream) cudaStreamDestroy(stream);
  }
  void set_rate_locked(float r) {  // gr_pfb_arb_resampler_ccf.h:159-163
    dec_rate = (unsigned)floor(int_rate / r);
    flt_rate = (int_rate / r) - dec_rate;
    rate = r;
  }
  // create_taps (:90-125): filter i holds tmp[i + j * int_rate], zero padded to T taps
  void partition(const std::vector<float>& nt, std::vector<float>& ours) const {
    ours.assign((size_t)int_rate * T, 0.f);
    for (unsigned i = 0; i < int_rate; i++)
      for (unsigned j = 0; j < T; j++) {
        const size_t k = i + (size_t)j * int_rate;
        if (k < nt.size()) ours[(size_t)i * T + j] = nt[k];
      }
  }
  // general_work's loop (:155-205) without the filter calls: fills the schedule of up to noutput items
  int schedule(int ninput, int noutput, int* cnt, unsigned short* flt, float* w, int* consumed) {
    int i = 0, count = start_index;
    unsigned j = last_filter;
    const int max_input = ninput - (int)T;
    while (i < noutput && count < max_input) {
      while (j < int_rate && i < noutput) {
        cnt[i] = count;
        flt[i] = (unsigned short)j;
        w[i] = acc;
        i++;
        acc += flt_rate;
        j += dec_rate + (int)floorf(acc);
        acc = fmodf(acc, 1.0f);
      }
      if (i < noutput) {
        const float ss = (float)(int)(j / int_rate);  // `float ss = (int)(j / d_int_rate); count += ss;`
        count = (int)((float)count + ss);
        j = j % int_rate;
      }
    }
    last_filter = j;
    start_index = std::max(0, count - ninput);
    *consumed = std::min(count, ninput);
    return i;
  }
};

extern "C" {

grcuda_pfb_arb* grcuda_pfb_arb_resampler_ccf_create(float rate, const float* taps, int ntaps, unsigned filter_size,
                                                    int nchan) {
  if (!taps || ntaps < 2) { set_error(GRCUDA_EINVAL, "pfb_arb_resampler_ccf: at least 2 prototype taps are needed"); return nullptr; }
  if (filter_size < 1 || filter_size > 65535u) { set_error(GRCUDA_EINVAL, "pfb_arb_resampler_ccf: filter_size %u", filter_size); return nullptr; }
  if (!(rate > 0.f)) { set_error(GRCUDA_EINVAL, "pfb_arb_resampler_ccf: rate must be > 0"); return nullptr; }
  if (nchan < 1) { set_error(GRCUDA_EINVAL, "pfb_arb_resampler_ccf: nchan < 1"); return nullptr; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
    cudaGetLastError();
    set_error(GRCUDA_ECUDA, "no CUDA device available; libgr_cuda has no CPU fallback");
    return nullptr;
  }
  grcuda_pfb_arb* h = new grcuda_pfb_arb;
  h->nchan = nchan;
  h->int_rate = filter_size;
  h->set_rate_locked(rate);
  h->T = (unsigned)ceil((double)ntaps / (double)filter_size);
  std::vector<float> nt(taps, taps + ntaps), dt(ntaps);
  float tap = 0.f;  // create_diff_taps (:127-140): first difference, last one duplicated
  for (int i = 0; i < ntaps - 1; i++) { tap = nt[i + 1] - nt[i]; dt[i] = tap; }
  dt[ntaps - 1] = tap;
  h->partition(nt, h->fwd);
  h->partition(dt, h->dfwd);
  std::vector<float> rt(h->fwd.size()), rdt(h->fwd.size());
  for (unsigned i = 0; i < h->int_rate; i++)
    for (unsigned k = 0; k < h->T; k++) {  // gr_fir stores the taps reversed (gr_fir_XXX.h.t:51,65)
      rt[(size_t)i * h->T + k] = h->fwd[(size_t)i * h->T + (h->T - 1 - k)];
      rdt[(size_t)i * h->T + k] = h->dfwd[(size_t)i * h->T + (h->T - 1 - k)];
    }
  const size_t tb = rt.size() * sizeof(float);
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_sched, cudaEventDisableTiming) != cudaSuccess || h->d_rt.reserve(tb) || h->d_rdt.reserve(tb) ||
      cudaMemcpy(h->d_rt.p, rt.data(), tb, cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(h->d_rdt.p, rdt.data(), tb, cudaMemcpyHostToDevice) != cudaSuccess) {
    set_error(GRCUDA_ECUDA, "pfb_arb_resampler_ccf: device setup failed: %s", cudaGetErrorString(cudaGetLastError()));
    delete h;
    return nullptr;
  }
  return h;
}
void grcuda_pfb_arb_resampler_ccf_destroy(grcuda_pfb_arb* h) { delete h; }
int grcuda_pfb_arb_resampler_ccf_set_rate(grcuda_pfb_arb* h, float rate) {
  if (!(rate > 0.f)) return set_error(GRCUDA_EINVAL, "pfb_arb_resampler_ccf: rate must be > 0");
  std::lock_guard<std::mutex> lk(h->mu);
  h->set_rate_locked(rate);
  return GRCUDA_OK;
}
unsigned grcuda_pfb_arb_resampler_ccf_history(grcuda_pfb_arb* h) { return h->T + 1; }  // :121
double grcuda_pfb_arb_resampler_ccf_relative_rate(grcuda_pfb_arb* h) { std::lock_guard<std::mutex> lk(h->mu); return h->rate; }
int grcuda_pfb_arb_resampler_ccf_taps_per_filter(grcuda_pfb_arb* h) { return (int)h->T; }
int grcuda_pfb_arb_resampler_ccf_filter_size(grcuda_pfb_arb* h) { return (int)h->int_rate; }
int grcuda_pfb_arb_resampler_ccf_get_taps(grcuda_pfb_arb* h, int filter, int derivative, float* out, int cap) {
  if (filter < 0 || (unsigned)filter >= h->int_rate) return set_error(GRCUDA_ERANGE, "pfb_arb_resampler_ccf: filter %d", filter);
  const std::vector<float>& v = derivative ? h->dfwd : h->fwd;
  const int n = std::min<int>(cap, (int)h->T);
  for (int k = 0; k < n; k++) out[k] = v[(size_t)filter * h->T + k];
  return (int)h->T;
}

int grcuda_pfb_arb_resampler_ccf_work_device(grcuda_pfb_arb* h, int noutput_items, int ninput_items,
                                             const grcuda_complex* d_in, grcuda_complex* d_out, int* consumed,
                                             void* stream_) {
  if (consumed) *consumed = 0;
  if (noutput_items < 0 || ninput_items < 0) return set_error(GRCUDA_EINVAL, "pfb_arb_resampler_ccf: negative item count");
  if (h->updated) {  // :166-169
    h->updated = false;
    return 0;
  }
  if (noutput_items == 0) return 0;
  cudaStream_t s = stream_ ? (cudaStream_t)stream_ : h->stream;
  // schedule arrays: [cnt int32 | acc float | flt uint16], pinned on the host, one copy to the device
  const size_t n = (size_t)noutput_items;
  const size_t bytes = n * 10;
  int rc;
  if (h->sched_pending) {  // the previous call's copy must have read the pinned buffer before it is rewritten
    GRB_CUDA(cudaEventSynchronize(h->ev_sched));
    h->sched_pending = false;
  }
  if ((rc = h->h_sched.reserve(bytes)) || (rc = h->d_sched.reserve(bytes))) return rc;
  int* cnt = (int*)h->h_sched.p;
  float* w = (float*)(cnt + n);
  unsigned short* flt = (unsigned short*)(w + n);
  int cons = 0, produced;
  {
    std::lock_guard<std::mutex> lk(h->mu);
    produced = h->schedule(ninput_items, noutput_items, cnt, flt, w, &cons);
  }
  if (consumed) *consumed = cons;
  if (produced == 0) return 0;
  GRB_CUDA(cudaMemcpyAsync(h->d_sched.p, h->h_sched.p, bytes, cudaMemcpyHostToDevice, s));
  GRB_CUDA(cudaEventRecord(h->ev_sched, s));
  h->sched_pending = true;
  ArbArgs a;
  a.in = (const float2*)d_in; a.out = (float2*)d_out;
  a.cnt = (const int*)h->d_sched.p; a.acc = (const float*)((const int*)h->d_sched.p + n);
  a.flt = (const unsigned short*)((const float*)a.acc + n);
  a.rt = h->d_rt.as<float>(); a.rdt = h->d_rdt.as<float>();
  a.T = (int)h->T; a.M = h->nchan; a.nitems = (long)produced * h->nchan;
  const size_t tap_bytes = (size_t)h->int_rate * h->T * 2 * sizeof(float);
  if (h->nchan >= 32 && tap_bytes <= 40 * 1024) {
    // tiled path: the chunk is sized so that its input rows fit the staging area at this rate (a chunk that does not
    // fit after all, e.g. right after set_rate, reads global memory directly inside the kernel)
    ArbTileArgs ta;
    ta.a = a; ta.nout = produced; ta.ntaps_all = (int)(h->int_rate * h->T); ta.one = 1.0f;
    const double rows_per_out = 1.0 / (double)h->rate;
    const int fit = (int)std::floor((ARB_MAXROWS - (int)h->T - 2) / std::max(rows_per_out, 1e-9));
    ta.chunk = std::max(4, std::min(128, fit));
    const size_t smem = (size_t)ARB_MAXROWS * ARB_CH * sizeof(arb_u64) + tap_bytes;
    GRB_CUDA(raise_dynamic_smem((const void*)pfb_arb_tile_kernel, (size_t)smem));
    dim3 grid((h->nchan + ARB_CH - 1) / ARB_CH, (produced + ta.chunk - 1) / ta.chunk);
    pfb_arb_tile_kernel<<<grid, ARB_THREADS, smem, s>>>(ta);
  } else {
    const long blocks = (a.nitems + 255) / 256;
    const int grid = (int)std::max<long>(1, std::min<long>(blocks, (long)sm_count() * 16));
    pfb_arb_kernel<<<grid, 256, 0, s>>>(a);
  }
  GRB_LAUNCH_CHECK();
  return produced;
}

int grcuda_pfb_arb_resampler_ccf_work(grcuda_pfb_arb* h, int noutput_items, int ninput_items, const grcuda_complex* in,
                                      grcuda_complex* out, int* consumed) {
  if (consumed) *consumed = 0;
  if (h->nchan != 1) return set_error(GRCUDA_EINVAL, "host work() is the single-stream form (nchan == 1)");
  if (noutput_items < 0 || ninput_items < 0) return set_error(GRCUDA_EINVAL, "pfb_arb_resampler_ccf: negative item count");
  if (h->updated) {
    h->updated = false;
    return 0;
  }
  if (noutput_items == 0 || ninput_items == 0) return 0;
  int rc;
  if ((rc = h->d_in.reserve((size_t)ninput_items * sizeof(float2))) ||
      (rc = h->d_out.reserve((size_t)noutput_items * sizeof(float2))))
    return rc;
  if ((rc = h->stager.h2d(h->d_in.p, in, (size_t)ninput_items * sizeof(float2), h->stream))) return rc;
  const int produced = grcuda_pfb_arb_resampler_ccf_work_device(h, noutput_items, ninput_items, (const grcuda_complex*)h->d_in.p,
                                                                (grcuda_complex*)h->d_out.p, consumed, h->stream);
  if (produced <= 0) return produced;
  if ((rc = h->stager.d2h(out, h->d_out.p, (size_t)produced * sizeof(float2), h->stream))) return rc;
  return produced;
}

}  // extern "C"
