// FIR-class kernels of the hot path: decimating complex FIR with real or complex taps
// (gr_fir_filter_ccf, gr_freq_xlating_fir_filter_ccf) and the polyphase branch filters of
// gr_pfb_channelizer_ccf.  All of them are HBM-bound by design: every
// input sample is read from HBM once, every output written once; reuse lives in registers /
// shared memory.
#pragma once
#include <cuda_runtime.h>
#include "fft_radix.cuh"
#include "tma.cuh"

namespace grb {

// ===========================================================================================
// Decimating FIR, complex input, real (CTAPS=false) or complex (CTAPS=true) taps.
//   out[o] = sum_{i<ntaps} rt[i] * in[o*D + i]        (in history-prefixed, rt = reversed taps;
//   gr_fir_XXX_generic.cc.t:83-103, gr_fir_filter_XXX.cc.t:81-85)
// Polyphase form so that a thread owns R CONSECUTIVE outputs and slides a register window:
//   i = D*q + p  ->  out[o] = sum_p sum_q rtp[p][q] * xp[p][o + q],   xp[p][n] = in[n*D + p]
// The CTA stages its input span once into shared memory, split by phase p (conflict-free both
// for the coalesced fill and for the strided register-window reads: row pitch = 4 mod 16 float2,
// in-row index padded n + n/8 so that threads R=8 apart hit distinct banks).
// Per phase and tap the thread does R (x2, x4 for complex taps) FMAs for ONE new LDS.64.
// Optional epilogue: multiply by the running rotator e^{j*(o_abs)*theta} (freq_xlating).
// ===========================================================================================
struct FirArgs {
  const float2* in;   // history-prefixed input
  float2* out;
  long nout;
  int decim;
  int ntaps;
  int J;              // taps per phase, padded to a multiple of FIR_R
  const float* rtp;   // [decim][J] floats (real taps) or float2 (complex taps), zero padded
  int tile_out;       // outputs per CTA tile = blockDim.x * FIR_R
  int pitch;          // smem float2 per phase row (padded)
  // rotator epilogue (freq_xlating): phase(o) = theta * (o + out_index0), evaluated in double
  int rotate;
  double theta;
  long out_index0;
  float2 rot_step[8];  // e^{j theta r}, r < FIR_R (evaluated in double on the host)
};

#define FIR_R 8

__device__ __forceinline__ int fir_phys(int n) { return n + (n >> 3); }

template <bool CTAPS>
__global__ void __launch_bounds__(128) fir_decim_kernel(const FirArgs a) {
  extern __shared__ __align__(16) float2 fir_smem[];
  const int D = a.decim, J = a.J;
  float2* xs = fir_smem;                                   // [D][pitch]
  float* taps_s = reinterpret_cast<float*>(xs + (size_t)D * a.pitch);  // [D][J] (x2 if complex)
  const int ntap_words = D * J * (CTAPS ? 2 : 1);
  for (int i = threadIdx.x; i < ntap_words; i += blockDim.x) taps_s[i] = a.rtp[i];

  const long ntiles = (a.nout + a.tile_out - 1) / a.tile_out;
  for (long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long o_base = tile * a.tile_out;
    const int tile_n = (int)min((long)a.tile_out, a.nout - o_base);
    // input span: phase rows hold n in [0, tile_n_pad + J), idx = (o_base + n)*D + p
    const int rows_n = a.tile_out + J;
    const long in_base = o_base * D;
    const long in_valid = (a.nout - 1) * (long)D + a.ntaps;  // number of valid input items
    __syncthreads();  // previous tile fully consumed (also orders the taps fill)
    const int span = rows_n * D;
    {
      // Phase-split fill, global -> shared without register staging (cp.async, 8 B per element: the
      // scatter by phase rules out a bulk copy) and without a division per element: element
      // s = tid + k*blockDim goes to (n, p) = (s / D, s % D), advanced incrementally.
      const int bd = blockDim.x, dn = bd / D, dp = bd - dn * D;
      int n = threadIdx.x / D, p = threadIdx.x - n * D;
      const float2* __restrict__ g = a.in + in_base;
      const long nvalid = in_valid - in_base;  // elements of this tile that exist
      const unsigned xs_s = smem_u32(xs);
      if (dp == 0 && (dn & 7) == 0 && (long)span <= nvalid) {
        // common case (D divides the block size, interior tile): the phase is fixed per thread and
        // n advances by a multiple of 8, so the padded address advances by a constant
        unsigned dst = xs_s + (unsigned)(p * a.pitch + fir_phys(n)) * 8u;
        const unsigned dstep = (unsigned)(dn + (dn >> 3)) * 8u;
        const float2* __restrict__ src = g + threadIdx.x;
        for (int s = threadIdx.x; s < span; s += bd) {
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
          dst += dstep;
          src += bd;
        }
      } else
      for (int s = threadIdx.x; s < span; s += bd) {
        const unsigned dst = xs_s + (unsigned)(p * a.pitch + fir_phys(n)) * 8u;
        if (s < nvalid) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(g + s) : "memory");
        else xs[(size_t)p * a.pitch + fir_phys(n)] = make_float2(0.f, 0.f);
        n += dn;
        p += dp;
        if (p >= D) { p -= D; n++; }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();

    const int o0 = threadIdx.x * FIR_R;  // first output of this thread within the tile
    float2 acc[FIR_R];
#pragma unroll
    for (int r = 0; r < FIR_R; r++) acc[r] = make_float2(0.f, 0.f);
    if (o0 < tile_n) {
      for (int p = 0; p < D; p++) {
        // o0 and q0 are multiples of 8 = FIR_R, so phys(o0 + q0 + u) = 9*(o0 + q0)/8 + u: one running
        // pointer per phase row and compile-time offsets u
        const float2* xw = xs + (size_t)p * a.pitch + fir_phys(o0);
        float2 w[FIR_R];  // sliding window: w[r] = xp[p][o0 + q + r]
#pragma unroll
        for (int r = 0; r < FIR_R; r++) w[r] = xw[r];
        const float* tq = taps_s + (size_t)p * J * (CTAPS ? 2 : 1);
        for (int q0 = 0; q0 < J; q0 += FIR_R, xw += FIR_R + 1, tq += FIR_R * (CTAPS ? 2 : 1)) {
          float tv[FIR_R * (CTAPS ? 2 : 1)];  // the 8 taps of this group: two (four) LDS.128
#pragma unroll
          for (int v4 = 0; v4 < FIR_R * (CTAPS ? 2 : 1) / 4; v4++) {
            const float4 t4 = reinterpret_cast<const float4*>(tq)[v4];
            tv[4 * v4] = t4.x; tv[4 * v4 + 1] = t4.y; tv[4 * v4 + 2] = t4.z; tv[4 * v4 + 3] = t4.w;
          }
#pragma unroll
          for (int u = 0; u < FIR_R; u++) {
            // tap q = q0+u multiplies window slot (u + r) mod R for output r
            // packed FP32 (FFMA2): one instruction updates the real and the imaginary accumulator
            if (CTAPS) {
              const float2 t = make_float2(tv[2 * u], tv[2 * u + 1]);
#pragma unroll
              for (int r = 0; r < FIR_R; r++) {
                const float2 x = w[(u + r) % FIR_R];
                // (t.x + j t.y)(x.x + j x.y) = t.x * (x.x, x.y) + t.y * (-x.y, x.x)
                acc[r] = cfma(make_float2(-x.y, x.x), t.y, cfma(x, t.x, acc[r]));
              }
            } else {
              const float t = tv[u];
#pragma unroll
              for (int r = 0; r < FIR_R; r++) acc[r] = cfma(w[(u + r) % FIR_R], t, acc[r]);
            }
            // slot u (holding xp[o0+q0+u]) is dead now: refill with xp[o0 + q0 + u + R]
            w[u] = xw[FIR_R + 1 + u];
          }
        }
      }
      if (a.rotate) {
        // closed form of the reference's running rotator (gr_rotator.h:40-50): e^{j theta (index)}.  One
        // double-precision sincos per thread for its first output, the next seven by the host's step table
        const double ph = a.theta * (double)(a.out_index0 + o_base + o0);
        double sn, cs;
        sincos(ph, &sn, &cs);
        const float2 rot0 = make_float2((float)cs, (float)sn);
#pragma unroll
        for (int r = 0; r < FIR_R; r++) acc[r] = cmul(acc[r], cmul(rot0, a.rot_step[r]));
      }
      // 8 consecutive float2 = 64 B per thread; neighbouring threads are contiguous
      float2* o = a.out + o_base + o0;
#pragma unroll
      for (int r = 0; r < FIR_R; r++)
        if (o0 + r < tile_n) o[r] = acc[r];
    }
  }
}

// ===========================================================================================
// Polyphase branch FIR of gr_pfb_channelizer_ccf (oversample_rate == 1):
//   u[m][k] = sum_{t<T} h[k + t*M] * X[m + H - t][M-1-k]          (gr_pfb_channelizer_ccf.cc:171-188)
// X = interleaved input rows [time][M] with H = T history rows in front (history = T+1, :136).
// One thread owns one column j = M-1-k and walks rows_per_thread consecutive rows keeping the
// last TT samples of its column in registers (TT >= T, taps zero padded); neighbouring threads
// read neighbouring columns, so every load/store is a fully coalesced 256 B warp access and each
// input sample is fetched from HBM exactly once per row tile (+ T-1 halo rows per tile).
// taps_t[t*M + j] = h[(M-1-j) + t*M]  (transposed so the per-thread tap fetch is coalesced too).
// ===========================================================================================
struct PfbFirArgs {
  const float2* x;     // [H + nrows][M]
  float2* u;           // [nrows][M]   FFT input order (index k)
  const float* taps_t; // [TT][M]
  int M;
  int T;               // real taps per branch (history rows H = T)
  long nrows;
  int rows_per_thread;
};

template <int TT>
__global__ void __launch_bounds__(128) pfb_fir_kernel(const PfbFirArgs a) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= a.M) return;
  const long m0 = (long)blockIdx.y * a.rows_per_thread;
  if (m0 >= a.nrows) return;
  const long m1 = min(a.nrows, m0 + a.rows_per_thread);
  const int M = a.M, H = a.T;
  float h[TT];
#pragma unroll
  for (int t = 0; t < TT; t++) h[t] = __ldg(a.taps_t + (size_t)t * M + j);
  // window slot s holds buffer row b with b % TT == s; output m needs buffer rows m+H-t, t<TT
  float2 w[TT];
  const float2* __restrict__ xc = a.x + j;
  {
    // preload the TT-1 rows preceding buffer row (m0 + H): rows m0+H-TT+1 .. m0+H-1
#pragma unroll
    for (int s = 0; s < TT; s++) w[s] = make_float2(0.f, 0.f);
#pragma unroll
    for (int d = 1; d < TT; d++) {
      const long b = m0 + H - d;
      if (b >= 0 && d < a.T + 0) {  // rows older than T-1 back only ever meet zero taps
        w[(TT - d) % TT] = __ldg(xc + b * (long)M);
      }
    }
  }
  float2* __restrict__ uc = a.u + (M - 1 - j);
  // rows in groups of TT so that window indices are compile-time; the loads of NB rows are issued
  // back to back BEFORE their arithmetic (NB x 256 B per warp in flight): the kernel is a pure
  // stream and only memory-level parallelism keeps HBM busy
  constexpr int NB = TT < 8 ? TT : 8;
  for (long mb = m0; mb < m1; mb += TT) {
#pragma unroll
    for (int sb = 0; sb < TT; sb += NB) {
      float2 nx[NB];
      const float2* __restrict__ xr = xc + (mb + sb + H) * (long)M;
#pragma unroll
      for (int s = 0; s < NB; s++) nx[s] = (mb + sb + s < m1) ? __ldg(xr + (long)s * M) : make_float2(0.f, 0.f);
#pragma unroll
      for (int s2 = 0; s2 < NB; s2++) {
        // slot for this row: relative index s in the group; the preload put row (m0+H-d) at
        // slot (TT-d)%TT, i.e. row (mb+H+s) belongs to slot s.
        const int s = sb + s2;
        w[s] = nx[s2];
        float2 acc = make_float2(0.f, 0.f);
#pragma unroll
        for (int t = 0; t < TT; t++) {
          const float2 x = w[(s - t + TT) % TT];
          acc.x += h[t] * x.x;
          acc.y += h[t] * x.y;
        }
        if (mb + s < m1) uc[(mb + s) * (long)M] = acc;
      }
    }
  }
}

// Same filter, sample windows staged through shared memory by the bulk copy engine (TMA).
// A CTA owns 256 neighbouring columns x rows_per_cta output rows.  The 2 KB row segments of its
// column tile arrive in stages of 16 rows (32 KB) through cp.async.bulk; four stages (128 KB) are
// in flight per SM whatever the register allocation, which is what a 16 B/sample stream with
// ~1 us of HBM latency needs (the register-window kernel above has 16 warps x 8 loads x 256 B =
// 32 KB in flight per SM and stops at a third of the HBM rate).  128 KB leaves room on the SM for a
// CTA of the clock-recovery kernel of the previous block, which runs concurrently.  Thread = column: one LDS.64 per row
// (conflict free), TT FFMA pairs on the register window, one coalesced store.
#define PFT_COLS 256
#define PFT_SR 16
#define PFT_NST 4
static inline size_t pfb_fir_tma_smem() { return (size_t)PFT_NST * PFT_SR * PFT_COLS * sizeof(float2) + 64; }

template <int TT>
__global__ void __launch_bounds__(PFT_COLS) pfb_fir_tma_kernel(const PfbFirArgs a) {
  extern __shared__ __align__(128) unsigned char pft_smem[];
  float2* stage = reinterpret_cast<float2*>(pft_smem);  // [NST][SR][COLS]
  uint64_t* full = reinterpret_cast<uint64_t*>(pft_smem + (size_t)PFT_NST * PFT_SR * PFT_COLS * sizeof(float2));
  const int tid = threadIdx.x;
  const int M = a.M, T = a.T, H = a.T;
  const int col0 = blockIdx.x * PFT_COLS;
  const int ncols = min(PFT_COLS, M - col0);
  const long m0 = (long)blockIdx.y * a.rows_per_thread;
  const long m1 = min(a.nrows, m0 + a.rows_per_thread);
  if (m0 >= m1) return;
  // sequence of buffer rows this CTA walks: bs + i, i in [0, nseq); output row of step i is m0 + i - (T-1)
  const long bs = m0 + H - (T - 1);
  const int nseq = (int)(m1 - m0) + T - 1;
  const int nstages = (nseq + PFT_SR - 1) / PFT_SR;
  const unsigned row_bytes = (unsigned)ncols * sizeof(float2);

  if (tid == 0) {
    for (int s = 0; s < PFT_NST; s++) mbar_init(full + s, 1);
    mbar_init_fence();
  }
  __syncthreads();
  auto issue = [&](int k) {  // warp 0: the PFT_SR row segments of stage k
    const int i0 = k * PFT_SR;
    const int nr = min(PFT_SR, nseq - i0);
    uint64_t* bar = full + (k % PFT_NST);
    if (tid == 0) mbar_expect_tx(bar, row_bytes * (unsigned)nr);
    __syncwarp();
    if (tid < nr) {
      float2* dst = stage + ((size_t)(k % PFT_NST) * PFT_SR + tid) * PFT_COLS;
      bulk_g2s(dst, a.x + (bs + i0 + tid) * (long)M + col0, row_bytes, bar);
    }
  };
  if (tid < 32)
    for (int k = 0; k < PFT_NST && k < nstages; k++) issue(k);

  const bool cok = tid < ncols;
  const int j = col0 + (cok ? tid : 0);
  float h[TT];
#pragma unroll
  for (int t = 0; t < TT; t++) h[t] = __ldg(a.taps_t + (size_t)t * M + j);
  float2 w[TT];  // slot s holds sequence row i with i % TT == s
#pragma unroll
  for (int s = 0; s < TT; s++) w[s] = make_float2(0.f, 0.f);
  float2* __restrict__ uc = a.u + (M - 1 - j) + (m0 - (T - 1)) * (long)M;  // + i * M = output row of step i

  for (int k = 0; k < nstages; k++) {
    mbar_wait(full + (k % PFT_NST), (unsigned)(k / PFT_NST) & 1u);
    const float2* __restrict__ sp = stage + (size_t)(k % PFT_NST) * PFT_SR * PFT_COLS + tid;
#pragma unroll
    for (int g = 0; g < PFT_SR / TT; g++) {
      const int ig = k * PFT_SR + g * TT;
      float2* __restrict__ up = uc + (long)ig * M;
      // two accumulation chains per component (even / odd taps): the 2 x TT FFMAs of a row are then
      // 4 chains of TT/2 instead of 2 chains of TT dependent instructions
#define PFT_ROW(s_, GUARD)                                                    \
      {                                                                       \
        w[s_] = sp[(g * TT + (s_)) * PFT_COLS];                               \
        float2 a0 = make_float2(0.f, 0.f), a1 = make_float2(0.f, 0.f);        \
        _Pragma("unroll") for (int t = 0; t < TT; t += 2) {                   \
          const float2 x0 = w[((s_) - t + TT) % TT];                          \
          const float2 x1 = w[((s_) - t - 1 + 2 * TT) % TT];                  \
          a0.x += h[t] * x0.x;                                                \
          a0.y += h[t] * x0.y;                                                \
          a1.x += h[t + 1] * x1.x;                                            \
          a1.y += h[t + 1] * x1.y;                                            \
        }                                                                     \
        if (GUARD) *up = make_float2(a0.x + a1.x, a0.y + a1.y);               \
        up += M;                                                              \
      }
      if (ig >= T - 1 && ig + TT <= nseq) {  // CTA uniform: every row of the group is an output row
#pragma unroll
        for (int s = 0; s < TT; s++) PFT_ROW(s, cok)
      } else if (ig < nseq) {                 // first (window fill) and last (ragged) groups
#pragma unroll
        for (int s = 0; s < TT; s++) PFT_ROW(s, cok && ig + s >= T - 1 && ig + s < nseq)
      }
#undef PFT_ROW
    }
    __syncthreads();  // every thread is done with this stage's buffer
    if (tid < 32 && k + PFT_NST < nstages) issue(k + PFT_NST);
  }
}

// Generic / oversampled branch FIR (any T, any oversample_rate = M/rr):  thread per (o, j).
// Reference: gr_pfb_channelizer_ccf.cc:169-196.
struct PfbFirGenArgs {
  const float2* x;     // [H + nrows_in][M], H = T
  float2* u;           // [nout][M]
  const float* taps;   // h[k + t*M] zero padded to T*M
  int M, T, rr;
  long nout;
};

__global__ void pfb_fir_generic_kernel(const PfbFirGenArgs a) {
  const int M = a.M, T = a.T;
  const long total = a.nout * M;
  for (long g = blockIdx.x * (long)blockDim.x + threadIdx.x; g < total; g += (long)gridDim.x * blockDim.x) {
    const long o = g / M;
    const int j = (int)(g - o * M);
    const long c = (o + 1) * (long)a.rr - 1;       // unwrapped filter phase
    const int last = (int)(c % M);
    const long n = 1 + c / M;                       // reference's n for this output
    const int filt = (last - j + M) % M;            // filter index applied to stream j
    const long nn = n - (j > last ? 1 : 0);         // &in[n] or &in[n-1]
    const int dst = M - ((j + a.rr) % M) - 1;       // d_idxlut[j]
    float2 acc = make_float2(0.f, 0.f);
    for (int t = 0; t < T; t++) {
      const float hv = __ldg(a.taps + filt + (size_t)t * M);
      const float2 xv = __ldg(a.x + (nn + T - 1 - t) * (long)M + j);
      acc.x += hv * xv.x;
      acc.y += hv * xv.y;
    }
    a.u[o * (long)M + dst] = acc;
  }
}

// gr_stream_to_streams layout change on the device: staging [stream][len] -> rows [len][M]
__global__ void transpose_streams_kernel(const float2* __restrict__ src, float2* __restrict__ dst, int M, int len) {
  __shared__ float2 tile[32][33];
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;  // bx: time, by: stream
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int s = by + r, t = bx + threadIdx.x;
    if (s < M && t < len) tile[r][threadIdx.x] = src[(size_t)s * len + t];
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int t = bx + r, s = by + threadIdx.x;
    if (s < M && t < len) dst[(size_t)t * M + s] = tile[threadIdx.x][r];
  }
}

// gr_pfb_decimator_ccf when the composite filter does not fit fir_decim_kernel's shared-memory tile (decim x taps
// too large): one warp per output, lanes stride over the window, so the loads of the window and of the composite taps
// are coalesced; the taps_per_filter-fold overlap of consecutive windows is served by L1/L2.
//   out[i] = sum_{n < L} g[n] * x[i * decim + n]     (g = composite complex taps in window order)
__global__ void __launch_bounds__(256) pfb_decim_warp_kernel(const float2* __restrict__ x, float2* __restrict__ out, long nout,
                                                             int decim, const float2* __restrict__ g, int L) {
  const int lane = threadIdx.x & 31;
  const long warp0 = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((long)gridDim.x * blockDim.x) >> 5;
  for (long i = warp0; i < nout; i += nwarps) {
    const float2* w = x + (size_t)i * decim;
    float ar = 0.f, ai = 0.f;
    for (int n = lane; n < L; n += 32) {
      const float2 v = __ldg(w + n), t = __ldg(g + n);
      ar = fmaf(t.x, v.x, fmaf(-t.y, v.y, ar));
      ai = fmaf(t.x, v.y, fmaf(t.y, v.x, ai));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      ar += __shfl_xor_sync(0xffffffffu, ar, o);
      ai += __shfl_xor_sync(0xffffffffu, ai, o);
    }
    if (lane == 0) out[i] = make_float2(ar, ai);
  }
}

}  // namespace grb
