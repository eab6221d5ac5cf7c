// Frequency-domain FIR for gr_fft_filter_ccc (reference:
// gnuradio-core/src/lib/filter/gri_fft_filter_ccc_generic.cc:62-165, gr_fft_filter_ccc.cc:82-103).
//
// The reference convolves block by block: forward FFT of nsamples new items zero padded to fftsize, product with the
// transformed taps, inverse FFT, add the previous block's tail (overlap-ADD), keep every decimation-th item.  The same
// linear convolution on the GPU is ONE kernel per call and one CTA per block, overlap-SAVE: the CTA reads the fftsize
// consecutive items [b*nsamples - (ntaps-1), b*nsamples + nsamples) (the plan keeps the last ntaps-1 items of the
// previous call in front of the new ones, where the reference keeps its tail), transforms them in shared memory,
// multiplies by H = FFT(taps)/fftsize, transforms back and stores the last nsamples items -- the first ntaps-1 are
// the circular wrap-around and are dropped.  Blocks are independent (no tail to hand from block to block), HBM sees
// fftsize/nsamples (about 2) x 8 B in and 8/decimation B out per item; nothing else leaves the SM.
//
// FFT: Stockham autosort in place (every thread holds its 16 points in registers across the barrier), radix 16 passes
// plus one radix 2/4/8 pass for the remainder (fftsize = 2 * 2^ceil(log2 ntaps) is a power of two,
// gri_fft_filter_ccc_generic.cc:106), butterflies of fft_radix.cuh.  Shared-memory rows are padded one item in 16:
// the first pass scatters with stride R.
#pragma once
#include "fft_radix.cuh"

namespace grb {

#define FFTF_MAX_PASSES 5
#define FFTF_ELEMS 16

struct FftFiltArgs {
  const float2* x;    // block b reads x[b * nsamples + i], i < n (items at or beyond x_limit read as zero)
  float2* out;        // total_items / decim items
  const float2* H;    // [n] FFT(taps of this partition) / n, natural order
  const float2* tw;   // forward twiddles e^{-2 pi i k / (Ns R)}, k < Ns, pass after pass (the inverse conjugates them)
  int n, ntaps, nsamples, decim;   // ntaps: taps per partition (n - nsamples + 1); nsamples: hop = kept items per block
  long nblk;
  long x_limit;       // items readable behind x
  long total_items;   // input-rate items this call produces (the last block may be partial)
  int accumulate;     // 1: out += (partitions after the first)
  int npass;
  int radix[FFTF_MAX_PASSES];
  int tw_off[FFTF_MAX_PASSES];
};

struct FftFiltPass {  // what a pass needs of the arguments, by value (registers, not a pointer into local memory)
  const float2* H;
  float2* out;
  int n, ntaps, decim, accumulate;
  long total_items;
  long in_left;       // readable items from this block's first one on
};

__device__ __forceinline__ int fftf_phys(int i) { return i + (i >> 4); }

// One Stockham pass of radix R over the CTA's row.  SRC: 0 = global items (first forward pass), 1 = shared memory,
// 2 = shared memory times H (first inverse pass).  DST: 0 = shared memory, 1 = the block's kept output items.
// (__noinline__: each instantiation is a function with its own register allocation, at most 71 registers; inlined into
// one kernel body the 24 of them make ptxas spill 7 KB)
template <int R, int DIR, int SRC, int DST>
__device__ __noinline__ void fftf_pass(const FftFiltPass a, float2* __restrict__ s, const float2* __restrict__ gin,
                                       int Ns, const float2* __restrict__ tw, long item0) {
  constexpr int B = FFTF_ELEMS / R;   // butterflies per thread
  const int n = a.n, nb = n / R;
  float2 v[B][R];
#pragma unroll
  for (int b = 0; b < B; b++) {
    const int j = threadIdx.x + b * blockDim.x;
    if (j < nb) {
#pragma unroll
      for (int r = 0; r < R; r++) {
        const int i = j + r * nb;
        if (SRC == 0) v[b][r] = i < a.in_left ? __ldg(gin + i) : make_float2(0.f, 0.f);
        else if (SRC == 1) v[b][r] = s[fftf_phys(i)];
        else v[b][r] = cmul(s[fftf_phys(i)], __ldg(a.H + i));
      }
    }
  }
  if (SRC != 0) __syncthreads();   // in place: every read precedes every write
#pragma unroll
  for (int b = 0; b < B; b++) {
    const int j = threadIdx.x + b * blockDim.x;
    if (j < nb) {
      const int k = j % Ns;
      if (Ns > 1) {
        float2 w = __ldg(tw + k);
        if (DIR > 0) w.y = -w.y;
        apply_twiddle_powers<R>(v[b], w);
      }
      butterfly<R, DIR>(v[b]);
      const int o0 = (j - k) * R + k;
#pragma unroll
      for (int r = 0; r < R; r++) {
        const int o = o0 + r * Ns;
        if (DST == 0) {
          s[fftf_phys(o)] = v[b][r];
        } else if (o >= a.ntaps - 1) {        // the first ntaps-1 items are the circular wrap-around
          const long i = item0 + (o - (a.ntaps - 1));
          if (i < a.total_items && (a.decim == 1 || i % a.decim == 0)) {
            float2* dst = a.out + (a.decim == 1 ? i : i / a.decim);
            *dst = a.accumulate ? cadd(*dst, v[b][r]) : v[b][r];
          }
        }
      }
    }
  }
  if (DST == 0) __syncthreads();
}

template <int DIR, int SRC, int DST>
__device__ __forceinline__ void fftf_pass_r(int R, const FftFiltPass& a, float2* s, const float2* gin, int Ns, const float2* tw,
                                            long item0) {
  switch (R) {
    case 16: fftf_pass<16, DIR, SRC, DST>(a, s, gin, Ns, tw, item0); break;
    case 8: fftf_pass<8, DIR, SRC, DST>(a, s, gin, Ns, tw, item0); break;
    case 4: fftf_pass<4, DIR, SRC, DST>(a, s, gin, Ns, tw, item0); break;
    default: fftf_pass<2, DIR, SRC, DST>(a, s, gin, Ns, tw, item0); break;
  }
}

// grid: any (blocks are claimed round robin), block: max(32, n/16) threads, dynamic shared memory: (n + n/16 + 1) float2
template <int MAXT>
__global__ void __launch_bounds__(MAXT, 1) fft_filter_ols_kernel(const FftFiltArgs a) {
  extern __shared__ float2 fftf_smem[];
  float2* s = fftf_smem;
  FftFiltPass pa;
  pa.H = a.H; pa.out = a.out; pa.n = a.n; pa.ntaps = a.ntaps; pa.decim = a.decim; pa.accumulate = a.accumulate;
  pa.total_items = a.total_items;
  for (long b = blockIdx.x; b < a.nblk; b += gridDim.x) {
    const float2* gin = a.x + b * (long)a.nsamples;
    pa.in_left = a.x_limit - b * (long)a.nsamples;
    const long item0 = b * (long)a.nsamples;
    int Ns = 1;
    // forward
    for (int p = 0; p < a.npass; p++) {
      const float2* tw = a.tw + a.tw_off[p];
      if (p == 0) fftf_pass_r<-1, 0, 0>(a.radix[p], pa, s, gin, Ns, tw, item0);
      else fftf_pass_r<-1, 1, 0>(a.radix[p], pa, s, gin, Ns, tw, item0);
      Ns *= a.radix[p];
    }
    // product with H on the way into the inverse; the last inverse pass stores the kept items
    Ns = 1;
    for (int p = 0; p < a.npass; p++) {
      const float2* tw = a.tw + a.tw_off[p];
      const bool first = p == 0, last = p == a.npass - 1;
      if (first && last) fftf_pass_r<+1, 2, 1>(a.radix[p], pa, s, gin, Ns, tw, item0);
      else if (first) fftf_pass_r<+1, 2, 0>(a.radix[p], pa, s, gin, Ns, tw, item0);
      else if (last) fftf_pass_r<+1, 1, 1>(a.radix[p], pa, s, gin, Ns, tw, item0);
      else fftf_pass_r<+1, 1, 0>(a.radix[p], pa, s, gin, Ns, tw, item0);
      Ns *= a.radix[p];
    }
    __syncthreads();   // the next block's first pass writes the row this block's last pass has just read
  }
}

}  // namespace grb
