// Batched shared-memory FFT engine (device side).  Replaces gri_fft_complex::execute +
// the window / fftshift loops of gr_fft_vcc_fftw::work
// (gnuradio-core/src/lib/general/gri_fft.cc:142-146, gr_fft_vcc_fftw.cc:64-96) and the
// M-point despinning FFT of gr_pfb_channelizer_ccf::general_work (gr_pfb_channelizer_ccf.cc:193).
//
// Algorithm: Stockham autosort, decimation in time, up to FFT_MAX_PASSES passes of radix
// R_p in {2,3,4,5,8,10,16,20}.  Pass p (Ns = R_0*...*R_{p-1}) for butterfly j in [0, N/R):
//     k = j mod Ns;  v[r] = src[j + r*N/R] * W_{Ns*R}^{k*r};  v = DFT_R(v);
//     dst[(j-k)*R + k + r*Ns] = v[r]
// The first pass reads straight from HBM (coalesced: consecutive j), applies the window /
// ifftshift; the last pass writes straight to HBM (coalesced) with the fftshift rotation;
// only the passes in between touch shared memory, in place, one row = N (+ padding) float2.
// One thread owns one butterfly per pass, so a row lives in registers across each barrier.
//
// HBM traffic = 8 B read + 8 B written per point: the algorithmic minimum (16 B/sample).
#pragma once
#include "fft_radix.cuh"
#include "tma.cuh"

namespace grb {

#define FFT_MAX_PASSES 6

struct FftArgs {
  const float2* in;      // [nrows][N]
  float2* out;           // [nrows][N]
  long nrows;
  const float* window;   // N floats or nullptr (gr_fft_vcc_fftw.cc:68-72)
  int in_rot;            // src element = in[(i + in_rot) % N]   (ifftshift, :74-79)
  int out_rot;           // out[(o + out_rot) % N] = X[o]        (fftshift, :89-93)
  const float2* tw[FFT_MAX_PASSES];  // per pass: Ns entries e^{DIR 2 pi j k/(Ns R)}
  int n;                 // N
  int npass;
  int radix[FFT_MAX_PASSES];
  int rows_per_cta;
  int threads_per_row;
  int row_stride;        // smem float2 per row (N + padding)
  int pad_div;           // phys(i) = i + i / pad_div  (0 = no padding)
  int* counter;          // staged kernels: row groups are claimed from this counter (zeroed before the launch);
                         // nullptr = static assignment g = blockIdx.x + k * gridDim.x
};

__device__ __forceinline__ int fft_phys(int i, int pad_div) { return pad_div ? i + i / pad_div : i; }

template <int PADDIV> __device__ __forceinline__ int fft_phys_c(int i) { return PADDIV ? i + i / (PADDIV ? PADDIV : 1) : i; }

// ---- one pass, compile-time radix; N / Ns may be compile-time (fixed kernels) or runtime ------
template <int R, int DIR, bool FROM_GLOBAL, bool TO_GLOBAL, int PADDIV, bool STAGED = false>
__device__ __forceinline__ void fft_pass(const FftArgs& a, int N, int Ns, int j, bool row_ok, long row,
                                         float2* __restrict__ srow, const float2* __restrict__ tw,
                                         const float2* __restrict__ staged_row = nullptr) {
  const int nb = N / R;
  const bool active = row_ok && j < nb;
  float2 v[R];
  if (active) {
    if (FROM_GLOBAL) {
      // STAGED: the row was brought into shared memory by a bulk copy while the previous row was computed
      const float2* __restrict__ g = STAGED ? staged_row : a.in + row * (long)N;
      if (a.window == nullptr && a.in_rot == 0) {  // uniform: plain rows (the channelizer's case): base + immediates
        const float2* __restrict__ gj = g + j;
#pragma unroll
        for (int r = 0; r < R; r++) v[r] = STAGED ? gj[r * nb] : __ldg(gj + r * nb);
      } else
#pragma unroll
      for (int r = 0; r < R; r++) {
        const int i = j + r * nb;
        int gi = i + a.in_rot;
        if (gi >= N) gi -= N;
        float2 x = STAGED ? g[gi] : __ldg(g + gi);
        if (a.window) {
          const float w = __ldg(a.window + i);
          x.x *= w;
          x.y *= w;
        }
        v[r] = x;
      }
    } else {
      // phys(j + r*nb) = phys(j) + r*(nb + nb/PADDIV) when PADDIV divides nb: one address register + immediates
      if (PADDIV && (nb % (PADDIV ? PADDIV : 1)) == 0) {
        const float2* __restrict__ sp = srow + fft_phys_c<PADDIV>(j);
        const int st = nb + nb / (PADDIV ? PADDIV : 1);
#pragma unroll
        for (int r = 0; r < R; r++) v[r] = sp[r * st];
      } else {
#pragma unroll
        for (int r = 0; r < R; r++) v[r] = srow[fft_phys_c<PADDIV>(j + r * nb)];
      }
    }
  }
  if (!FROM_GLOBAL && !TO_GLOBAL) __syncthreads();  // in-place: every read precedes every write
  if (active) {
    const int k = j % Ns;
    if (Ns > 1) apply_twiddle_powers<R>(v, __ldg(tw + k));
    butterfly<R, DIR>(v);
    const int o0 = (j - k) * R + k;
    if (TO_GLOBAL) {
      float2* __restrict__ g = a.out + row * (long)N;
      if (a.out_rot == 0) {  // uniform
        float2* __restrict__ go = g + o0;
#pragma unroll
        for (int r = 0; r < R; r++) go[r * Ns] = v[r];
      } else
#pragma unroll
      for (int r = 0; r < R; r++) {
        int o = o0 + r * Ns + a.out_rot;
        if (o >= N) o -= N;
        g[o] = v[r];
      }
    } else {
      if (PADDIV && (Ns % (PADDIV ? PADDIV : 1)) == 0) {  // same strength reduction on the way out
        float2* __restrict__ sp = srow + fft_phys_c<PADDIV>(o0);
        const int st = Ns + Ns / (PADDIV ? PADDIV : 1);
#pragma unroll
        for (int r = 0; r < R; r++) sp[r * st] = v[r];
      } else {
#pragma unroll
        for (int r = 0; r < R; r++) srow[fft_phys_c<PADDIV>(o0 + r * Ns)] = v[r];
      }
    }
  }
  if (!TO_GLOBAL) __syncthreads();
}

// ---- fixed plans: N = R0*R1*R2*R3 (trailing radices may be 1) ----------------------------------
template <int R0, int R1, int R2, int R3> struct FftFixedCfg {
  static constexpr int N = R0 * R1 * R2 * R3;
  static constexpr int NP = (R3 > 1) ? 4 : ((R2 > 1) ? 3 : ((R1 > 1) ? 2 : 1));
  static constexpr int cmax(int a, int b) { return a > b ? a : b; }
  static constexpr int TPR =
      cmax(cmax(N / R0, R1 > 1 ? N / R1 : 1), cmax(R2 > 1 ? N / R2 : 1, R3 > 1 ? N / R3 : 1));
  static constexpr int ROWS = cmax(1, 256 / TPR);
  static constexpr int THREADS = ROWS * TPR;
};

// MINB = resident CTAs per SM the register allocation must allow (occupancy vs. spills trade-off,
// chosen per plan from measurements; see profiles/).
// STAGED: persistent CTAs prefetch their input rows into shared memory with cp.async.bulk (TMA): the
// load of group g + gridDim.x is issued as soon as the first pass has pulled group g out of the
// stage buffer, and lands while the remaining passes and the HBM write of group g run, so the
// HBM read, the butterflies and the HBM write of successive rows overlap even at one CTA per SM
// (an 8000-point row needs 64 KB of work space and ~100 registers x 400 threads).  One stage
// buffer only: work + stage = 131 KB, which leaves room for a CTA of the (long, latency bound)
// clock-recovery kernel of the previous block on the same SM.
template <int DIR, int R0, int R1, int R2, int R3, bool STAGED>
__device__ __forceinline__ void fft_fixed_body(const FftArgs& a) {
  constexpr int N = R0 * R1 * R2 * R3;
  constexpr int NP = FftFixedCfg<R0, R1, R2, R3>::NP;
  constexpr int PADDIV = (NP > 1 && (R0 % 2 == 0)) ? R0 : 0;
  extern __shared__ __align__(128) float2 fft_smem[];
  const int tpr = a.threads_per_row;
  const int lrow = threadIdx.x / tpr;
  const int j = threadIdx.x - lrow * tpr;
  float2* srow = fft_smem + (size_t)lrow * a.row_stride;
  const long ngroups = (a.nrows + a.rows_per_cta - 1) / a.rows_per_cta;
  // staged layout: [work rows][stage][mbarrier]; stage = rows_per_cta * N float2
  const size_t work = (((size_t)a.rows_per_cta * a.row_stride * sizeof(float2)) + 127) / 128 * 128;
  const size_t stage_bytes = (size_t)a.rows_per_cta * N * sizeof(float2);
  unsigned char* base = reinterpret_cast<unsigned char*>(fft_smem);
  uint64_t* full = reinterpret_cast<uint64_t*>(base + work + stage_bytes);
  auto issue = [&](long g) {  // one thread
    const long r0 = g * a.rows_per_cta;
    const unsigned bytes = (unsigned)(min((long)a.rows_per_cta, a.nrows - r0) * N * sizeof(float2));
    mbar_expect_tx(full, bytes);
    bulk_g2s(base + work, a.in + r0 * (long)N, bytes, full);
  };
  // Row groups are CLAIMED, not assigned (staged kernels): a CTA that shares its SM with a CTA of another kernel
  // (the clock-recovery kernel of the previous block, an NCCL receive spinning on its peer) or that is not resident
  // at all for a while simply takes fewer groups, instead of the whole launch waiting for it.  A CTA holds its
  // current group and the next one (whose bulk load it issues as soon as the stage buffer is free) and claims the
  // one after that at the top of an iteration, so that the atomic's latency hides under the butterflies.
  __shared__ int s_claim[2];
  int stat = blockIdx.x;  // static sequence when there is no counter
  const int ng = (int)min(ngroups, 0x7fffffffL);
  auto claim = [&]() -> int {
    if (STAGED && a.counter) return atomicAdd(a.counter, 1);
    const int v = stat;
    stat += gridDim.x;
    return v;
  };
  int g, gn;
  if (STAGED) {
    if (threadIdx.x == 0) {
      mbar_init(full, 1);
      mbar_init_fence();
      const int g0 = claim();
      if (g0 < ng) issue(g0);
      s_claim[0] = g0;
      s_claim[1] = claim();
    }
    __syncthreads();
    g = s_claim[0];
    gn = s_claim[1];
  } else {
    g = claim();
    gn = claim();
  }
  for (int it = 0; g < ng; it++) {
    int gnn = 0;
    if (!STAGED || threadIdx.x == 0) gnn = claim();
    const long row = (long)g * a.rows_per_cta + lrow;
    const bool row_ok = lrow < a.rows_per_cta && row < a.nrows;
    const float2* st = nullptr;
    if (STAGED) {
      mbar_wait(full, (unsigned)it & 1u);
      st = reinterpret_cast<const float2*>(base + work) + (size_t)lrow * N;
    }
    if (NP == 1) {
      fft_pass<R0, DIR, true, true, PADDIV, STAGED>(a, N, 1, j, row_ok, row, srow, nullptr, st);
      if (STAGED) {
        __syncthreads();  // every thread has its inputs in registers: the stage can be refilled
        if (threadIdx.x == 0 && gn < ng) issue(gn);
      }
    } else {
      fft_pass<R0, DIR, true, false, PADDIV, STAGED>(a, N, 1, j, row_ok, row, srow, nullptr, st);  // ends with a barrier
      if (STAGED && threadIdx.x == 0 && gn < ng) issue(gn);
      if (NP == 2) {
        fft_pass<R1, DIR, false, true, PADDIV>(a, N, R0, j, row_ok, row, srow, a.tw[1]);
      } else {
        fft_pass<R1, DIR, false, false, PADDIV>(a, N, R0, j, row_ok, row, srow, a.tw[1]);
        if (NP == 3) {
          fft_pass<R2, DIR, false, true, PADDIV>(a, N, R0 * R1, j, row_ok, row, srow, a.tw[2]);
        } else {
          fft_pass<R2, DIR, false, false, PADDIV>(a, N, R0 * R1, j, row_ok, row, srow, a.tw[2]);
          fft_pass<R3, DIR, false, true, PADDIV>(a, N, R0 * R1 * R2, j, row_ok, row, srow, a.tw[3]);
        }
      }
    }
    if (STAGED) {
      if (threadIdx.x == 0) { s_claim[0] = gn; s_claim[1] = gnn; }
      __syncthreads();  // publishes the claims; also: the next group's first pass overwrites the rows read above
      g = s_claim[0];
      gn = s_claim[1];
    } else {
      if (NP != 1) __syncthreads();  // the next group's first pass overwrites the rows read above
      g = gn;
      gn = gnn;
    }
  }
}

// Two entry points over the same body: MINB = resident CTAs per SM the register allocation must allow, or an
// explicit register cap NREG (chosen so that a CTA co-resides with the clock-recovery kernel of the
// previous block: 13 warps x 128 registers leave no room for it).
template <int DIR, int R0, int R1, int R2, int R3, int MINB, bool STAGED>
__global__ void __launch_bounds__(FftFixedCfg<R0, R1, R2, R3>::THREADS, MINB) fft_fixed_kernel(const FftArgs a) {
  fft_fixed_body<DIR, R0, R1, R2, R3, STAGED>(a);
}
template <int DIR, int R0, int R1, int R2, int R3, int NREG, bool STAGED>
__global__ void __maxnreg__(NREG) fft_fixed_kernel_r(const FftArgs a) {
  fft_fixed_body<DIR, R0, R1, R2, R3, STAGED>(a);
}

// ---- generic plan: runtime radices, ping-pong rows, any thread count ---------------------------
template <int R, int DIR>
__device__ __forceinline__ void fft_generic_pass(const FftArgs& a, int p, int Ns, long row, const float2* src,
                                                 float2* dst, bool from_global, bool to_global, int j0, int jstep) {
  const int N = a.n, nb = N / R;
  for (int j = j0; j < nb; j += jstep) {
    float2 v[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
      const int i = j + r * nb;
      if (from_global) {
        int gi = i + a.in_rot;
        if (gi >= N) gi -= N;
        float2 x = __ldg(a.in + row * (long)N + gi);
        if (a.window) { const float w = __ldg(a.window + i); x.x *= w; x.y *= w; }
        v[r] = x;
      } else {
        v[r] = src[i];
      }
    }
    const int k = j % Ns;
    if (Ns > 1) apply_twiddle_powers<R>(v, __ldg(a.tw[p] + k));
    butterfly<R, DIR>(v);
    const int o0 = (j - k) * R + k;
#pragma unroll
    for (int r = 0; r < R; r++) {
      if (to_global) {
        int o = o0 + r * Ns + a.out_rot;
        if (o >= N) o -= N;
        a.out[row * (long)N + o] = v[r];
      } else {
        dst[o0 + r * Ns] = v[r];
      }
    }
  }
}

template <int DIR>
__global__ void __launch_bounds__(256) fft_generic_kernel(const FftArgs a) {
  extern __shared__ float2 fft_smem[];
  const int tpr = a.threads_per_row;
  const int lrow = threadIdx.x / tpr;
  const int j0 = threadIdx.x - lrow * tpr;
  float2* buf0 = fft_smem + (size_t)lrow * 2 * a.n;
  float2* buf1 = buf0 + a.n;
  const long ngroups = (a.nrows + a.rows_per_cta - 1) / a.rows_per_cta;
  for (long g = blockIdx.x; g < ngroups; g += gridDim.x) {
    const long row = g * a.rows_per_cta + lrow;
    const bool row_ok = lrow < a.rows_per_cta && row < a.nrows;
    int Ns = 1;
    float2* src = buf0;
    float2* dst = buf1;
    for (int p = 0; p < a.npass; p++) {
      const bool fg = (p == 0), tg = (p == a.npass - 1);
      if (row_ok) {
        switch (a.radix[p]) {
          case 2: fft_generic_pass<2, DIR>(a, p, Ns, row, src, dst, fg, tg, j0, tpr); break;
          case 3: fft_generic_pass<3, DIR>(a, p, Ns, row, src, dst, fg, tg, j0, tpr); break;
          case 4: fft_generic_pass<4, DIR>(a, p, Ns, row, src, dst, fg, tg, j0, tpr); break;
          case 5: fft_generic_pass<5, DIR>(a, p, Ns, row, src, dst, fg, tg, j0, tpr); break;
          case 8: fft_generic_pass<8, DIR>(a, p, Ns, row, src, dst, fg, tg, j0, tpr); break;
          default: break;
        }
      }
      Ns *= a.radix[p];
      __syncthreads();
      float2* t = src; src = dst; dst = t;
    }
  }
}

// ---- fallback for lengths with a prime factor > 5: direct O(N^2) DFT, one thread per output ---
// tw[0] = e^{DIR 2 pi j k / N}, k in [0, N).
template <int DIR>
__global__ void fft_naive_kernel(const FftArgs a) {
  const int N = a.n;
  for (long row = blockIdx.y; row < a.nrows; row += gridDim.y) {
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < N; k += gridDim.x * blockDim.x) {
      float2 acc = make_float2(0.f, 0.f);
      int idx = 0;
      for (int n = 0; n < N; n++) {
        int gi = n + a.in_rot;
        if (gi >= N) gi -= N;
        float2 x = __ldg(a.in + row * (long)N + gi);
        if (a.window) { const float w = __ldg(a.window + n); x.x *= w; x.y *= w; }
        const float2 w = __ldg(a.tw[0] + idx);
        acc.x += x.x * w.x - x.y * w.y;
        acc.y += x.x * w.y + x.y * w.x;
        idx += k;
        if (idx >= N) idx -= N;
      }
      int o = k + a.out_rot;
      if (o >= N) o -= N;
      a.out[row * (long)N + o] = acc;
    }
  }
}

}  // namespace grb
