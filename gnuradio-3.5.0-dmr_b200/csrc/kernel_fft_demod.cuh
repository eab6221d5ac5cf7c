// The channelizer's M-point FFT with gr_quadrature_demod_cf fused into its last pass
// (gr_pfb_channelizer_ccf.cc:193-199 feeding gr_quadrature_demod_cf.cc:46-62 on every channel).
//
// In the chain the channelizer output Y is only ever read by the discriminator d[t][c] = gain * atan2(Y[t][c] conj Y[t-1][c]).
// The last Stockham pass holds a whole output row in registers (thread j owns channels j, j + Ns, ...); a CTA that
// transforms CONSECUTIVE rows keeps the previous row's values of its own channels in registers as well, so the
// discriminator costs no memory traffic at all: Y (8 B written + 8 B read per sample) never reaches HBM, D (4 B) does.
//
// Rows are claimed in CHUNKS of consecutive rows from an atomic counter (same reason as the plain kernel: a CTA that
// shares its SM takes fewer).  A chunk that does not start at row 0 first transforms the row in front of it without
// storing anything ("warm" row, 1 / chunk of extra work); the chunk at row 0 takes the previous block's last row from
// prev_y, and whoever transforms the last row leaves it in last_y for the next block.
// Same staging as fft_fixed_body<STAGED>: the next row's input arrives by cp.async.bulk while this one is computed.
// The discriminator is gr_math.cuh's quad_demod, bit exact on the values it is given.  Those values are this kernel's
// own transform: the compiler contracts multiply-adds of the last pass differently here than in the plain FFT kernel, so
// the two kernels' outputs differ in the last bit for most channels (both within 1e-6 of the float64 DFT; the
// reference's FFTW output is not bit-pinned either).  For parity tests y_out makes the kernel ALSO store the transform
// it computed: the demod tail is then checked bit for bit against the oracle on exactly those values.
#pragma once
#include "fft_engine.cuh"
#include "gr_math.cuh"

namespace grb {

struct FftDemodArgs {
  FftArgs f;             // in = branch-filter rows, nrows, twiddles, geometry (out, window, rotations unused)
  float* d;              // [nrows][N] discriminator output
  float2* y_out;         // nullptr, or [nrows][N]: ALSO store the transform (parity tests: the values the discriminator saw)
  const float2* prev_y;  // [N] channelizer output of the row before row 0
  float2* last_y;        // [N] receives the channelizer output of row nrows - 1
  const float* atan_table;
  float gain;
  int chunk;             // rows per claim
  int nchunks;
};

template <int DIR, int R0, int R1, int R2>
__device__ __forceinline__ void fft_demod_body(const FftDemodArgs& A) {
  const FftArgs& a = A.f;
  constexpr int N = R0 * R1 * R2;
  constexpr int PADDIV = (R0 % 2 == 0) ? R0 : 0;
  constexpr int NB = N / R2;   // butterflies of the last pass = threads
  constexpr int NS = R0 * R1;  // its stride: thread j owns outputs j + r * NS
  static_assert(NB == NS, "the last pass must give every thread the same channels for every row");
  extern __shared__ __align__(128) float2 fft_smem[];
  const int j = threadIdx.x;
  float2* srow = fft_smem;
  const size_t work = (((size_t)a.row_stride * sizeof(float2)) + 127) / 128 * 128;
  const size_t stage_bytes = (size_t)N * sizeof(float2);
  unsigned char* base = reinterpret_cast<unsigned char*>(fft_smem);
  uint64_t* full = reinterpret_cast<uint64_t*>(base + work + stage_bytes);
  float* tab = reinterpret_cast<float*>(base + work + stage_bytes + 16);
  for (int i = threadIdx.x; i < 257; i += blockDim.x) tab[i] = A.atan_table[i];
  __shared__ long s_code[2];   // row * 4 + (warm ? 1 : 0) + (take prev_y ? 2 : 0); -1: no more rows

  // thread 0 walks the CTA's row sequence
  long it_r = -1, it_e = -1;
  int stat = blockIdx.x;
  auto advance = [&]() -> long {
    if (it_r + 1 < it_e) { it_r++; return it_r * 4; }
    int c;
    if (a.counter) c = atomicAdd(a.counter, 1);
    else { c = stat; stat += gridDim.x; }
    if (c >= A.nchunks) { it_e = -1; it_r = -1; return -1; }
    it_e = min((long)(c + 1) * A.chunk, a.nrows);
    if (c == 0) { it_r = 0; return 2; }
    it_r = (long)c * A.chunk - 1;
    return it_r * 4 + 1;
  };
  auto issue = [&](long row) {
    mbar_expect_tx(full, (unsigned)stage_bytes);
    bulk_g2s(base + work, a.in + row * (long)N, (unsigned)stage_bytes, full);
  };
  if (threadIdx.x == 0) {
    mbar_init(full, 1);
    mbar_init_fence();
    const long c0 = advance();
    if (c0 >= 0) issue(c0 >> 2);
    s_code[0] = c0;
    s_code[1] = c0 >= 0 ? advance() : -1;
  }
  __syncthreads();
  long cur = s_code[0], nxt = s_code[1];
  float2 prev[R2];
#pragma unroll
  for (int r = 0; r < R2; r++) prev[r] = make_float2(0.f, 0.f);
  const bool active = j < NB;
  for (int it = 0; cur >= 0; it++) {
    long nn = -1;
    if (threadIdx.x == 0 && nxt >= 0) nn = advance();
    const long row = cur >> 2;
    const bool warm = cur & 1;
    if ((cur & 2) && active) {
#pragma unroll
      for (int r = 0; r < R2; r++) prev[r] = __ldg(A.prev_y + j + r * NS);
    }
    mbar_wait(full, (unsigned)it & 1u);
    const float2* st = reinterpret_cast<const float2*>(base + work);
    fft_pass<R0, DIR, true, false, PADDIV, true>(a, N, 1, j, true, row, srow, nullptr, st);  // ends with a barrier
    if (threadIdx.x == 0 && nxt >= 0) issue(nxt >> 2);
    fft_pass<R1, DIR, false, false, PADDIV>(a, N, R0, j, true, row, srow, a.tw[1]);
    // last pass (fft_pass<R2, DIR, false, true>) with the discriminator where the store was
    if (active) {
      float2 v[R2];
      if (PADDIV && (NB % (PADDIV ? PADDIV : 1)) == 0) {
        const float2* __restrict__ sp = srow + fft_phys_c<PADDIV>(j);
        const int sst = NB + NB / (PADDIV ? PADDIV : 1);
#pragma unroll
        for (int r = 0; r < R2; r++) v[r] = sp[r * sst];
      } else {
#pragma unroll
        for (int r = 0; r < R2; r++) v[r] = srow[fft_phys_c<PADDIV>(j + r * NB)];
      }
      apply_twiddle_powers<R2>(v, __ldg(a.tw[2] + j));   // k = j mod NS = j
      butterfly<R2, DIR>(v);
      if (!warm) {
        float* __restrict__ drow = A.d + row * (long)N + j;
#pragma unroll
        for (int r = 0; r < R2; r++) drow[r * NS] = quad_demod(v[r], prev[r], A.gain, tab);
        if (A.y_out) {
          float2* __restrict__ yrow = A.y_out + row * (long)N + j;
#pragma unroll
          for (int r = 0; r < R2; r++) yrow[r * NS] = v[r];
        }
      }
#pragma unroll
      for (int r = 0; r < R2; r++) prev[r] = v[r];
      if (row == a.nrows - 1) {
#pragma unroll
        for (int r = 0; r < R2; r++) A.last_y[j + r * NS] = v[r];
      }
    }
    if (threadIdx.x == 0) { s_code[0] = nxt; s_code[1] = nn; }
    __syncthreads();  // publishes the sequence; also: the next row's first pass overwrites the work row read above
    cur = s_code[0];
    nxt = s_code[1];
  }
}

template <int DIR, int R0, int R1, int R2, int NREG>
__global__ void __maxnreg__(NREG) fft_demod_kernel(const FftDemodArgs A) {
  fft_demod_body<DIR, R0, R1, R2>(A);
}

}  // namespace grb
