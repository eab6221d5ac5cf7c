// Fused front half of the per-channel demod tail: gr_quadrature_demod_cf -> gr_fir_filter_fff (RRC),
// batched over all channels in [time][channel] layout, bit exact against the reference's
// x86-64 path (gr_fast_atan2f table arctangent + float_dotprod_sse64.S summation order).
//   reads  Y [hist + nrows][M] complex  (8 B/sample)      -- the discriminator output never
//   writes F [nrows][M] float           (4 B/sample)         touches HBM (it lives in a smem tile)
//
// A CTA owns 32 neighbouring channels x DF_RT consecutive rows.  Phase 1 computes the discriminator
// for the tile (+ the FIR history rows above it) into shared memory; phase 2 runs the FIR with a
// lane = a channel and each thread producing groups of four consecutive outputs a0..a0+3
// (a0 = 0 mod 4 in ABSOLUTE stream index), so that the four outputs share every loaded sample.
//
// SSE order restated per output a (see gr_math.cuh dot_sse): the window starts at s = a-(ntaps-1);
// samples are grouped in ABSOLUTE aligned blocks of four (lane = absolute index mod 4); the output's
// first nb%4 blocks accumulate into xmm4, the remaining blocks rotate over xmm4..xmm7; the result
// is lanes (d0+d2)+(d1+d3) of (acc0+acc1)+(acc3+acc2).  With ntaps-1 = 4q+rho the four outputs of
// a group fall in two classes (r < rho starts one block earlier) and all of that bookkeeping is a
// function of (rho, q mod 4) only: the kernel is instantiated for the 16 combinations so that
// every accumulator index is a compile-time register name.
#pragma once
#include <cuda_runtime.h>
#include "gr_math.cuh"

namespace grb {

#define DF_RT 128        // output rows per CTA tile (multiple of 4)
#define DF_THREADS 256
#define DF_MAXB 34       // max aligned blocks per output window -> ntaps <= 4*DF_MAXB - 7

struct DemodFrontArgs {
  const float2* y;       // row 0 of the buffer is absolute row (abs_row0 - hist)
  float* f;              // [nrows][M], row 0 = absolute row abs_row0
  long abs_row0;
  int nrows, M, hist;
  float gain;
  const float* atan_table;
  int ntaps, q;          // ntaps - 1 = 4*q + rho
  float tp[4][DF_MAXB * 4];  // tp[al][p] = rt[p - al] (reversed taps shifted by al, zero padded):
                             // the reference's four pre-aligned tap copies (gr_fir_fff_simd.cc:69-94)
};

template <int RHO, int QM>
__global__ void __launch_bounds__(DF_THREADS, 2) demod_front_kernel(const DemodFrontArgs a) {
  extern __shared__ float df_smem[];  // d tile [drows][32] then atan table [257]
  constexpr int DELTA_B = RHO > 0 ? 1 : 0;                 // class B (r >= RHO) starts one union block later
  const int q = a.q;
  const int J = q + 1 + DELTA_B;                            // union blocks per group of four outputs
  const int drows = DF_RT + 4 * (J - 1);                    // d rows the tile touches
  float* dtile = df_smem;
  float* tab = df_smem + (size_t)drows * 32;
  for (int i = threadIdx.x; i < 257; i += DF_THREADS) tab[i] = a.atan_table[i];

  const int c0 = blockIdx.x * 32;
  const long A0 = (a.abs_row0 >> 2) << 2;                   // tiles start on absolute multiples of 4
  const long tile_start = A0 + (long)blockIdx.y * DF_RT;
  const long d_row0 = tile_start - 4L * (J - 1);            // absolute row of dtile row 0
  const long ybase = a.abs_row0 - a.hist;                   // absolute row of y row 0
  const long yrows = (long)a.hist + a.nrows;
  __syncthreads();

  // ---- phase 1: discriminator into shared memory (gr_quadrature_demod_cf.cc:56-59) -----------
  {
    const int ch = threadIdx.x & 31;
    const int c = c0 + ch;
    for (int r = threadIdx.x >> 5; r < drows; r += DF_THREADS / 32) {
      const long p = d_row0 + r;              // absolute row
      const long yi = p - ybase;              // buffer row of Y[p]
      float d = 0.f;
      if (c < a.M && yi >= 0 && yi < yrows) {
        const float2 cur = __ldg(a.y + yi * a.M + c);
        float2 prev = make_float2(0.f, 0.f);  // before the buffer = before the stream: zeros
        if (yi >= 1) prev = __ldg(a.y + (yi - 1) * a.M + c);
        d = quad_demod(cur, prev, a.gain, tab);
      }
      dtile[r * 32 + ch] = d;
    }
  }
  __syncthreads();

  // ---- phase 2: RRC FIR, four outputs per thread per step --------------------------------------
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = c0 + lane;
  for (int g = warp; g < DF_RT / 4; g += DF_THREADS / 32) {
    const long a0 = tile_start + 4L * g;
    float acc[4][4][4];  // [output r][slot][lane]
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
      for (int s = 0; s < 4; s++)
#pragma unroll
        for (int l = 0; l < 4; l++) acc[r][s][l] = 0.f;
    const float* dcol = dtile + (size_t)(4 * g) * 32 + lane;  // union block j lane l -> dcol[(4j+l)*32]

    // one union block for all four outputs; SLOTJ = j for the first four blocks (prologue rules), -1 after
#define DF_BLOCK(j_, SLOT_OF)                                                              \
    {                                                                                      \
      float x[4];                                                                          \
      _Pragma("unroll") for (int l = 0; l < 4; l++) x[l] = dcol[(4 * (j_) + l) * 32];      \
      _Pragma("unroll") for (int r = 0; r < 4; r++) {                                      \
        /* folded to constants once the r loop is unrolled */                              \
        const int delta = (RHO > 0 && r >= RHO) ? 1 : 0;                                   \
        const int nbm = (RHO > 0 && r < RHO) ? ((QM + 2) & 3) : ((QM + 1) & 3);            \
        const int P = delta + nbm;                                                         \
        const int al = ((r - RHO) % 4 + 4) % 4;                                            \
        const int slot = SLOT_OF;                                                          \
        if ((j_) >= delta) {                                                               \
          const float* t = a.tp[al] + 4 * ((j_) - delta);                                  \
          _Pragma("unroll") for (int l = 0; l < 4; l++)                                    \
            acc[r][slot][l] = GR_FADD(acc[r][slot][l], GR_FMUL(t[l], x[l]));               \
        }                                                                                  \
      }                                                                                    \
    }
    // first four union blocks: slot = (j < P) ? P&3 : j&3, all compile-time
#define DF_SLOT_PRO(jc) (((jc) < P) ? (P & 3) : ((jc) & 3))
    if (J > 0) DF_BLOCK(0, DF_SLOT_PRO(0))
    if (J > 1) DF_BLOCK(1, DF_SLOT_PRO(1))
    if (J > 2) DF_BLOCK(2, DF_SLOT_PRO(2))
    if (J > 3) DF_BLOCK(3, DF_SLOT_PRO(3))
    int j = 4;
    for (; j + 4 <= J; j += 4) {
      DF_BLOCK(j + 0, 0)
      DF_BLOCK(j + 1, 1)
      DF_BLOCK(j + 2, 2)
      DF_BLOCK(j + 3, 3)
    }
    if (j < J) { DF_BLOCK(j, 0) j++; }
    if (j < J) { DF_BLOCK(j, 1) j++; }
    if (j < J) { DF_BLOCK(j, 2) j++; }
#undef DF_BLOCK
#undef DF_SLOT_PRO
    // combine: true accumulator a lives in slot (a + P) & 3
#pragma unroll
    for (int r = 0; r < 4; r++) {
      const int delta = (RHO > 0 && r >= RHO) ? 1 : 0;
      const int nbm = (RHO > 0 && r < RHO) ? ((QM + 2) & 3) : ((QM + 1) & 3);
      const int P = delta + nbm;
      float d[4];
#pragma unroll
      for (int l = 0; l < 4; l++)
        d[l] = GR_FADD(GR_FADD(acc[r][(0 + P) & 3][l], acc[r][(1 + P) & 3][l]),
                       GR_FADD(acc[r][(3 + P) & 3][l], acc[r][(2 + P) & 3][l]));
      const float out = GR_FADD(GR_FADD(d[0], d[2]), GR_FADD(d[1], d[3]));
      const long row = a0 + r - a.abs_row0;
      if (c < a.M && row >= 0 && row < a.nrows) a.f[row * a.M + c] = out;
    }
  }
}

}  // namespace grb
