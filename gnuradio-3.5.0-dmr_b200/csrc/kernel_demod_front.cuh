// Fused front half of the per-channel demod tail: gr_quadrature_demod_cf -> gr_fir_filter_fff (RRC),
// batched over all channels in [time][channel] layout, bit exact against the reference's
// x86-64 path (gr_fast_atan2f table arctangent + float_dotprod_sse64.S summation order).
//   reads  Y [hist + nrows][M] complex  (8 B/sample)      -- the discriminator output never
//   writes F [nrows][M] float           (4 B/sample)         touches HBM (it lives in a smem tile)
//
// A CTA owns 32 neighbouring channels x DF_RT consecutive rows.  Phase 1 computes the discriminator
// for the tile (+ the FIR history rows above it) into shared memory: a warp walks CONSECUTIVE rows
// of its 32 channels, so the previous sample is the register it loaded one row earlier and the
// loads of several rows are in flight together.  Phase 2 runs the FIR with a lane = a channel and
// each thread producing groups of four consecutive outputs a0..a0+3 (a0 = 0 mod 4 in ABSOLUTE
// stream index), so that the four outputs share every loaded sample; the taps come from shared
// memory as broadcast LDS.128 (one instruction per output and block of four taps).
//
// This kernel is FP32-issue bound, not HBM bound: the reference order forbids FMA contraction
// (separate IEEE multiply and add per tap: 2 x 36 instructions per output for 29 taps, + the 16-way
// accumulator tree, + ~50 for the table arctangent with a correctly rounded division), i.e.
// ~170 instructions per 12 bytes of HBM traffic against a machine balance of ~11 FP32 lanes-ops/B.
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
#include "packed_f32.cuh"

namespace grb {

#define DF_RT 256        // output rows per CTA tile (multiple of 4)
#define DF_THREADS 256  // 2 CTAs per SM at 96 registers, and still 2 next to a CTA of the clock-recovery kernel (320 x 320 tiles are 6 % faster alone, slower in the pipelined chain)
#define DF_MAXB 34       // max aligned blocks per output window -> ntaps <= 4*DF_MAXB - 7

// Discriminator tile layout: four consecutive rows of a channel are one 16-byte quad, element (row r, lane l) at
// [(r >> 2) * 128 + l * 4 + (r & 3)] -- the FIR phase takes the four samples of a union block with ONE LDS.128 (already a
// packed register pair for FMUL2), the discriminator phase stores one STS.128 per four rows.
__device__ __forceinline__ int df_dt(int r, int lane) { return ((r >> 2) << 7) + (lane << 2) + (r & 3); }

struct DemodFrontArgs {
  const float* dsrc;     // non-null: the discriminator output already exists (the FFT kernel of the channelizer made
                         // it, kernel_fft_demod.cuh): same row numbering as y, phase 1 is a plain copy into the tile
  const float2* y;       // row 0 of the buffer is absolute row (abs_row0 - hist)
  float* f;              // [nrows][M], row 0 = absolute row abs_row0
  long abs_row0;
  int nrows, M, hist;
  float gain;
  float one;             // 1.0f, opaque to the compiler (see df_acc2)
  const float* atan_table;
  int ntaps, q;          // ntaps - 1 = 4*q + rho
  const float* tp;       // [4][DF_MAXB * 4] device: tp[al][p] = rt[p - al] (reversed taps shifted by al, zero
                         // padded): the reference's four pre-aligned tap copies (gr_fir_fff_simd.cc:69-94)
  unsigned long long tpc[4][8][2];  // the same for the first 8 blocks as packed pairs IN THE KERNEL PARAMETERS: the JFIX
                         // instantiation addresses them with compile-time indices, i.e. as constant-bank operands --
                         // no shared-memory load per tap block (the FIR phase is bound by the LDS pipe, not by issue)
};

// accumulator slot of union block j (first four blocks: the reference's "first nblocks%4 blocks go to xmm4"
// prologue; then round robin), and whether block j is the first one to touch its slot
__host__ __device__ constexpr int df_slot(int j, int P) { return j < 4 ? (j < P ? (P & 3) : (j & 3)) : (j & 3); }
__host__ __device__ constexpr bool df_first(int j, int delta, int P) {
  for (int jj = delta; jj < j; jj++)
    if (df_slot(jj, P) == df_slot(j, P)) return false;
  return true;
}

// JFIX != 0: the number of union blocks is a compile-time constant (the instantiation for the chain's 29-tap matched
// filter: every loop bound, the tile height and the accumulator bookkeeping fold; same arithmetic, fewer instructions)
template <int RHO, int QM, int JFIX = 0>
__global__ void __maxnreg__(96) demod_front_kernel(const DemodFrontArgs a) {
  extern __shared__ __align__(16) float df_smem[];  // taps [4][DF_MAXB*4], atan table [260], d tile [drows][32]
  constexpr int DELTA_B = RHO > 0 ? 1 : 0;                 // class B (r >= RHO) starts one union block later
  const int q = a.q;
  const int J = JFIX ? JFIX : q + 1 + DELTA_B;              // union blocks per group of four outputs
  const int drows = DF_RT + 4 * (J - 1);                    // d rows the tile touches
  float* taps_s = df_smem;
  float* tab = taps_s + 4 * DF_MAXB * 4;
  float* dtile = tab + 260;
  for (int i = threadIdx.x; i < 4 * DF_MAXB * 4; i += DF_THREADS) taps_s[i] = a.tp[i];
  for (int i = threadIdx.x; i < 257; i += DF_THREADS) tab[i] = a.atan_table[i];

  const int c0 = blockIdx.x * 32;
  const long A0 = (a.abs_row0 >> 2) << 2;                   // tiles start on absolute multiples of 4
  const long tile_start = A0 + (long)blockIdx.y * DF_RT;
  const long d_row0 = tile_start - 4L * (J - 1);            // absolute row of dtile row 0
  const long ybase = a.abs_row0 - a.hist;                   // absolute row of y row 0
  const long yrows = (long)a.hist + a.nrows;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = c0 + lane;
  __syncthreads();

  // ---- phase 1: discriminator into shared memory (gr_quadrature_demod_cf.cc:56-59) -----------
  if (a.dsrc) {  // (kernel uniform) the discriminator ran in the channelizer's FFT kernel: fetch its rows
    constexpr int NW = DF_THREADS / 32;
    const int per = ((drows / 4 + NW - 1) / NW) * 4;        // rows per warp: whole quads (drows is a multiple of 4)
    const int r0 = min(drows, warp * per), r1 = min(drows, r0 + per);
    const bool cok = c < a.M;
    const long yi0 = d_row0 + r0 - ybase;
    const float* __restrict__ dcolumn = a.dsrc + (cok ? c : 0);
    int r = r0;
    for (; r + 8 <= r1; r += 8) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const long yi = yi0 + (r - r0) + u;
        v[u] = (cok && yi >= 0 && yi < yrows) ? __ldg(dcolumn + yi * a.M) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 8; u++) dtile[df_dt(r + u, lane)] = v[u];
    }
    for (; r < r1; r++) {
      const long yi = yi0 + (r - r0);
      dtile[df_dt(r, lane)] = (cok && yi >= 0 && yi < yrows) ? __ldg(dcolumn + yi * a.M) : 0.f;
    }
  } else {
    constexpr int NW = DF_THREADS / 32;
    const int per = ((drows / 4 + NW - 1) / NW) * 4;        // consecutive rows per warp: whole quads
    const int r0 = min(drows, warp * per), r1 = min(drows, r0 + per);
    const bool cok = c < a.M;
    const float2* __restrict__ ycol = a.y + (cok ? c : 0);
    auto ld = [&](long yi) -> float2 {                      // rows outside the buffer: before the stream / not there yet
      if (cok && yi >= 0 && yi < yrows) return __ldg(ycol + yi * a.M);
      return make_float2(0.f, 0.f);
    };
    const long yi0 = d_row0 + r0 - ybase;                   // buffer row of the first d row of this warp
    // interior tiles (all but the first / last row tile and a ragged last channel group): no bounds
    // checks, one running pointer
    if (c0 + 32 <= a.M && d_row0 - 1 >= ybase && d_row0 + drows <= ybase + yrows) {
      const size_t M = (size_t)a.M;
      const float2* __restrict__ p = a.y + (size_t)(yi0 - 1) * M + c;
      float2 prev = __ldg(p);
      p += M;
      float4* dt = reinterpret_cast<float4*>(dtile) + (r0 >> 2) * 32 + lane;   // r0 is a multiple of 4
      // Batches of DF_PB rows (three quads), the next batch's loads in flight while this one is computed: the rows come
      // straight from HBM (each is read once), and with 16 warps per SM it is their latency, not the arctangent, that
      // the phase would otherwise wait for (ncu: 2.3 long-scoreboard stalls per issue with batches of 4 and no overlap).
      constexpr int DF_PB = 12;
      float2 nx[DF_PB];
      const int nfull = (r1 - r0) / DF_PB;      // whole batches (36 rows per warp = 3 x 12 for 29 taps)
      int r = r0;
      if (nfull > 0) {
#pragma unroll
        for (int u = 0; u < DF_PB; u++) nx[u] = __ldg(p + u * M);
        for (int k = 0; k < nfull; k++) {
          float2 v[DF_PB];
#pragma unroll
          for (int u = 0; u < DF_PB; u++) v[u] = nx[u];
          p += DF_PB * M;
          if (k + 1 < nfull) {
#pragma unroll
            for (int u = 0; u < DF_PB; u++) nx[u] = __ldg(p + u * M);
          }
#pragma unroll
          for (int qd = 0; qd < DF_PB / 4; qd++) {
            float4 d4;
            d4.x = quad_demod(v[4 * qd + 0], prev, a.gain, tab);
            d4.y = quad_demod(v[4 * qd + 1], v[4 * qd + 0], a.gain, tab);
            d4.z = quad_demod(v[4 * qd + 2], v[4 * qd + 1], a.gain, tab);
            d4.w = quad_demod(v[4 * qd + 3], v[4 * qd + 2], a.gain, tab);
            prev = v[4 * qd + 3];
            dt[qd * 32] = d4;
          }
          dt += (DF_PB / 4) * 32;
        }
        r += nfull * DF_PB;
      }
      for (; r < r1; r += 4) {                  // remaining quads of the warp's share
        float2 v[4];
#pragma unroll
        for (int u = 0; u < 4; u++) v[u] = __ldg(p + u * M);
        p += 4 * M;
        float4 d4;
        d4.x = quad_demod(v[0], prev, a.gain, tab);
        d4.y = quad_demod(v[1], v[0], a.gain, tab);
        d4.z = quad_demod(v[2], v[1], a.gain, tab);
        d4.w = quad_demod(v[3], v[2], a.gain, tab);
        prev = v[3];
        *dt = d4;
        dt += 32;
      }
    } else {
    float2 prev = ld(yi0 - 1);
    int r = r0;
    for (; r + 4 <= r1; r += 4) {
      float2 v[4];
#pragma unroll
      for (int u = 0; u < 4; u++) v[u] = ld(yi0 + (r - r0) + u);
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const long yi = yi0 + (r - r0) + u;
        // d = 0 where the reference has not produced a sample (before the stream start the history is
        // zeros, gr_buffer.cc:201-214; rows past the end of the buffer only ever meet zero taps)
        const float d = (cok && yi >= 0 && yi < yrows) ? quad_demod(v[u], prev, a.gain, tab) : 0.f;
        dtile[df_dt(r + u, lane)] = d;
        prev = v[u];
      }
    }
    for (; r < r1; r++) {
      const long yi = yi0 + (r - r0);
      const float2 cur = ld(yi);
      dtile[df_dt(r, lane)] = (cok && yi >= 0 && yi < yrows) ? quad_demod(cur, prev, a.gain, tab) : 0.f;
      prev = cur;
    }
    }
  }
  __syncthreads();

  // ---- phase 2: RRC FIR, four outputs per thread per step --------------------------------------
  const ulonglong2* tp2 = reinterpret_cast<const ulonglong2*>(taps_s);  // tp2[al * DF_MAXB + b] = taps of block b for alignment al
  const df_u64 ones = df_pack(a.one, a.one);
  const int rel0 = (int)(tile_start - a.abs_row0);          // output row of the tile's first row (negative in the first tile)
  for (int g = warp; g < DF_RT / 4; g += DF_THREADS / 32) {
    const int rel = rel0 + 4 * g;
    if (rel + 3 < 0 || rel >= a.nrows) continue;            // warp uniform
    df_u64 acc[4][4][2];  // [output r][slot][lane pair]
    const float4* dcol = reinterpret_cast<const float4*>(dtile) + g * 32 + lane;  // union block j = the quad dcol[j * 32]
    int j;
    // one union block for all four outputs.  SLOT_OF = accumulator slot of this block, FIRST_OF = the block is the
    // first to touch that slot (its products initialise the accumulator); both fold to constants once the r loop
    // is unrolled
#define DF_BLOCK(j_, SLOT_OF, FIRST_OF)                                                    \
    {                                                                                      \
      const float4 xq = dcol[(j_) * 32];                                                   \
      const df_u64 x01 = df_pack(xq.x, xq.y), x23 = df_pack(xq.z, xq.w);                   \
      _Pragma("unroll") for (int r = 0; r < 4; r++) {                                      \
        const int delta = (RHO > 0 && r >= RHO) ? 1 : 0;                                   \
        const int nbm = (RHO > 0 && r < RHO) ? ((QM + 2) & 3) : ((QM + 1) & 3);            \
        const int P = delta + nbm;                                                         \
        const int al = ((r - RHO) % 4 + 4) % 4;                                            \
        const int slot = SLOT_OF;                                                          \
        (void)P;                                                                           \
        if ((j_) >= delta) {                                                               \
          ulonglong2 t;                                                                    \
          if (JFIX) t = make_ulonglong2(a.tpc[al][((j_) - delta) & 7][0], a.tpc[al][((j_) - delta) & 7][1]); \
          else t = tp2[al * DF_MAXB + ((j_) - delta)];                                     \
          const df_u64 p01 = df_mul2(t.x, x01), p23 = df_mul2(t.y, x23);                   \
          if (FIRST_OF) {                                                                  \
            acc[r][slot][0] = p01;                                                         \
            acc[r][slot][1] = p23;                                                         \
          } else {                                                                         \
            acc[r][slot][0] = df_acc2(p01, ones, acc[r][slot][0]);                         \
            acc[r][slot][1] = df_acc2(p23, ones, acc[r][slot][1]);                         \
          }                                                                                \
        }                                                                                  \
      }                                                                                    \
    }
    if (J >= 8) {
      // Every slot is first touched by one of the union blocks 0..7 at a compile-time known place: that
      // product initialises the accumulator instead of being added to zero.  0 + p and p differ only for
      // p = -0 (the reference's accumulators are never -0), which can only change the sign of an all-zero
      // result: the `+ 0.0f` at the end restores it.  Saves the zeroings and 64 of the additions.
#define DF_BLOCK8(j_) DF_BLOCK(j_, df_slot((j_), P), df_first((j_), delta, P))
      DF_BLOCK8(0) DF_BLOCK8(1) DF_BLOCK8(2) DF_BLOCK8(3) DF_BLOCK8(4) DF_BLOCK8(5) DF_BLOCK8(6) DF_BLOCK8(7)
#undef DF_BLOCK8
      j = 8;
    } else {
#pragma unroll
      for (int r = 0; r < 4; r++)
#pragma unroll
        for (int s = 0; s < 4; s++) acc[r][s][0] = acc[r][s][1] = 0ull;
      j = 0;
    }
    // first four union blocks: slot = (j < P) ? P&3 : j&3, all compile-time
#define DF_SLOT_PRO(jc) (((jc) < P) ? (P & 3) : ((jc) & 3))
    if (j == 0) {
      if (J > 0) DF_BLOCK(0, DF_SLOT_PRO(0), false)
      if (J > 1) DF_BLOCK(1, DF_SLOT_PRO(1), false)
      if (J > 2) DF_BLOCK(2, DF_SLOT_PRO(2), false)
      if (J > 3) DF_BLOCK(3, DF_SLOT_PRO(3), false)
      j = 4;
    }
    for (; j + 4 <= J; j += 4) {
      DF_BLOCK(j + 0, 0, false)
      DF_BLOCK(j + 1, 1, false)
      DF_BLOCK(j + 2, 2, false)
      DF_BLOCK(j + 3, 3, false)
    }
    if (j < J) { DF_BLOCK(j, 0, false) j++; }
    if (j < J) { DF_BLOCK(j, 1, false) j++; }
    if (j < J) { DF_BLOCK(j, 2, false) j++; }
#undef DF_BLOCK
#undef DF_SLOT_PRO
    // combine: true accumulator a lives in slot (a + P) & 3
    float outs[4];
#pragma unroll
    for (int r = 0; r < 4; r++) {
      const int delta = (RHO > 0 && r >= RHO) ? 1 : 0;
      const int nbm = (RHO > 0 && r < RHO) ? ((QM + 2) & 3) : ((QM + 1) & 3);
      const int P = delta + nbm;
      df_u64 d[2];  // lanes (0, 1) and (2, 3) of (acc0 + acc1) + (acc3 + acc2)
#pragma unroll
      for (int h = 0; h < 2; h++)
        d[h] = df_add2(df_add2(acc[r][(0 + P) & 3][h], acc[r][(1 + P) & 3][h]),
                       df_add2(acc[r][(3 + P) & 3][h], acc[r][(2 + P) & 3][h]));
      float e0, e1;
      df_unpack(df_add2(d[0], d[1]), e0, e1);  // (d0 + d2, d1 + d3)
      outs[r] = GR_FADD(GR_FADD(e0, e1), 0.0f);  // (-0) + 0 = +0, else unchanged
    }
    if (c < a.M) {
      float* __restrict__ fp = a.f + (long)rel * a.M + c;
      if (rel >= 0 && rel + 3 < a.nrows) {       // all four rows inside the block: one address, three increments
        fp[0] = outs[0];
        fp[a.M] = outs[1];
        fp[2 * (long)a.M] = outs[2];
        fp[3 * (long)a.M] = outs[3];
      } else {
#pragma unroll
        for (int r = 0; r < 4; r++)
          if (rel + r >= 0 && rel + r < a.nrows) fp[r * (long)a.M] = outs[r];
      }
    }
  }
}

}  // namespace grb
