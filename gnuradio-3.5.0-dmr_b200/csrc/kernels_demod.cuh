// Batched per-channel demod tail on the channelizer output, layout [time][channel] so that a
// warp's 32 lanes are 32 neighbouring channels at the same time step (coalesced 128/256 B rows).
// Every kernel reproduces the reference arithmetic bit for bit (csrc/gr_math.cuh).
#pragma once
#include <cuda_runtime.h>
#include "gr_math.cuh"

namespace grb {

// ---- gr_quadrature_demod_cf::work, batched (gr_quadrature_demod_cf.cc:46-62) -------------------
// in: [1 + nrows][nchan] complex (row 0 = history), out: [nrows][nchan].  12 B/sample of HBM.
__global__ void __launch_bounds__(256) quad_demod_kernel(const float2* __restrict__ in, float* __restrict__ out,
                                                         long nrows, int nchan, float gain,
                                                         const float* __restrict__ atan_table) {
  __shared__ float tab[257];
  for (int i = threadIdx.x; i < 257; i += blockDim.x) tab[i] = atan_table[i];
  __syncthreads();
  const long total = nrows * nchan;
  for (long g = blockIdx.x * (long)blockDim.x + threadIdx.x; g < total; g += (long)gridDim.x * blockDim.x) {
    const float2 prev = __ldg(in + g);
    const float2 cur = __ldg(in + g + nchan);
    out[g] = quad_demod(cur, prev, gain, tab);
  }
}

__global__ void fast_atan2f_kernel(const float* __restrict__ y, const float* __restrict__ x, float* __restrict__ out,
                                   long n, const float* __restrict__ atan_table) {
  __shared__ float tab[257];
  for (int i = threadIdx.x; i < 257; i += blockDim.x) tab[i] = atan_table[i];
  __syncthreads();
  for (long g = blockIdx.x * (long)blockDim.x + threadIdx.x; g < n; g += (long)gridDim.x * blockDim.x)
    out[g] = fast_atan2f(y[g], x[g], tab);
}

// ---- gr_fir_filter_fff::work, batched over channels (gr_fir_filter_XXX.cc.t:66-88) ------------
// in: [ntaps-1 + nout*decim][nchan] (history-prefixed), out: [nout][nchan]; out[o][c] =
// dot(rt, in[o*decim .. +ntaps)[c]) in the selected reference summation order.
// abs0 = absolute stream index of in row 0 (decides the SSE lane phase).
#define FFF_MAX_TAPS_SMEM 4096
__global__ void __launch_bounds__(128) fir_fff_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                      long nout, int nchan, int decim, const float* __restrict__ rt_g,
                                                      int ntaps, int order, long abs0, int rows_per_thread) {
  extern __shared__ float rt[];
  for (int i = threadIdx.x; i < ntaps; i += blockDim.x) rt[i] = rt_g[i];
  __syncthreads();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nchan) return;
  const long o0 = (long)blockIdx.y * rows_per_thread;
  const long o1 = min(nout, o0 + rows_per_thread);
  for (long o = o0; o < o1; o++) {
    const float* p = in + (o * decim) * (long)nchan + c;
    const float v = (order == GR_ORDER_SSE) ? dot_sse(rt, ntaps, p, nchan, mod4(abs0 + o * decim))
                                            : dot_generic(rt, ntaps, p, nchan);
    out[o * (long)nchan + c] = v;
  }
}

// ---- digital_clock_recovery_mm_ff::general_work, batched (digital_clock_recovery_mm_ff.cc:102-139)
// One thread per channel: the loop is sequential in time (mu/omega/last_sample feed back), but
// the channels are independent and, in [time][channel] layout, the lanes of a warp walk the
// same rows at (nearly) the same pace, so their loads share 128 B lines.  The kernel itself is in
// kernel_mm.cuh (warp specialised: loader / recursion / slicer+correlator).
struct MMChanState {  // persists across work calls (the block's members d_mu, d_omega, ...)
  float mu, omega, last_sample, slicer_avg;
  long long next_abs;  // absolute input index of the next in[ii] for this channel
  int clamped;         // times the loop tried to step before the first buffered row (see mm_kernel)
  int overflow;        // times the per-call output capacity stopped the loop before the input did
};

// per-channel registers of digital_correlate_access_code_bb (d_data_reg, d_flag_reg) + bit counter
struct CorrChanState {
  unsigned long long data_reg, flag_reg;
  long long nbits;  // absolute number of bits already processed on this channel
};
struct CorrHit { int channel; int pad; long long bit_index; };

// Optional fused epilogue of the M&M kernel: slicer decision -> gr_map_bb -> gr_unpack_k_bits_bb(k)
// -> correlator, in the same thread that produced the symbol (no second latency-bound pass over
// the symbol stream, no re-read of it from HBM).
struct MMCorrFuse {
  int on;
  int bits_per_symbol;
  unsigned char map[256];
  unsigned char* out;      // [k*max_out][nchan] correlator bytes or nullptr
  CorrChanState* state;
  CorrParams p;
  CorrHit* hits;
  int max_hits;
  int* nhits;
};

struct MMArgs {
  MMCorrFuse corr;
  const float* in;       // [ninput][nchan], row 0 has absolute index abs_row0
  long ninput;
  long abs_row0;
  int nchan;
  float* out;            // [max_out][nchan] soft symbols
  unsigned char* sliced; // [max_out][nchan] or nullptr
  int max_out;
  int* counts;           // [nchan]
  MMChanState* state;    // [nchan]
  const MMChanState* state_in;  // nullptr, or where the initial state is read instead (a time shard: the buffer
                                // the left neighbour's final state was received into)
  MMChanState* state_out2;      // nullptr, or a second place the final state is written to (the send buffer)
  MMParams p;
  int order;
  int slicer_levels;     // 0, 2 or 4
  float slicer_alpha, slicer_beta;
  const float* mmse_eff; // [129][8] coefficients applied to in[ii+0..7]
  float one;             // 1.0f, opaque to the compiler: acc + p is issued as fma(p, 1, acc) so that ptxas cannot
                         // contract a packed multiply into the addition (kernel_demod_front.cuh)
};

// stand-alone slicer (single stream; the recurrence on d_avg is sequential when alpha != 0)
__global__ void slicer_kernel(const float* __restrict__ in, unsigned char* __restrict__ out, long n, int levels,
                              float alpha, float beta, float* avg_state) {
  if (levels == 2 || alpha == 0.0f) {
    // no state: fully parallel (avg stays avg*beta + 0 = avg*1 -> constant)
    const float avg0 = *avg_state;
    for (long g = blockIdx.x * (long)blockDim.x + threadIdx.x; g < n; g += (long)gridDim.x * blockDim.x) {
      if (levels == 2) out[g] = slice2(in[g]);
      else { float avg = avg0; out[g] = slice4(in[g], avg, 0.0f, beta); }
    }
  } else if (blockIdx.x == 0 && threadIdx.x == 0) {
    float avg = *avg_state;
    for (long g = 0; g < n; g++) out[g] = slice4(in[g], avg, alpha, beta);
    *avg_state = avg;
  }
}

// ---- gr_map_bb -> gr_unpack_k_bits_bb -> digital_correlate_access_code_bb, batched ------------
// (gr_map_bb.cc:49-61, gr_unpack_k_bits_bb.cc:53-70, digital_correlate_access_code_bb.cc:87-133)
struct CorrArgs {
  const unsigned char* symbols;  // [sym_rows][nchan] (slicer decisions) or bits when bits_per_symbol == 0
  const int* counts;             // [nchan] valid symbols per channel (nullptr: fixed_count)
  int fixed_count;
  int nchan;
  unsigned char map[256];        // gr_map_bb table
  int bits_per_symbol;           // k of gr_unpack_k_bits_bb (0: input already is one bit per byte)
  unsigned char* out;            // [out_rows][nchan] or nullptr
  CorrChanState* state;
  CorrParams p;
  CorrHit* hits;
  int max_hits;
  int* nhits;
};

__global__ void __launch_bounds__(128) corr_kernel(const CorrArgs a) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= a.nchan) return;
  CorrChanState st = a.state[c];
  const int n = a.counts ? a.counts[c] : a.fixed_count;
  const int k = a.bits_per_symbol;
  long ob = 0;
  for (int s = 0; s < n; s++) {
    const unsigned sym = a.symbols[(long)s * a.nchan + c];
    if (k == 0) {
      const unsigned char t = corr_step(st.data_reg, st.flag_reg, a.p, sym);
      if (a.out) a.out[ob * a.nchan + c] = t;
      if (t & 2) {
        const int h = atomicAdd(a.nhits, 1);
        if (h < a.max_hits) { a.hits[h].channel = c; a.hits[h].pad = 0; a.hits[h].bit_index = st.nbits + ob; }
      }
      ob++;
    } else {
      const unsigned d = a.map[sym];
      for (int b = k - 1; b >= 0; b--) {  // MSB first
        const unsigned char t = corr_step(st.data_reg, st.flag_reg, a.p, (d >> b) & 1u);
        if (a.out) a.out[ob * a.nchan + c] = t;
        if (t & 2) {
          const int h = atomicAdd(a.nhits, 1);
          if (h < a.max_hits) { a.hits[h].channel = c; a.hits[h].pad = 0; a.hits[h].bit_index = st.nbits + ob; }
        }
        ob++;
      }
    }
  }
  st.nbits += ob;
  a.state[c] = st;
}

// ---- gr_map_bb -> gr_unpack_k_bits_bb(2) -> digital_correlate_access_code_bb, parallel in TIME as well ----------
// (digital_correlate_access_code_bb.cc:87-133.)  The block's output at bit i is a pure function of the 64 bits before
// it: data_reg = bits [i-64, i-1]; a match at bit i (popcount((data_reg ^ code) & mask) <= threshold) raises the
// flag that reaches bit 63 of flag_reg, i.e. the output, exactly `len` bits later.  So a thread that starts 64
// symbols (128 bits >= 64 + len) before its chunk with zeroed registers has the sequential registers bit for bit by
// the time its chunk begins, and chunks are independent: thread = (channel, chunk of CORR_CHUNK symbols).  Chunk 0
// starts from the carried registers instead.  The chunk holding a channel's last symbol writes the registers to
// state_out (a separate array: every chunk reads the old bit count); hits go to the shared list in any order.
// Requirements (else the caller uses the sequential kernel): k = 2 bits per symbol, code length >= 16, no byte output.
#define CORR_CHUNK 128
struct CorrParArgs {
  const unsigned char* symbols;  // [sym_rows][nchan] slicer decisions
  const int* counts;             // [nchan]
  int nchan;
  unsigned char map[256];
  const CorrChanState* state_in;
  CorrChanState* state_out;
  CorrParams p;
  CorrHit* hits;
  int max_hits;
  int* nhits;
};

__global__ void __launch_bounds__(128) corr_par_kernel(const CorrParArgs a) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= a.nchan) return;
  const int n = a.counts[c];
  const int k = blockIdx.y;
  const int s0 = k * CORR_CHUNK;
  if (k > 0 && s0 >= n) return;
  const int s1 = min(n, s0 + CORR_CHUNK);
  const CorrChanState old = a.state_in[c];
  unsigned long long data_reg = 0, flag_reg = 0;
  int s = 0;
  if (k == 0) { data_reg = old.data_reg; flag_reg = old.flag_reg; }
  else s = s0 - 64;
  const CorrParams cp = a.p;
  const int code_len = cp.flag_bit ? 64 - (__ffsll((long long)cp.flag_bit) - 1) : 0;
  const int flag_shift = 64 - code_len;
  const unsigned code_hi = (unsigned)(cp.access_code >> 32), code_lo = (unsigned)cp.access_code;
  const unsigned mask_hi = (unsigned)(cp.mask >> 32), mask_lo = (unsigned)cp.mask;
  const size_t nchan = (size_t)a.nchan;
  const unsigned char* sp = a.symbols + (size_t)s * nchan + c;
  const long long first_bit = 2ll * s0;  // hits at bit positions >= this one are this chunk's to report
  auto hit = [&](long long bit) {
    if (bit < first_bit) return;
    const int h = atomicAdd(a.nhits, 1);
    if (h < a.max_hits) { a.hits[h].channel = c; a.hits[h].pad = 0; a.hits[h].bit_index = old.nbits + bit; }
  };
  // eight symbols = 16 bits at a time (same window form as the fused epilogue of the clock-recovery kernel)
  for (; s + 8 <= s1; s += 8) {
    unsigned bits16 = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) bits16 |= ((unsigned)a.map[sp[(size_t)i * nchan]] & 3u) << (14 - 2 * i);
    sp += 8 * nchan;
    const unsigned dhi = (unsigned)(data_reg >> 32), dlo = (unsigned)data_reg;
    const unsigned inb = bits16 << 16;
    unsigned mm = 0;
#pragma unroll
    for (int j = 0; j < 16; j++) {
      const unsigned shi = __funnelshift_l(dlo, dhi, j);
      if (__popc((shi ^ code_hi) & mask_hi) <= cp.threshold) {
        const unsigned slo = __funnelshift_l(inb, dlo, j);
        const unsigned nwrong = __popc((shi ^ code_hi) & mask_hi) + __popc((slo ^ code_lo) & mask_lo);
        mm |= (nwrong <= cp.threshold ? 1u : 0u) << (15 - j);
      }
    }
    const unsigned hits16 = (unsigned)(flag_reg >> 48);
    if (hits16) {
      for (int j = 0; j < 16; j++)
        if (hits16 & (0x8000u >> j)) hit(2ll * s + j);
    }
    data_reg = (data_reg << 16) | bits16;
    flag_reg = (flag_reg << 16) | ((unsigned long long)mm << flag_shift);
  }
  for (; s < s1; s++) {  // fewer than eight symbols left: bit serial
    const unsigned dib = a.map[*sp];
    sp += nchan;
    for (int b = 1; b >= 0; b--) {
      const unsigned char t = corr_step(data_reg, flag_reg, cp, (dib >> b) & 1u);
      if (t & 2) hit(2ll * s + (1 - b));
    }
  }
  if (s1 == n && (k == 0 ? n <= CORR_CHUNK : true)) {  // this chunk holds the channel's last symbol (or there is none)
    CorrChanState o;
    o.data_reg = data_reg; o.flag_reg = flag_reg; o.nbits = old.nbits + 2ll * n;
    a.state_out[c] = o;
  }
}

}  // namespace grb
