// digital_clock_recovery_mm_ff::general_work, batched over channels and warp specialised
// (digital_clock_recovery_mm_ff.cc:102-139 + gri_mmse_fir_interpolator.cc:61-71), with the
// slicer -> gr_map_bb -> gr_unpack_k_bits_bb -> digital_correlate_access_code_bb epilogue.
//
// Why it looks like this.  The loop is sequential in time (mu, omega, last_sample feed back into
// the NEXT input index), so a channel can never use more than one thread, and 8000 channels are
// only 250 warps on 592 warp schedulers: the kernel runs at the latency of ONE warp's dependent
// instruction chain per symbol, whatever the grid.  Everything that is not on that chain is
// therefore moved out of the warp that carries it.  A CTA owns 64 neighbouring channels (two
// groups of 32, lane = channel) and six warps:
//
//   warps 0,1  CORE    the Mueller & Mueller recursion only.  Branch free: the 8 input samples
//                      and the interpolator taps are fetched from shared memory speculatively,
//                      the step is computed, and a predicate (input landed, queue slot free)
//                      decides whether the new state is committed.  floor()/rint() are done with
//                      the 1.5*2^23 trick (one FADD/FFMA instead of a conversion-unit round trip);
//                      the timing-error term picks one of the four exactly equivalent sums.
//   warps 2,3  POST    takes soft symbols from a shared-memory queue, eight at a time: soft symbol
//                      -> HBM, 4-level (or binary) slicer, dibit map, bit unpack, access-code
//                      correlation (__popcll over the 64-bit shift register), sync-hit list.
//   warps 4,5  LOADER  keeps a per-lane ring of the lane's input column in shared memory filled
//                      RING-8 rows ahead of the loop with cp.async (LDGSTS: no register staging) and
//                      publishes how far the data has landed.  HBM latency never meets the loop.
//
// Warps talk through shared memory only (all six are co-resident by construction, so spinning is
// safe).  Queue slots carry their own full/empty state (a reserved NaN pattern = empty), so no
// ordering between different words is needed there; the loader publishes a row count AFTER
// cp.async.wait_group + a CTA fence.
#pragma once
#include <cuda_runtime.h>
#include "gr_math.cuh"
#include "kernels_demod.cuh"

namespace grb {

#define MMW_BACK 8            // rows kept behind the furthest position for (rare) backward steps
#define MMW_Q 32              // soft-symbol queue depth per lane
#define MMW_PB 8              // symbols the post warp takes per batch
#define MMW_EMPTY 0x7fc0deadu // queue slot is empty (a quiet-NaN payload the arithmetic cannot produce
                              // from finite data; a colliding input NaN is re-encoded as 0x7fc00000)
#define MMW_CH 64             // channels per CTA
#define MMW_THREADS 192
#define MMW_MAGIC 12582912.0f // 1.5 * 2^23: adding it leaves round(x) / floor(x) in the low mantissa bits
#define MMW_MAGIC_BITS 0x4b400000

__device__ __forceinline__ int mmw_ldv(const int* p) {
  int v;
  asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"((unsigned)__cvta_generic_to_shared(p)) : "memory");
  return v;
}
__device__ __forceinline__ void mmw_stv(int* p, int v) {
  asm volatile("st.volatile.shared.s32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(p)), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned mmw_ldq(unsigned addr) {
  unsigned v;
  asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void mmw_stq(unsigned addr, unsigned v) {
  asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

static inline size_t mm_ws_smem_bytes(int ring) {
  return (size_t)(ring + 8) * MMW_CH * 4 + 129 * 8 * 4 + MMW_Q * MMW_CH * 4 + 3 * MMW_CH * 4 + 256;
}

// The loop update (mm_update in gr_math.cuh) restated for the shortest dependent chain.  Identical
// results: fmul(+-1, x) is exact, so mm_val = slice(last)*out - slice(out)*last is one of the four
// sums below, each rounded once exactly like the reference's subtraction; floor() by adding
// 1.5*2^23 with round-towards-minus-infinity (exact for |mu| < 2^22, else the slow way).
__device__ __forceinline__ int mmw_update(float& mu, float& omega, float& last, const MMParams& p, float out) {
  const bool ln = last < 0.f, on = out < 0.f;
  const float mm_val = ln ? (on ? __fadd_rn(-out, last) : __fsub_rn(-out, last))
                          : (on ? __fadd_rn(out, last) : __fsub_rn(out, last));
  last = out;
  float om = __fadd_rn(omega, __fmul_rn(p.gain_omega, mm_val));
  om = __fadd_rn(p.omega_mid, branchless_clip(__fsub_rn(om, p.omega_mid), p.omega_relative_limit));
  omega = om;
  const float m2 = __fadd_rn(__fadd_rn(mu, om), __fmul_rn(p.gain_mu, mm_val));
  int adv;
  if (fabsf(m2) < 4194304.0f) {
    const float t = __fadd_rd(m2, MMW_MAGIC);
    adv = __float_as_int(t) - MMW_MAGIC_BITS;
    mu = __fsub_rn(m2, __fsub_rn(t, MMW_MAGIC));
  } else {  // also NaN
    const float fl = floorf(m2);
    mu = __fsub_rn(m2, fl);
    adv = (int)fl;
  }
  return adv;
}

template <int RING>
__global__ void __launch_bounds__(MMW_THREADS) mm_ws_kernel(const MMArgs a) {
  extern __shared__ __align__(16) float mmw_smem[];
  float* ring = mmw_smem;                            // [RING + 8][64]; rows RING..RING+7 mirror rows 0..7
  float* tab = ring + (RING + 8) * MMW_CH;           // [129][8] interpolator coefficients
  unsigned* q = reinterpret_cast<unsigned*>(tab + 129 * 8);    // [MMW_Q][64]
  int* pub_ii = reinterpret_cast<int*>(q + MMW_Q * MMW_CH);    // [64] core -> loader: current input position
  int* pub_filled = pub_ii + MMW_CH;                           // [64] loader -> core: rows < this have landed
  int* pub_done = pub_filled + MMW_CH;                         // [64] core -> loader/post: symbols produced + 1
  unsigned char* smap = reinterpret_cast<unsigned char*>(pub_done + MMW_CH);  // [256] gr_map_bb table

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int role = warp >> 1;                        // 0 core, 1 post, 2 loader
  const int cl = (warp & 1) * 32 + lane;             // channel within the CTA
  const int c = blockIdx.x * MMW_CH + cl;
  const bool valid = c < a.nchan;
  for (int i = threadIdx.x; i < 129 * 8; i += MMW_THREADS) tab[i] = a.mmse_eff[i];
  for (int i = threadIdx.x; i < 256; i += MMW_THREADS) smap[i] = a.corr.map[i];
  for (int i = threadIdx.x; i < MMW_Q * MMW_CH; i += MMW_THREADS) q[i] = MMW_EMPTY;

  const int ninput = (int)a.ninput;
  const int ni = ninput - 8;  // :112
  MMChanState st;
  st.mu = 0.f; st.omega = 0.f; st.last_sample = 0.f; st.slicer_avg = 0.f; st.next_abs = a.abs_row0; st.clamped = 0; st.overflow = 0;
  if (valid) st = a.state[c];
  // floor(mu) can be negative when gain_mu*mm_val < -omega (unnormalised input): the reference then
  // re-reads older items of its circular buffer.  The caller keeps a carry of older rows in front of
  // each block for that; stepping even further back is clamped (and counted) instead of reading
  // out of bounds, which is where the reference's behaviour is undefined anyway.
  int ii0 = (int)(st.next_abs - a.abs_row0);  // may be > 0: samples already consumed
  const bool clamp0 = ii0 < 0;
  if (clamp0) ii0 = 0;
  if (role == 0) {
    pub_ii[cl] = ii0;
    pub_filled[cl] = ii0;
    pub_done[cl] = valid ? 0 : 1;
  }
  __syncthreads();

  const size_t nchan = (size_t)a.nchan;
  const float* __restrict__ col = a.in + (valid ? c : 0);
  const unsigned ring_lane = (unsigned)__cvta_generic_to_shared(ring + cl);
  const unsigned q_lane = (unsigned)__cvta_generic_to_shared(q + cl);
  constexpr unsigned RP = MMW_CH * 4;  // ring / queue row pitch in bytes

  if (role == 2) {
    // ---------------------------------------------------------------------------- LOADER
    int filled = ii0, g1 = ii0, g2 = ii0;  // fill level after the last / the second to last committed group
    const float* gp = col + (size_t)ii0 * nchan;
    while (true) {
      const int cur = mmw_ldv(pub_ii + cl);
      const int fin = mmw_ldv(pub_done + cl);
      const int want = (valid && !fin) ? min(cur + (RING - MMW_BACK), ninput) : filled;
      const bool any = filled < want;
      while (filled < want) {
        const unsigned slot = (unsigned)filled & (RING - 1);
        const unsigned dst = ring_lane + slot * RP;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(gp) : "memory");
        if (slot < 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + RING * RP), "l"(gp) : "memory");
        gp += nchan;
        filled++;
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      // everything but the two most recent groups has landed: rows < g2
      asm volatile("cp.async.wait_group 2;" ::: "memory");
      __threadfence_block();
      mmw_stv(pub_filled + cl, g2);
      g2 = g1;
      g1 = filled;
      if (__all_sync(0xffffffffu, fin != 0)) break;
      if (!__any_sync(0xffffffffu, any)) __nanosleep(64);
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    return;
  }

  if (role == 0) {
    // ---------------------------------------------------------------------------- CORE
    float mu = st.mu, omega = st.omega, last = st.last_sample;
    int ii = ii0, oo = 0, hi = ii0, filled_seen = ii0;
    int clamped = clamp0 ? 1 : 0;
    const MMParams mp = a.p;
    const int order = a.order, max_out = a.max_out;
    const unsigned tab_s = (unsigned)__cvta_generic_to_shared(tab);
    while (true) {
      bool need_slow = false;
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const bool act = valid && oo < max_out && ii < ni;
        // speculative fetch: any address inside the ring is readable; `ready` below says whether
        // rows ii..ii+7 of this lane's column are really the ones in these slots
        const unsigned src = ring_lane + ((unsigned)ii & (RING - 1)) * RP;
        float v[8], cf[8];
#pragma unroll
        for (int i = 0; i < 8; i++) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v[i]) : "r"(src + i * RP));
        // imu = (int) rint(mu * 128): the product is exact, the FFMA rounds once to nearest even
        const unsigned imu = (unsigned)__float_as_int(__fmaf_rn(mu, 128.0f, MMW_MAGIC)) & 0xffu;
        const unsigned ta = tab_s + imu * 32u;
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(cf[0]), "=f"(cf[1]), "=f"(cf[2]), "=f"(cf[3]) : "r"(ta));
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4+16];" : "=f"(cf[4]), "=f"(cf[5]), "=f"(cf[6]), "=f"(cf[7]) : "r"(ta));
        const unsigned qslot = q_lane + (unsigned)(oo & (MMW_Q - 1)) * RP;
        const unsigned qw = mmw_ldq(qslot);
        // rows [max(hi - BACK, ii0), pub_filled) of this lane's column are in the ring and cannot be
        // recycled by the loader (it never overwrites rows >= pub_ii - BACK, and pub_ii <= hi)
        const bool behind = ii < max(hi - MMW_BACK, ii0);
        if (ii + 8 > filled_seen) filled_seen = mmw_ldv(pub_filled + cl);
        const bool ready = act && !behind && ii + 8 <= filled_seen && qw == MMW_EMPTY;
        need_slow = need_slow || (act && behind);

        const float o = mmse8(cf, v, order);
        float mu2 = mu, om2 = omega, la2 = last;
        int ii2 = ii + mmw_update(mu2, om2, la2, mp, o);
        unsigned ob = __float_as_uint(o);
        if (ob == MMW_EMPTY) ob = 0x7fc00000u;
        if (ready) {
          mmw_stq(qslot, ob);
          if (ii2 < 0) { ii2 = 0; clamped++; }
          mu = mu2; omega = om2; last = la2; ii = ii2;
          oo++;
          hi = max(hi, ii);
          mmw_stv(pub_ii + cl, ii);
        }
      }
      const bool act = valid && oo < max_out && ii < ni;
      if (!__any_sync(0xffffffffu, act)) break;
      if (__any_sync(0xffffffffu, need_slow)) {
        // a lane stepped back past what its ring still holds (unnormalised input): one step straight
        // from global memory; the ring catches up with it as ii grows again
        const bool behind = ii < max(hi - MMW_BACK, ii0);
        const unsigned qslot = q_lane + (unsigned)(oo & (MMW_Q - 1)) * RP;
        if (act && behind && mmw_ldq(qslot) == MMW_EMPTY) {
          float v[8], cf[8];
#pragma unroll
          for (int i = 0; i < 8; i++) v[i] = __ldg(col + (size_t)(ii + i) * nchan);
          const float* tp = tab + 8 * mm_imu(mu);
#pragma unroll
          for (int i = 0; i < 8; i++) cf[i] = tp[i];
          const float o = mmse8(cf, v, order);
          unsigned ob = __float_as_uint(o);
          if (ob == MMW_EMPTY) ob = 0x7fc00000u;
          mmw_stq(qslot, ob);
          oo++;
          ii += mmw_update(mu, omega, last, mp, o);
          if (ii < 0) { ii = 0; clamped++; }
          hi = max(hi, ii);
          mmw_stv(pub_ii + cl, ii);
        }
      }
    }
    mmw_stv(pub_done + cl, oo + 1);
    if (valid) {
      MMChanState* sp = a.state + c;
      sp->mu = mu; sp->omega = omega; sp->last_sample = last;
      sp->next_abs = a.abs_row0 + ii;
      sp->clamped = st.clamped + clamped;
      sp->overflow = st.overflow + (ii < ni ? 1 : 0);
      a.counts[c] = oo;
    }
    return;
  }

  // ------------------------------------------------------------------------------ POST
  {
    float avg = st.slicer_avg;
    CorrChanState cs;
    cs.data_reg = 0; cs.flag_reg = 0; cs.nbits = 0;
    const bool corr_on = a.corr.on != 0;
    if (corr_on && valid) cs = a.corr.state[c];
    float* op = a.out + (valid ? c : 0);
    unsigned char* sp = a.sliced ? a.sliced + (valid ? c : 0) : nullptr;
    unsigned char* bp = (corr_on && a.corr.out) ? a.corr.out + (valid ? c : 0) : nullptr;
    const int slv = a.slicer_levels, kbits = a.corr.bits_per_symbol;
    const float s_alpha = a.slicer_alpha, s_beta = a.slicer_beta;
    const CorrParams cp = a.corr.p;
    const bool discard = a.debug == 1;
    int consumed = 0, ob = 0;
    bool finished = !valid;

    // one soft symbol: HBM store, slicer, symbol store, dibit -> bits -> correlator
    auto emit = [&](float o) {
      *op = o;
      op += nchan;
      unsigned char d = 0;
      if (slv == 4) d = slice4(o, avg, s_alpha, s_beta);
      else if (slv == 2) d = slice2(o);
      if (sp) { *sp = d; sp += nchan; }
      if (corr_on) {
        const unsigned dib = smap[d];
        for (int b = kbits - 1; b >= 0; b--) {  // gr_unpack_k_bits_bb: MSB first
          const unsigned char t = corr_step(cs.data_reg, cs.flag_reg, cp, (dib >> b) & 1u);
          if (bp) { *bp = t; bp += nchan; }
          if (t & 2) {
            const int h = atomicAdd(a.corr.nhits, 1);
            if (h < a.corr.max_hits) { a.corr.hits[h].channel = c; a.corr.hits[h].pad = 0; a.corr.hits[h].bit_index = cs.nbits + ob; }
          }
          ob++;
        }
      }
    };

    while (true) {
      if (!finished) {
        unsigned slot[MMW_PB], w[MMW_PB];
#pragma unroll
        for (int i = 0; i < MMW_PB; i++) slot[i] = q_lane + (unsigned)((consumed + i) & (MMW_Q - 1)) * RP;
#pragma unroll
        for (int i = 0; i < MMW_PB; i++) w[i] = mmw_ldq(slot[i]);
        bool full = true;
#pragma unroll
        for (int i = 0; i < MMW_PB; i++) full = full && (w[i] != MMW_EMPTY);
        if (full) {
#pragma unroll
          for (int i = 0; i < MMW_PB; i++) mmw_stq(slot[i], MMW_EMPTY);
          if (!discard) {
            if (kbits == 2 && corr_on) {
              // the common DMR case, fully unrolled so that the eight symbols' independent work
              // (stores, slicer, table lookups, popcounts) overlaps; only the two shift registers
              // are carried from bit to bit
              unsigned dib[MMW_PB];
#pragma unroll
              for (int i = 0; i < MMW_PB; i++) {
                const float o = __uint_as_float(w[i]);
                op[(size_t)i * nchan] = o;
                unsigned char d = 0;
                if (slv == 4) d = slice4(o, avg, s_alpha, s_beta);
                else if (slv == 2) d = slice2(o);
                if (sp) sp[(size_t)i * nchan] = d;
                dib[i] = smap[d];
              }
              op += MMW_PB * nchan;
              if (sp) sp += MMW_PB * nchan;
#pragma unroll
              for (int i = 0; i < MMW_PB; i++) {
#pragma unroll
                for (int b = 1; b >= 0; b--) {
                  const unsigned char t = corr_step(cs.data_reg, cs.flag_reg, cp, (dib[i] >> b) & 1u);
                  if (bp) bp[(size_t)(2 * i + 1 - b) * nchan] = t;
                  if (t & 2) {
                    const int h = atomicAdd(a.corr.nhits, 1);
                    if (h < a.corr.max_hits) {
                      a.corr.hits[h].channel = c; a.corr.hits[h].pad = 0; a.corr.hits[h].bit_index = cs.nbits + ob + 2 * i + 1 - b;
                    }
                  }
                }
              }
              if (bp) bp += 2 * MMW_PB * nchan;
              ob += 2 * MMW_PB;
            } else {
#pragma unroll 1
              for (int i = 0; i < MMW_PB; i++) emit(__uint_as_float(w[i]));
            }
          }
          consumed += MMW_PB;
        } else {
          // fewer than a batch queued: only drain one by one once the core has finished
          const int dn = mmw_ldv(pub_done + cl);
          if (dn != 0) {
            const unsigned w0 = mmw_ldq(slot[0]);
            if (w0 != MMW_EMPTY) {
              mmw_stq(slot[0], MMW_EMPTY);
              if (!discard) emit(__uint_as_float(w0));
              consumed++;
            } else if (consumed == dn - 1) {
              finished = true;
            }
          }
        }
      }
      if (__all_sync(0xffffffffu, finished)) break;
    }
    if (valid) {
      a.state[c].slicer_avg = avg;
      if (corr_on) { cs.nbits += ob; a.corr.state[c] = cs; }
    }
  }
}

}  // namespace grb
