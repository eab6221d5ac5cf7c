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
//   warps 2,3  CORE    the Mueller & Mueller recursion only, MMW_TRIP symbols per trip, branch free.  The 8
//                      input samples and the interpolator row are fetched from shared memory
//                      speculatively, the step is computed, and one predicate (input landed, queue
//                      slots free, ordinary forward step) selects whether the new state is kept; the
//                      soft symbol of a step that is not kept is stored to a scratch word.  floor()
//                      and rint() are done with the 1.5*2^23 trick (one FADD / FFMA instead of a
//                      conversion-unit round trip), and the two shared-memory addresses of the NEXT
//                      step (ring slot of row ii, interpolator row of mu) are carried as state and
//                      derived from those two results by integer multiply-adds, so the loads of step
//                      k+1 hang off the last addition of step k by three instructions.  The
//                      timing-error term picks one of the four exactly equivalent sums.  Anything
//                      unusual (a backward step, the last few symbols of a call, mu >= 2^15, more
//                      than 2^22 rows in one call) takes a plain one-symbol path that reads the input
//                      straight from global memory.
//   warps 4,5  POST    takes soft symbols from a shared-memory queue, eight at a time: soft symbol
//                      -> HBM, 4-level (or binary) slicer, dibit map, bit unpack, access-code
//                      correlation over a 16-bit window (__popc over the 64-bit shift register at
//                      all 16 positions at once; only the registers are carried), sync-hit list.
//   warps 0,1  LOADER  keeps a per-lane ring of the lane's input column in shared memory filled
//                      RING-8 rows ahead of the loop with cp.async (LDGSTS: no register staging) and
//                      publishes how far the data has landed.  HBM latency never meets the loop.
//   (warp id % 4 = scheduler: the core warps have theirs to themselves, loader and post share.)
//
// Measured (ncu source counters, profiles/README.md): ~66 instructions and ~320 cycles per symbol; the
// interpolator rows (one per lane, 32 bytes) cost 13.6 shared-memory wavefronts per LDS.128 instead of 4.
// A table replicated per bank group removes those conflicts but its 16-33 KB push the kernel past the
// 48 KB at which it co-resides with the front kernels, and the kernel time did not move: not kept.
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
#define MMW_TRIP 8            // symbols the core warp attempts per trip (between two looks at the other warps' words)
#define MMW_EMPTY 0x7fc0deadu // queue slot is empty (a quiet-NaN payload the arithmetic cannot produce
                              // from finite data; a colliding input NaN is re-encoded as 0x7fc00000)
#define MMW_CH 64             // channels per CTA
#define MMW_THREADS 192
#define MMW_TABREP 1          // copies of the interpolator table in shared memory (1 = the plain table; see below)
#define MMW_TABROW (2 * MMW_TABREP * 16)  // bytes per interpolator row: [2 halves][MMW_TABREP copies][4 floats]
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
  return (size_t)(ring + 8) * MMW_CH * 4 + 129 * MMW_TABROW + MMW_Q * MMW_CH * 4 + 4 * MMW_CH * 4 + 256;
}

// 6 warps x 64 registers leave room on the SM for the big-tile front kernels this kernel runs next to; at 48 the
// core loop's ten shared-memory loads per symbol were issued one at a time, each just before its use
#ifndef MMW_REGS
#define MMW_REGS 48
#endif
template <int RING, int ORDER>
__global__ void __maxnreg__(MMW_REGS) mm_ws_kernel(const MMArgs a) {
  extern __shared__ __align__(16) float mmw_smem[];
  // Interpolator table, [129 rows][2 halves][MMW_TABREP copies][4 floats].  With MMW_TABREP = 1 this is the plain
  // [129][8] table: every lane has its own mu, so a lane's two LDS.128 of its row cost 13.6 shared-memory wavefronts
  // each instead of 4.  MMW_TABREP > 1 sends lane l to copy l % MMW_TABREP, i.e. spreads the eight lanes of a quarter
  // warp over disjoint bank groups (8 copies: exactly 4 wavefronts), but the extra 16-33 KB push the kernel past the
  // 48 KB at which it co-resides with the 128 KB tiles of the branch filter, and the kernel time did not move
  // (profiles/README.md): 1 is what ships.
  float* tab = mmw_smem;
  float* ring = tab + 129 * (MMW_TABROW / 4);        // [RING + 8][64]; rows RING..RING+7 mirror rows 0..7
  unsigned* q = reinterpret_cast<unsigned*>(ring + (RING + 8) * MMW_CH);    // [MMW_Q][64]
  int* pub_ii = reinterpret_cast<int*>(q + MMW_Q * MMW_CH);    // [64] core -> loader: current input position
  int* pub_filled = pub_ii + MMW_CH;                           // [64] loader -> core: rows < this have landed
  int* pub_done = pub_filled + MMW_CH;                         // [64] core -> loader/post: symbols produced + 1
  int* dump = pub_done + MMW_CH;                               // [64] where the queue store of an uncommitted step lands
  unsigned char* smap = reinterpret_cast<unsigned char*>(dump + MMW_CH);  // [256] gr_map_bb table

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // warp -> scheduler is warp id % 4: the core warps (2, 3) have a scheduler each to themselves; the loader (0, 1)
  // and post (4, 5) warps share the other two
  const int role = warp < 2 ? 0 : warp < 4 ? 2 : 1;  // 0 loader, 1 post, 2 core
  const int cl = (warp & 1) * 32 + lane;             // channel within the CTA
  const int c = blockIdx.x * MMW_CH + cl;
  const bool valid = c < a.nchan;
  for (int i = threadIdx.x; i < 129 * 8 * MMW_TABREP; i += MMW_THREADS) {
    const int rep = i % MMW_TABREP, j = (i / MMW_TABREP) & 7, row = i / (8 * MMW_TABREP);
    tab[row * (MMW_TABROW / 4) + (j >> 2) * (MMW_TABREP * 4) + rep * 4 + (j & 3)] = a.mmse_eff[row * 8 + j];
  }
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
  if (role == 2) {
    pub_ii[cl] = ii0;
    pub_filled[cl] = ii0;
    pub_done[cl] = valid ? 0 : 1;
  }
  __syncthreads();

  const size_t nchan = (size_t)a.nchan;
  const float* __restrict__ col = a.in + (valid ? c : 0);
  const unsigned ring_lane = (unsigned)__cvta_generic_to_shared(ring + cl);
  const unsigned q_lane = (unsigned)__cvta_generic_to_shared(q + cl);
  const unsigned dump_lane = (unsigned)__cvta_generic_to_shared(dump + cl);
  constexpr unsigned RP = MMW_CH * 4;  // ring / queue row pitch in bytes

  if (role == 0) {
    // ---------------------------------------------------------------------------- LOADER
    // Rows [filled - RING, filled) of the lane's column sit in ring slot (row % RING); the loader
    // runs ahead to pub_ii + RING - BACK, so it only ever overwrites rows < pub_ii - BACK.
    int filled = ii0, g1 = ii0;  // g1: fill level after the previous trip's group
    const float* gp = col + (size_t)ii0 * nchan;
    while (true) {
      const int cur = mmw_ldv(pub_ii + cl);
      const int fin = mmw_ldv(pub_done + cl);
      const int want = (valid && !fin) ? min(cur + (RING - MMW_BACK), ninput) : filled;
      while (filled < want) {
        const unsigned slot = (unsigned)filled & (RING - 1);
        const unsigned dst = ring_lane + slot * RP;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(gp) : "memory");
        if (slot < 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + RING * RP), "l"(gp) : "memory");
        gp += nchan;
        filled++;
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      // everything but the group just committed has landed: rows < g1
      asm volatile("cp.async.wait_group 1;" ::: "memory");
      __threadfence_block();
      mmw_stv(pub_filled + cl, g1);
      g1 = filled;
      if (__all_sync(0xffffffffu, fin != 0)) break;
      __nanosleep(200);  // ~4 symbols of the loop; the ring holds > 40
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    return;
  }

  if (role == 2) {
    // ---------------------------------------------------------------------------- CORE
    float mu = st.mu, omega = st.omega, last = st.last_sample;
    int ii = ii0, oo = 0, hi = ii0;
    int clamped = clamp0 ? 1 : 0;
    bool careful = false;  // the lane's next step goes through the one-symbol path below
    const MMParams mp = a.p;
    constexpr int order = ORDER;
    const int max_out = valid ? a.max_out : 0;
    float sl = last < 0.f ? -1.0f : 1.0f;  // slice(last_sample) (:89-93), carried so that it is off the critical chain
    const unsigned tab_s = (unsigned)__cvta_generic_to_shared(tab);
    const unsigned ring_s = (unsigned)__cvta_generic_to_shared(ring);
    // The two shared-memory addresses of a step are carried as state, so that the next step's loads hang off the
    // floor / rint of this step by three integer instructions instead of going through mu and ii:
    //   iib = ii * 256 + cl * 4   byte offset of (row ii, this lane) in an unbounded ring; & RMASK = the ring slot
    //   ta  = imu * MMW_TABROW + (lane % MMW_TABREP) * 16   shared-memory address of this lane's copy of the interpolator row for the CURRENT mu
    constexpr unsigned RMASK = RING * RP - 4;
    constexpr int FAR = 1 << 22;  // rows beyond this go through the one-symbol path (iib stays inside 31 bits)
    const unsigned rep16 = (unsigned)(lane % MMW_TABREP) * 16u;
    const unsigned tak = (128u * MMW_TABROW - MMW_TABROW) * (unsigned)MMW_MAGIC_BITS + rep16 + tab_s;
    auto ta_of = [&](float m) {
      return ((unsigned)__float_as_int(__fmaf_rn(m, 128.0f, MMW_MAGIC)) & 0xffu) * (unsigned)MMW_TABROW + rep16 + tab_s;
    };
    int iib = ii * 256 + cl * 4;
    unsigned ta = ta_of(mu);
    while (true) {
      // ---- MMW_TRIP symbols, committed only while nothing unusual happens ---------------------------
      // rows < pub_filled have landed; (ii <= fs8) == (ii + 8 <= pub_filled && ii < ni)
      const int fs8 = min(min(mmw_ldv(pub_filled + cl), ninput - 1) - 8, FAR);
      const int fs8b = fs8 * 256 + 255;  // ii <= fs8  <=>  iib <= fs8b  (cl * 4 < 256)
      // the post warp empties slots in order, so a free slot oo+TRIP-1 means oo..oo+TRIP-1 are free
      const bool qfree = mmw_ldq(q_lane + (unsigned)((oo + MMW_TRIP - 1) & (MMW_Q - 1)) * RP) == MMW_EMPTY;
      const bool fast = !careful && oo + MMW_TRIP <= max_out && qfree;
#pragma unroll
      for (int k = 0; k < MMW_TRIP; k++) {
        // speculative fetch: any address inside the ring is readable; the predicate below says
        // whether rows ii..ii+7 of this lane's column are really the ones in these slots
        const unsigned src = ring_s + ((unsigned)iib & RMASK);
        float v[8], cf[8];
        const unsigned tad = ta;
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(cf[0]), "=f"(cf[1]), "=f"(cf[2]), "=f"(cf[3]) : "r"(tad));
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4+%5];" : "=f"(cf[4]), "=f"(cf[5]), "=f"(cf[6]), "=f"(cf[7]) : "r"(tad), "n"(MMW_TABREP * 16));
#pragma unroll
        for (int i = 0; i < 8; i++) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v[i]) : "r"(src + i * RP));
        const float o = mmse8(cf, v, order);
        // mm_update (gr_math.cuh) restated for the shortest dependent chain.  fmul(+-1, x) is exact,
        // so mm_val = slice(last)*o - slice(o)*last is one of four sums, each rounded once exactly
        // like the reference's subtraction
        // slice(last)*o is +-o exactly; slice(o)*last is +-last exactly, and x - (-y) == x + y: both
        // candidates are computed, the sign of o picks one
        const float so_o = __fmul_rn(sl, o);
        const bool on = o < 0.f;
        const float mm_val = on ? __fadd_rn(so_o, last) : __fsub_rn(so_o, last);
        float om = __fadd_rn(omega, __fmul_rn(mp.gain_omega, mm_val));
        om = __fadd_rn(mp.omega_mid, branchless_clip(__fsub_rn(om, mp.omega_mid), mp.omega_relative_limit));
        const float m2 = __fadd_rn(__fadd_rn(mu, om), __fmul_rn(mp.gain_mu, mm_val));
        // floor(m2) by adding 1.5*2^23 rounding towards minus infinity, rint(m2 * 128) by the FFMA (the product is
        // exact, the sum rounds once to nearest even): both exact for 0 <= m2 < 2^15.  The next mu is
        // m2 - floor(m2) (exact), so the next imu = rint(mu * 128) = rint(m2 * 128) - 128 * floor(m2): the
        // interpolator row of the next step needs neither the new mu nor a conversion
        const unsigned tb = (unsigned)__float_as_int(__fadd_rd(m2, MMW_MAGIC));
        const unsigned ub = (unsigned)__float_as_int(__fmaf_rn(m2, 128.0f, MMW_MAGIC));
        // imu * ROW = ub * ROW - tb * 128 * ROW + (128 * ROW - ROW) * MMW_MAGIC_BITS (mod 2^32); no mask is needed: tan
        // is only ever committed for a plain step, where imu is in [0, 128]
        const unsigned tan = ub * (unsigned)MMW_TABROW + (tb * (0u - 128u * MMW_TABROW) + tak);
        const int iibn = (int)(tb * 256u + ((unsigned)iib - 256u * (unsigned)MMW_MAGIC_BITS));  // ii + floor(m2)
        const float mu2 = __fsub_rn(m2, __fsub_rn(__uint_as_float(tb), MMW_MAGIC));
        unsigned ob = __float_as_uint(o);
        if (ob == MMW_EMPTY) ob = 0x7fc00000u;
        // forward step with the tricks valid: 0 <= m2 < 2^15, one unsigned compare on the bit pattern
        // (negative values, -0, NaN and Inf all have larger patterns)
        const bool plain = __float_as_uint(m2) < 0x47000000u;
        // branch free commit: the state registers are selected and the queue store of an uncommitted step goes to
        // a scratch word (a branch here costs more than the whole arithmetic of the step and splits the four
        // symbols into basic blocks)
        const bool ok = fast && iib <= fs8b;
        const bool commit = ok && plain;
        mmw_stq(commit ? q_lane + (unsigned)(oo & (MMW_Q - 1)) * RP : dump_lane, ob);
        mu = commit ? mu2 : mu;
        omega = commit ? om : omega;
        last = commit ? o : last;
        sl = commit ? (on ? -1.0f : 1.0f) : sl;
        iib = commit ? iibn : iib;
        ta = commit ? tan : ta;
        oo += commit ? 1 : 0;
        careful = careful || (ok && !plain);  // nothing committed: the step is redone below
      }
      ii = iib >> 8;
      careful = careful || ii >= FAR;
      mmw_stv(pub_ii + cl, ii);
      // ---- one symbol the plain way: backward steps, the tail of a call, out-of-range mu --------
      // (the whole input is in global memory before the kernel starts; the ring is only a latency
      // optimisation, so this path depends on nobody)
      const bool act = oo < max_out && ii < ni;
      if (!__any_sync(0xffffffffu, act)) break;
      const bool tail = oo + MMW_TRIP > max_out;  // fewer than a trip's output slots left in this call
      if (__any_sync(0xffffffffu, act && (careful || tail))) {
        const unsigned qslot = q_lane + (unsigned)(oo & (MMW_Q - 1)) * RP;
        if (act && (careful || tail) && mmw_ldq(qslot) == MMW_EMPTY) {
          float v[8], cf[8];
#pragma unroll
          for (int i = 0; i < 8; i++) v[i] = __ldg(col + (size_t)(ii + i) * nchan);
          const float* tp = tab + (MMW_TABROW / 4) * mm_imu(mu);  // copy 0 of the row
#pragma unroll
          for (int i = 0; i < 8; i++) cf[i] = tp[(i >> 2) * (MMW_TABREP * 4) + (i & 3)];
          const float o = mmse8(cf, v, order);
          unsigned ob = __float_as_uint(o);
          if (ob == MMW_EMPTY) ob = 0x7fc00000u;
          mmw_stq(qslot, ob);
          oo++;
          MMState s;
          s.mu = mu; s.omega = omega; s.last_sample = last;
          hi = max(hi, ii);
          ii += mm_update(s, mp, o);
          mu = s.mu; omega = s.omega; last = s.last_sample;
          sl = last < 0.f ? -1.0f : 1.0f;
          if (ii < 0) { ii = 0; clamped++; }
          // the ring still holds rows >= hi - BACK (the loader never overwrites rows >= pub_ii - BACK
          // and every published position is <= hi); older rows keep coming from global memory
          careful = ii < max(hi - MMW_BACK, ii0);
          iib = ii * 256 + cl * 4;
          ta = ta_of(mu);
          mmw_stv(pub_ii + cl, min(ii, hi));
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
    // window form of the correlator: valid when a match cannot raise its flag inside the same 16 bits
    const int code_len = cp.flag_bit ? 64 - (__ffsll((long long)cp.flag_bit) - 1) : 0;
    const bool windowed = corr_on && kbits == 2 && bp == nullptr && code_len >= 16;
    const int flag_shift = 64 - code_len;
    const unsigned code_hi = (unsigned)(cp.access_code >> 32), code_lo = (unsigned)cp.access_code;
    const unsigned mask_hi = (unsigned)(cp.mask >> 32), mask_lo = (unsigned)cp.mask;
    int consumed = 0, ob = 0;
    bool finished = !valid;

    auto hit = [&](int bit) {
      const int h = atomicAdd(a.corr.nhits, 1);
      if (h < a.corr.max_hits) { a.corr.hits[h].channel = c; a.corr.hits[h].pad = 0; a.corr.hits[h].bit_index = cs.nbits + bit; }
    };
    // one soft symbol the plain way: HBM store, slicer, symbol store, dibit -> bits -> correlator
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
          if (t & 2) hit(ob);
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
          if (windowed) {
            // Eight dibits = 16 bits at once.  Before bit j the data register is (data << j) | (the
            // first j new bits), so all 16 mismatch counts are independent funnel shifts + popcounts;
            // a match at bit j lands in the flag register at bit (64 - len) + (15 - j) after the 16
            // shifts, and the flags that reach bit 63 during these 16 bits are the register's top 16
            // bits as they are now (len >= 16: no match of this window can get there yet).
            unsigned bits16 = 0;
#pragma unroll
            for (int i = 0; i < MMW_PB; i++) {
              const float o = __uint_as_float(w[i]);
              op[(size_t)i * nchan] = o;
              unsigned char d = 0;
              if (slv == 4) d = slice4(o, avg, s_alpha, s_beta);
              else if (slv == 2) d = slice2(o);
              if (sp) sp[(size_t)i * nchan] = d;
              bits16 |= ((unsigned)smap[d] & 3u) << (14 - 2 * i);
            }
            op += MMW_PB * nchan;
            if (sp) sp += MMW_PB * nchan;
            const unsigned dhi = (unsigned)(cs.data_reg >> 32), dlo = (unsigned)cs.data_reg;
            const unsigned inb = bits16 << 16;
            unsigned mm = 0;
#pragma unroll
            for (int j = 0; j < 16; j++) {
              const unsigned shi = __funnelshift_l(dlo, dhi, j), slo = __funnelshift_l(inb, dlo, j);
              const unsigned nwrong = __popc((shi ^ code_hi) & mask_hi) + __popc((slo ^ code_lo) & mask_lo);
              mm |= (nwrong <= cp.threshold ? 1u : 0u) << (15 - j);
            }
            const unsigned hits16 = (unsigned)(cs.flag_reg >> 48);
            if (hits16) {
              for (int j = 0; j < 16; j++)
                if (hits16 & (0x8000u >> j)) hit(ob + j);
            }
            cs.data_reg = (cs.data_reg << 16) | bits16;
            cs.flag_reg = (cs.flag_reg << 16) | ((unsigned long long)mm << flag_shift);
            ob += 16;
          } else {
#pragma unroll 1
            for (int i = 0; i < MMW_PB; i++) emit(__uint_as_float(w[i]));
          }
          consumed += MMW_PB;
        } else {
          // fewer than a batch queued: only drain one by one once the core has finished
          const int dn = mmw_ldv(pub_done + cl);
          if (dn != 0) {
            const unsigned w0 = mmw_ldq(slot[0]);
            if (w0 != MMW_EMPTY) {
              mmw_stq(slot[0], MMW_EMPTY);
              emit(__uint_as_float(w0));
              consumed++;
            } else if (consumed == dn - 1) {
              finished = true;
            }
          } else {
            __nanosleep(100);
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
