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
#include "packed_f32.cuh"
#include "tma.cuh"

namespace grb {

#define MMW_BACK 8            // rows kept behind the furthest position for (rare) backward steps
#define MMW_Q 32              // soft-symbol queue depth per lane
#define MMW_PB 8              // symbols the post warp takes per batch
#define MMW_TRIP 8            // symbols the core warp attempts per trip (between two looks at the other warps' words)
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

#ifdef MMW_STATS
// lab build only (tools/mm_microbench.py --stats): cycles and trip counts of the core warps
// [0] cycles in trips every lane committed, [1] such trips, [2] cycles in other trips (incl. replay), [3] such trips,
// [4] cycles in the one-symbol section, [5] one-symbol iterations, [6] total core cycles, [7] core warps
__device__ unsigned long long mmw_stats[12];  // [8] lane-trips blocked by the queue, [9] by input, [10] stopped half way
__device__ __forceinline__ long long mmw_clk() { long long c; asm volatile("mov.u64 %0, %%clock64;" : "=l"(c)); return c; }
#endif

// a ^ (b & c) in ONE instruction (the compiler shares b & c between two users and serialises two LOP3 otherwise)
__device__ __forceinline__ unsigned mmw_xor_and(unsigned a, unsigned b, unsigned c) {
  unsigned d;
  asm("lop3.b32 %0, %1, %2, %3, 0x78;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

#define MMW_G 8  // rows per bulk-copy group of the TMA loader (one mbarrier per group slot)
static inline size_t mm_ws_smem_bytes(int ring, int tabrep) {
  return (size_t)(ring + 8) * MMW_CH * 4 + 129 * (size_t)(2 * tabrep * 16) + MMW_Q * MMW_CH * 4 + 4 * MMW_CH * 4 + 256 +
         (size_t)(ring / MMW_G) * 8;
}
__device__ __forceinline__ bool mmw_mbar_test(uint64_t* bar, unsigned parity) {
  unsigned ok;
  asm volatile("{\n .reg .pred p;\n mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}

// ---- POST warp: takes soft symbols from the shared-memory queue, eight at a time: soft symbol -> HBM, slicer, dibit
// map, bit unpack, access-code correlation, sync-hit list (shared by both clock-recovery kernels)
template <int QD>
__device__ __forceinline__ void mmw_post_warp(const MMArgs& a, const MMChanState& st, const bool valid, const int c, const int cl,
                                              const unsigned q_lane, int* pub_done, const unsigned char* smap) {
  const size_t nchan = (size_t)a.nchan;
  constexpr unsigned RP = MMW_CH * 4;
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
  // the 4-entry dibit map as one byte (2 bits per slicer decision) when it fits: no table load per symbol
  unsigned map8 = 0;
  bool map_packed = true;
  for (int d = 0; d < 4; d++) { map_packed = map_packed && smap[d] < 4; map8 |= ((unsigned)smap[d] & 3u) << (2 * d); }
  // alpha == 0 (no DC tracking): avg = avg * 1 + sample * 0 stays avg exactly, nothing to carry
  const bool static_avg = s_alpha == 0.0f && s_beta == 1.0f && slv == 4;
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
    bool got = false;
    if (!finished) {
      unsigned slot[MMW_PB], w[MMW_PB];
#pragma unroll
      for (int i = 0; i < MMW_PB; i++) slot[i] = q_lane + (unsigned)((consumed + i) & (QD - 1)) * RP;
      // The core fills the slots in order (in-order shared-memory stores of one warp), so the LAST slot of a batch
      // being full means the whole batch is: ONE load per poll.  (Eight loads per poll from two post warps that had
      // caught up with the core took ~40 % of the shared-memory pipe away from the loop they were waiting for.)
      w[MMW_PB - 1] = mmw_ldq(slot[MMW_PB - 1]);
      const bool full = w[MMW_PB - 1] != MMW_EMPTY;
      if (full) {
        got = true;
#pragma unroll
        for (int i = 0; i < MMW_PB - 1; i++) w[i] = mmw_ldq(slot[i]);
#pragma unroll
        for (int i = 0; i < MMW_PB; i++) mmw_stq(slot[i], MMW_EMPTY);
        if (windowed || !corr_on) {
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
            unsigned d = 0;
            if (static_avg) {
              const float t = __fsub_rn(o, avg);  // pager_slicer_fb.cc:52-68 with d_avg constant
              d = t > 0.f ? (t > 2.0f ? 3u : 2u) : (t < -2.0f ? 0u : 1u);
            } else if (slv == 4) d = slice4(o, avg, s_alpha, s_beta);
            else if (slv == 2) d = slice2(o);
            if (sp) sp[(size_t)i * nchan] = (unsigned char)d;
            const unsigned dib = map_packed ? (map8 >> (2 * d)) : (unsigned)smap[d];
            bits16 |= (dib & 3u) << (14 - 2 * i);
          }
          op += MMW_PB * nchan;
          if (sp) sp += MMW_PB * nchan;
          if (corr_on) {  // (the stand-alone block stops here: soft symbols + slicer only)
          const unsigned dhi = (unsigned)(cs.data_reg >> 32), dlo = (unsigned)cs.data_reg;
          const unsigned inb = bits16 << 16;
          // two stages: the mismatches of the upper word alone already exceed the threshold at almost every
          // position; the lower word is only looked at for the (rare) candidates
          unsigned mm = 0, cand = 0;
#pragma unroll
          for (int j = 0; j < 16; j++) {
            const unsigned shi = __funnelshift_l(dlo, dhi, j);
            if (__popc((shi ^ code_hi) & mask_hi) <= cp.threshold) cand |= 0x8000u >> j;
          }
          while (cand) {
            const int j = __clz(cand) - 16;
            cand &= ~(0x8000u >> j);
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
          }
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
            got = true;
            mmw_stq(slot[0], MMW_EMPTY);
            emit(__uint_as_float(w0));
            consumed++;
          } else if (consumed == dn - 1) {
            finished = true;
          }
        }
      }
    }
    if (__all_sync(0xffffffffu, finished)) break;
    if (!__any_sync(0xffffffffu, got)) __nanosleep(200);  // a trip of the core takes ~600 ns, the queue holds 4 to 8
  }
  if (valid) {
    a.state[c].slicer_avg = avg;
    if (a.state_out2) a.state_out2[c].slicer_avg = avg;
    if (corr_on) { cs.nbits += ob; a.corr.state[c] = cs; }
  }
}

// Template parameters: RING = rows of look-ahead ring per lane; ORDER = summation order of the interpolator;
// NREG = register cap (6 warps x 48 registers co-reside with the big-tile front kernels of a single-GPU chain; the
// stand-alone build, which is what a time shard runs, takes what the loop wants); TR = copies of the interpolator
// table (1: 4 KB, 13.6 shared-memory wavefronts per LDS.128 of a lane's row; 8: 33 KB, the conflict-free 4);
// CORE = 1: round-1 recursion (state selected under one commit predicate), 2: shortest dependent chain, 3: fewest
// instructions (below).  LD = 0: per-lane look-ahead ring filled with 4-byte cp.async (any channel count, any
// alignment), 1: rows shared by the CTA's 64 channels, moved by the bulk copy engine (TMA), 256 bytes per row.
template <int RING, int ORDER, int NREG, int TR, int CORE, int LD>
__global__ void __maxnreg__(NREG) mm_ws_kernel(const MMArgs a) {
  extern __shared__ __align__(16) float mmw_smem[];
  constexpr int TABROW = 2 * TR * 16;  // bytes per interpolator row: [2 halves][TR copies][4 floats]
  // The ring comes FIRST: its shared-memory address is then a link-time constant that ptxas folds into the
  // immediate offset of the core loop's LDS, so a ring address is one LOP3 away from the row index.
  float* ring = mmw_smem;                            // [RING + 8][64]; rows RING..RING+7 mirror rows 0..7
  unsigned* q = reinterpret_cast<unsigned*>(ring + (RING + 8) * MMW_CH);    // [MMW_Q][64]
  int* pub_ii = reinterpret_cast<int*>(q + MMW_Q * MMW_CH);    // [64] core -> loader: current input position
  int* pub_filled = pub_ii + MMW_CH;                           // [64] loader -> core: rows < this have landed
  int* pub_done = pub_filled + MMW_CH;                         // [64] core -> loader/post: symbols produced + 1
  int* dump = pub_done + MMW_CH;                               // [64] where the queue store of an uncommitted step lands
  unsigned char* smap = reinterpret_cast<unsigned char*>(dump + MMW_CH);  // [256] gr_map_bb table
  // Interpolator table, [129 rows][2 halves][TR copies][4 floats].  Every lane has its own mu, so with the plain
  // table (TR = 1) a lane's two LDS.128 of its row cost 13.6 shared-memory wavefronts each; TR = 8 sends lane l to
  // copy l % 8, i.e. spreads the eight lanes of a quarter warp over disjoint bank groups: exactly 4 wavefronts.
  float* tab = reinterpret_cast<float*>(smap + 256);
  uint64_t* gbar = reinterpret_cast<uint64_t*>(tab + 129 * (TABROW / 4));  // [RING / MMW_G] (LD = 1)

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // warp -> scheduler is warp id % 4: the core warps (2, 3) have a scheduler each to themselves; the loader (0, 1)
  // and post (4, 5) warps share the other two
  const int role = warp < 2 ? 0 : warp < 4 ? 2 : 1;  // 0 loader, 1 post, 2 core
  const int cl = (warp & 1) * 32 + lane;             // channel within the CTA
  const int c = blockIdx.x * MMW_CH + cl;
  const bool valid = c < a.nchan;
  for (int e = threadIdx.x; e < 129 * 8; e += MMW_THREADS) {  // one global load per coefficient, TR stores
    const float cv = __ldg(a.mmse_eff + e);
    const int j = e & 7, row = e >> 3;
#pragma unroll
    for (int rep = 0; rep < TR; rep++) tab[row * (TABROW / 4) + (j >> 2) * (TR * 4) + rep * 4 + (j & 3)] = cv;
  }
  for (int i = threadIdx.x; i < 256; i += MMW_THREADS) smap[i] = a.corr.map[i];
  for (int i = threadIdx.x; i < MMW_Q * MMW_CH; i += MMW_THREADS) q[i] = MMW_EMPTY;

  const int ninput = (int)a.ninput;
  const int ni = ninput - 8;  // :112
  MMChanState st;
  st.mu = 0.f; st.omega = 0.f; st.last_sample = 0.f; st.slicer_avg = 0.f; st.next_abs = a.abs_row0; st.clamped = 0; st.overflow = 0;
  if (valid) st = (a.state_in ? a.state_in : a.state)[c];
  // floor(mu) can be negative when gain_mu*mm_val < -omega (unnormalised input): the reference then
  // re-reads older items of its circular buffer.  The caller keeps a carry of older rows in front of
  // each block for that; stepping even further back is clamped (and counted) instead of reading
  // out of bounds, which is where the reference's behaviour is undefined anyway.
  int ii0 = (int)(st.next_abs - a.abs_row0);  // may be > 0: samples already consumed
  const bool clamp0 = ii0 < 0;
  if (clamp0) ii0 = 0;
  if (role == 2) {
    pub_ii[cl] = (LD == 1 && !valid) ? ninput : ii0;
    pub_filled[cl] = ii0;
    pub_done[cl] = valid ? 0 : 1;
  }
  if (LD == 1 && threadIdx.x < RING / MMW_G) mbar_init(gbar + threadIdx.x, 1);
  if (LD == 1) mbar_init_fence();
  __syncthreads();

  const size_t nchan = (size_t)a.nchan;
  const float* __restrict__ col = a.in + (valid ? c : 0);
  const unsigned ring_lane = (unsigned)__cvta_generic_to_shared(ring + cl);
  const unsigned q_lane = (unsigned)__cvta_generic_to_shared(q + cl);
  const unsigned dump_lane = (unsigned)__cvta_generic_to_shared(dump + cl);
  constexpr unsigned RP = MMW_CH * 4;  // ring / queue row pitch in bytes

  if (role == 0 && LD == 1) {
    // ---------------------------------------------------------------------------- LOADER (bulk copy engine)
    // The 64 channels of the CTA walk their columns at (nearly) the same pace: every call starts all of them within
    // a few rows of each other (each stopped where the previous block ran out of input) and omega is clipped to
    // +- omega_relative_limit, so one fill level serves all 64.  Rows are moved in groups of MMW_G by cp.async.bulk
    // (256 bytes per row, no per-lane instruction stream: this warp shares its scheduler with a post warp), each
    // group completes on its own mbarrier; rows [landed - RING, landed) sit in ring slot (row % RING).  The group
    // that would overwrite rows >= min(pub_ii) - BACK is not issued yet.  A lane that runs far ahead of the slowest
    // one simply waits for it (core: `fed`).
    if (warp != 0) return;
    constexpr int NB = RING / MMW_G;
    const int c0 = blockIdx.x * MMW_CH;
    const unsigned row_bytes = (unsigned)min(MMW_CH, a.nchan - c0) * 4u;
    int gmin = min(mmw_ldv(pub_ii + lane), mmw_ldv(pub_ii + lane + 32));
    gmin = __reduce_min_sync(0xffffffffu, gmin);
    const int first_g = max(gmin - MMW_BACK, 0) / MMW_G;  // groups first_g .. first_g + NB - 1 are the first use (phase 0) of the NB barriers
    int next_g = first_g, landed_g = first_g;
    const float* base = a.in + c0;
    while (true) {
      int cur = min(mmw_ldv(pub_ii + lane), mmw_ldv(pub_ii + lane + 32));
      cur = __reduce_min_sync(0xffffffffu, cur);
      const bool fin = mmw_ldv(pub_done + lane) != 0 && mmw_ldv(pub_done + lane + 32) != 0;
      if (__all_sync(0xffffffffu, fin)) break;
      const int limit_row = cur - MMW_BACK + RING;  // rows below this may be in the ring
      bool moved = false;
      while ((next_g + 1) * MMW_G <= limit_row && next_g * MMW_G < ninput && next_g - landed_g < NB) {
        const int slot_g = next_g & (NB - 1);
        const int r0 = next_g * MMW_G;
        const int nrows = min(MMW_G, ninput - r0);
        uint64_t* bar = gbar + slot_g;
        if (lane == 0) mbar_expect_tx(bar, row_bytes * (unsigned)nrows * (slot_g == 0 ? 2u : 1u));
        __syncwarp();
        if (lane < nrows) {
          const float* src = base + (size_t)(r0 + lane) * nchan;
          float* dst = ring + (size_t)(slot_g * MMW_G + lane) * MMW_CH;
          bulk_g2s(dst, src, row_bytes, bar);
          if (slot_g == 0) bulk_g2s(dst + (size_t)RING * MMW_CH, src, row_bytes, bar);  // rows RING..RING+7 mirror 0..7
        }
        next_g++;
        moved = true;
      }
      while (landed_g < next_g && mmw_mbar_test(gbar + (landed_g & (NB - 1)), (unsigned)((landed_g - first_g) / NB) & 1u)) {
        landed_g++;
        moved = true;
      }
      if (moved) {
        __threadfence_block();
        const int lv = min(landed_g * MMW_G, ninput);
        mmw_stv(pub_filled + lane, lv);
        mmw_stv(pub_filled + lane + 32, lv);
      } else {
        __nanosleep(40);
      }
    }
    // nothing may still be in flight into this CTA's shared memory when it exits
    while (landed_g < next_g) {
      if (mmw_mbar_test(gbar + (landed_g & (NB - 1)), (unsigned)((landed_g - first_g) / NB) & 1u)) landed_g++;
    }
    return;
  }

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

  if (role == 2 && CORE == 1) {
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
    //   ta  = imu * TABROW + (lane % TR) * 16   shared-memory address of this lane's copy of the interpolator row for the CURRENT mu
    constexpr unsigned RMASK = RING * RP - 4;
    constexpr int FAR = 1 << 22;  // rows beyond this go through the one-symbol path (iib stays inside 31 bits)
    const unsigned rep16 = (unsigned)(lane % TR) * 16u;
    const unsigned tak = (128u * TABROW - TABROW) * (unsigned)MMW_MAGIC_BITS + rep16 + tab_s;
    auto ta_of = [&](float m) {
      return ((unsigned)__float_as_int(__fmaf_rn(m, 128.0f, MMW_MAGIC)) & 0xffu) * (unsigned)TABROW + rep16 + tab_s;
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
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4+%5];" : "=f"(cf[4]), "=f"(cf[5]), "=f"(cf[6]), "=f"(cf[7]) : "r"(tad), "n"(TR * 16));
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
        const unsigned tan = ub * (unsigned)TABROW + (tb * (0u - 128u * TABROW) + tak);
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
          const float* tp = tab + (TABROW / 4) * mm_imu(mu);  // copy 0 of the row
#pragma unroll
          for (int i = 0; i < 8; i++) cf[i] = tp[(i >> 2) * (TR * 4) + (i & 3)];
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
      if (a.state_out2) {  // the same state where the right-hand time shard reads it (no copy on the serial chain)
        MMChanState* s2 = a.state_out2 + c;
        s2->mu = mu; s2->omega = omega; s2->last_sample = last;
        s2->next_abs = sp->next_abs; s2->clamped = sp->clamped; s2->overflow = sp->overflow;
      }
      a.counts[c] = oo;
    }
    return;
  }

  if (role == 2 && CORE == 2) {
    // ---------------------------------------------------------------------------- CORE, shortest dependent chain
    // Per symbol the recursion is one chain: (table row, 8 samples) -> interpolator -> timing error -> omega ->
    // clip -> mu -> (floor, rint) -> the next step's two shared-memory addresses.  Everything else is kept off it:
    //  * the LIVE state registers run ahead unconditionally; whether a step counted (input landed, queue slot free,
    //    ordinary forward step) only selects a COMMITTED copy of the state and the address of the queue store.  A
    //    lane that failed a step is dead until the end of the trip and then resumes from its committed copy;
    //  * mm_val = slice(last)*o - slice(o)*last = s * (|o| - |last|) with s = sign(o) * sign(last) (the products
    //    with +-1 are exact and IEEE rounding is symmetric in sign): ONE subtraction with |.| operand modifiers,
    //    and the sign goes into the two gains (gain ^ sign bits, one LOP3 each, computed beside the subtraction).
    //    -0.0 is the one value whose sign bit disagrees with slice(): such a step is left to the plain path;
    //  * 0.5 * x1 is exact, so omega_mid + 0.5 * x1 of branchless_clip is one FFMA, rounded once like the
    //    reference's addition;
    //  * the ring sits at shared-memory offset 0: slot address = (iib & RMASK), the base is an LDS immediate.
    const MMParams mp = a.p;
    constexpr int order = ORDER;
    const int max_out = valid ? a.max_out : 0;
    float mu = st.mu, omega = st.omega, last = st.last_sample;
    int ii = ii0, oo = 0, hi = ii0;
    int clamped = clamp0 ? 1 : 0;
    constexpr unsigned SIGN = 0x80000000u;
    const unsigned go_b = __float_as_uint(mp.gain_omega), gm_b = __float_as_uint(mp.gain_mu);  // gains are >= 0 (create)
    const float mid = mp.omega_mid, lim = mp.omega_relative_limit;
    const unsigned tab_s = (unsigned)__cvta_generic_to_shared(tab);
    constexpr unsigned RMASK = RING * RP - 4;
    constexpr int FAR = 1 << 22;  // rows beyond this go through the one-symbol path (iib stays inside 31 bits)
    const unsigned rep16 = (unsigned)(lane % TR) * 16u;
    const unsigned tak = (128u * TABROW - TABROW) * (unsigned)MMW_MAGIC_BITS + rep16 + tab_s;
    auto ta_of = [&](float m) {
      return min((unsigned)__float_as_int(__fmaf_rn(m, 128.0f, MMW_MAGIC)) & 0xffu, 128u) * (unsigned)TABROW + rep16 + tab_s;
    };
    int iib = ii * 256 + cl * 4;  // byte offset of (row ii, this lane) in an unbounded ring; & RMASK = the ring slot
    unsigned ta = ta_of(mu);      // shared-memory address of this lane's copy of the interpolator row of the current mu
    bool careful = __float_as_uint(last) == SIGN;  // the lane's next step goes through the one-symbol path below
    // words of the other warps, read half a trip ahead of their use (an older fill level is only conservative)
    int nf = mmw_ldv(pub_filled + cl);
    unsigned qw = mmw_ldq(q_lane + (unsigned)((oo + MMW_TRIP - 1) & (MMW_Q - 1)) * RP);
    float cf[8];
#pragma unroll
    for (int i = 0; i < 8; i++) cf[i] = 0.f;
    while (true) {
      // rows < pub_filled have landed; (ii <= fs8) == (ii + 8 <= pub_filled && ii < ni)
      const int fs8 = min(min(nf, ninput - 1) - 8, FAR);
      const int fs8b = fs8 * 256 + 255;  // ii <= fs8  <=>  iib <= fs8b  (cl * 4 < 256)
      // the post warp empties slots in order, so a free slot oo+TRIP-1 means oo..oo+TRIP-1 are free
      bool alive = !careful && oo + MMW_TRIP <= max_out && qw == MMW_EMPTY;
      bool redo = false;
      float c_mu = mu, c_om = omega, c_last = last;
      int c_iib = iib;
      unsigned c_ta = ta;
      unsigned gl_o = go_b ^ (__float_as_uint(last) & SIGN), gl_m = gm_b ^ (__float_as_uint(last) & SIGN);
      int tp = 1;  // ta is a table address (the step that produced it was a plain one)
      const int oo0 = oo;
#pragma unroll
      for (int k = 0; k < MMW_TRIP; k++) {
        const unsigned ra = (unsigned)iib & RMASK;
        float v[8];
        asm volatile(
            "{\n .reg .pred p;\n setp.ne.s32 p, %9, 0;\n"
            " @p ld.shared.v4.f32 {%0,%1,%2,%3}, [%8];\n"
            " @p ld.shared.v4.f32 {%4,%5,%6,%7}, [%8+%10];\n}"
            : "+f"(cf[0]), "+f"(cf[1]), "+f"(cf[2]), "+f"(cf[3]), "+f"(cf[4]), "+f"(cf[5]), "+f"(cf[6]), "+f"(cf[7])
            : "r"(ta), "r"(tp), "n"(TR * 16));
#pragma unroll
        for (int i = 0; i < 8; i++)
          v[i] = *reinterpret_cast<volatile float*>(reinterpret_cast<char*>(ring) + ra + i * RP);
        if (k == MMW_TRIP / 2) {
          nf = mmw_ldv(pub_filled + cl);
          qw = mmw_ldq(q_lane + (unsigned)((oo0 + 2 * MMW_TRIP - 1) & (MMW_Q - 1)) * RP);
        }
        const float o = mmse8(cf, v, order);
        const unsigned ob = __float_as_uint(o);
        const float D = __fsub_rn(fabsf(o), fabsf(last));
        const float dO = __fmul_rn(__uint_as_float(mmw_xor_and(gl_o, ob, SIGN)), D);
        const float dM = __fmul_rn(__uint_as_float(mmw_xor_and(gl_m, ob, SIGN)), D);
        const float x = __fsub_rn(__fadd_rn(omega, dO), mid);
        const float r = __fsub_rn(fabsf(__fadd_rn(x, lim)), fabsf(__fsub_rn(x, lim)));
        const float om = __fmaf_rn(0.5f, r, mid);
        const float m2 = __fadd_rn(__fadd_rn(mu, om), dM);
        // floor(m2) by adding 1.5*2^23 rounding towards minus infinity, rint(m2 * 128) by the FFMA (the product is
        // exact, the sum rounds once to nearest even): both exact for 0 <= m2 < 2^15.  The next mu is
        // m2 - floor(m2) (exact), so the next imu = rint(mu * 128) = rint(m2 * 128) - 128 * floor(m2)
        const unsigned tb = (unsigned)__float_as_int(__fadd_rd(m2, MMW_MAGIC));
        const unsigned ub = (unsigned)__float_as_int(__fmaf_rn(m2, 128.0f, MMW_MAGIC));
        const unsigned tan = ub * (unsigned)TABROW + (tb * (0u - 128u * TABROW) + tak);
        // ii + floor(m2), plus 256 * MAGIC_BITS = 2^30 (mod 2^32) per step: the live iib is biased by k * 2^30 after k
        // steps (0 again after a trip of 8); the ring mask drops the bias, the true value is recovered off the chain
        const int iibn = (int)(tb * 256u + (unsigned)iib);
        const int iibt = (int)((unsigned)iib - ((unsigned)(k & 3) << 30));          // true iib of this step
        const int iibnt = (int)((unsigned)iibn - ((unsigned)((k + 1) & 3) << 30));  // true iib of the next one
        const float mu2 = __fsub_rn(m2, __fsub_rn(__uint_as_float(tb), MMW_MAGIC));
        // forward step with the tricks valid: 0 <= m2 < 2^15, one unsigned compare on the bit pattern
        // (negative values, -0, NaN and Inf all have larger patterns)
        const bool plain = __float_as_uint(m2) < 0x47000000u;
        const bool good = plain && ob != SIGN;
        const bool ok = alive && iibt <= fs8b;  // the state is real and rows ii..ii+7 are in the ring
        redo = redo || (ok && !good);           // nothing committed: the step is redone the plain way
        alive = ok && good;
        unsigned obq = ob;
        if (obq == MMW_EMPTY) obq = 0x7fc00000u;
        mmw_stq(alive ? q_lane + (unsigned)(oo & (MMW_Q - 1)) * RP : dump_lane, obq);
        oo += alive ? 1 : 0;
        c_mu = alive ? mu2 : c_mu;
        c_om = alive ? om : c_om;
        c_last = alive ? o : c_last;
        c_iib = alive ? iibnt : c_iib;
        c_ta = alive ? tan : c_ta;
        mu = mu2; omega = om; last = o; iib = iibn; ta = tan;
        tp = plain ? 1 : 0;
        gl_o = go_b ^ (ob & SIGN);
        gl_m = gm_b ^ (ob & SIGN);
      }
      static_assert(MMW_TRIP % 4 == 0, "the bias of the live iib must be 0 again at the end of a trip");
      if (!alive) { mu = c_mu; omega = c_om; last = c_last; iib = c_iib; ta = c_ta; }
      careful = careful || redo;
      ii = iib >> 8;
      careful = careful || ii >= FAR;
      mmw_stv(pub_ii + cl, ii);
      // ---- one symbol the plain way: backward steps, the tail of a call, out-of-range mu, -0.0 --------
      // (the whole input is in global memory before the kernel starts; the ring is only a latency
      // optimisation, so this path depends on nobody)
      const bool act = oo < max_out && ii < ni;
      if (!__any_sync(0xffffffffu, act)) break;
      const bool tail = oo + MMW_TRIP > max_out;  // fewer than a trip's output slots left in this call
      if (__any_sync(0xffffffffu, act && (careful || tail))) {
        const unsigned qslot = q_lane + (unsigned)(oo & (MMW_Q - 1)) * RP;
        if (act && (careful || tail) && mmw_ldq(qslot) == MMW_EMPTY) {
          float v[8], c8[8];
#pragma unroll
          for (int i = 0; i < 8; i++) v[i] = __ldg(col + (size_t)(ii + i) * nchan);
          const float* tp8 = tab + (TABROW / 4) * mm_imu(mu);  // copy 0 of the row
#pragma unroll
          for (int i = 0; i < 8; i++) c8[i] = tp8[(i >> 2) * (TR * 4) + (i & 3)];
          const float o = mmse8(c8, v, order);
          unsigned obq = __float_as_uint(o);
          if (obq == MMW_EMPTY) obq = 0x7fc00000u;
          mmw_stq(qslot, obq);
          oo++;
          MMState s;
          s.mu = mu; s.omega = omega; s.last_sample = last;
          hi = max(hi, ii);
          ii += mm_update(s, mp, o);
          mu = s.mu; omega = s.omega; last = s.last_sample;
          if (ii < 0) { ii = 0; clamped++; }
          // the ring still holds rows >= hi - BACK (the loader never overwrites rows >= pub_ii - BACK
          // and every published position is <= hi); older rows keep coming from global memory
          careful = ii < max(hi - MMW_BACK, ii0) || __float_as_uint(last) == SIGN;
          iib = ii * 256 + cl * 4;
          ta = ta_of(mu);
          mmw_stv(pub_ii + cl, min(ii, hi));
        }
      }
      if (!alive) {  // the words read ahead were for a full trip
        nf = mmw_ldv(pub_filled + cl);
        qw = mmw_ldq(q_lane + (unsigned)((oo + MMW_TRIP - 1) & (MMW_Q - 1)) * RP);
      }
    }
    mmw_stv(pub_done + cl, oo + 1);
    if (valid) {
      MMChanState* sp = a.state + c;
      sp->mu = mu; sp->omega = omega; sp->last_sample = last;
      sp->next_abs = a.abs_row0 + ii;
      sp->clamped = st.clamped + clamped;
      sp->overflow = st.overflow + (ii < ni ? 1 : 0);
      if (a.state_out2) {  // the same state where the right-hand time shard reads it (no copy on the serial chain)
        MMChanState* s2 = a.state_out2 + c;
        s2->mu = mu; s2->omega = omega; s2->last_sample = last;
        s2->next_abs = sp->next_abs; s2->clamped = sp->clamped; s2->overflow = sp->overflow;
      }
      a.counts[c] = oo;
    }
    return;
  }

  if (role == 2 && CORE == 3) {
    // ---------------------------------------------------------------------------- CORE, fewest instructions
    // A lone warp per scheduler issues one instruction every other cycle at best (measured: ~2.2 "selected" cycles
    // per instruction, profiles/README.md), so a symbol costs 2 x instructions + the exposed part of the dependent
    // chain.  This build keeps the chain of CORE 2 and removes everything else it can from the trip:
    //  * no committed copy of the state: the registers of the trip start simply stay live, and a lane whose trip did
    //    not go through (input not landed yet, end of the block, an unusual step) replays its committed steps from
    //    there one at a time (rare);
    //  * one sticky predicate per lane (3 ISETP per step), a predicated queue store with an immediate slot offset
    //    (trips start on a multiple of 8 symbols), a predicated count;
    //  * the interpolator as packed FP32 (FMUL2 / FFMA2 / FADD2): 8 instructions instead of 15, same roundings;
    //  * floor / rint with the magic number 2^23 (0x4b000000): 0x4b000000 * 256 == 0 (mod 2^32), so
    //    iib + 256 * floor(m2) is ONE IMAD on the raw bit pattern.
    const MMParams mp = a.p;
    constexpr int order = ORDER;
    const int max_out = valid ? a.max_out : 0;
    float mu = st.mu, omega = st.omega, last = st.last_sample;
    int ii = ii0, oo = 0, hi = ii0;
    int clamped = clamp0 ? 1 : 0;
    constexpr unsigned SIGN = 0x80000000u;
    constexpr float MAGIC = 8388608.0f;  // 2^23: x + 2^23 leaves floor(x) / rint(x) in the mantissa for 0 <= x < 2^23
    constexpr unsigned MB = 0x4b000000u;
    const unsigned go_b = __float_as_uint(mp.gain_omega), gm_b = __float_as_uint(mp.gain_mu);  // gains are >= 0 (create)
    const float mid = mp.omega_mid, lim = mp.omega_relative_limit;
    const unsigned tab_s = (unsigned)__cvta_generic_to_shared(tab);
    const unsigned ring_s = (unsigned)__cvta_generic_to_shared(ring);
    constexpr unsigned RMASK = RING * RP - 4;
    constexpr int FAR = 1 << 22;  // rows beyond this go through the one-symbol path (iib stays inside 31 bits)
    // rows an ordinary trip consumes at most: 8 steps of floor(omega_max + 1 + gain_mu * |mm|) with |mm| up to ~6, + 8
    const int LA = min(MMW_TRIP * ((int)(mid + lim) + 3) + 8, RING - MMW_BACK - 8);
    const unsigned rep16 = (unsigned)(lane % TR) * 16u;
    const unsigned tak = (128u * TABROW - TABROW) * MB + rep16 + tab_s;
    const df_u64 ones = df_pack(a.one, a.one);
    auto ta_of = [&](float m) {
      return min((unsigned)__float_as_int(__fmaf_rn(m, 128.0f, MAGIC)) & 0xffu, 128u) * (unsigned)TABROW + rep16 + tab_s;
    };
    // one symbol the plain way (reference arithmetic as written: gr_math.cuh); input rows from the ring when they
    // are there, else from global memory (the whole input is there before the kernel starts)
    auto slow_step = [&](bool from_ring) -> float {
      float v[8], c8[8];
      if (from_ring) {
#pragma unroll
        for (int i = 0; i < 8; i++)
          v[i] = *reinterpret_cast<volatile float*>(ring + (size_t)((ii + i) & (RING - 1)) * MMW_CH + cl);
      } else {
#pragma unroll
        for (int i = 0; i < 8; i++) v[i] = __ldg(col + (size_t)(ii + i) * nchan);
      }
      const float* tp8 = tab + (TABROW / 4) * min(mm_imu(mu), 128);  // copy 0 of the row
#pragma unroll
      for (int i = 0; i < 8; i++) c8[i] = tp8[(i >> 2) * (TR * 4) + (i & 3)];
      const float o = mmse8(c8, v, order);
      MMState s;
      s.mu = mu; s.omega = omega; s.last_sample = last;
      hi = max(hi, ii);
      ii += mm_update(s, mp, o);
      mu = s.mu; omega = s.omega; last = s.last_sample;
      return o;
    };
    int iib = ii * 256 + cl * 4;  // byte offset of (row ii, this lane) in an unbounded ring; & RMASK = the ring slot
    unsigned ta = ta_of(mu);      // shared-memory address of this lane's copy of the interpolator row of the current mu
    bool careful = __float_as_uint(last) == SIGN;  // the lane's next step goes through the one-symbol path below
    // words of the other warps, read half a trip ahead of their use (an older fill level is only conservative)
    int nf = mmw_ldv(pub_filled + cl);
    unsigned qw = mmw_ldq(q_lane + (unsigned)((oo + MMW_TRIP - 1) & (MMW_Q - 1)) * RP);
    float cf[8];
#pragma unroll
    for (int i = 0; i < 8; i++) cf[i] = 0.f;
#ifdef MMW_STATS
    unsigned long long sc[6] = {0, 0, 0, 0, 0, 0};
    const long long t_begin = mmw_clk();
#endif
    while (true) {
#ifdef MMW_STATS
      const long long t_trip = mmw_clk();
#endif
      // rows < pub_filled have landed; (ii <= fs8) == (ii + 8 <= pub_filled && ii < ni)
      const int fs8 = min(min(nf, ninput - 1) - 8, FAR);
      const int fs8b = fs8 * 256 + 255;  // ii <= fs8  <=>  iib <= fs8b  (cl * 4 < 256)
      // the post warp empties slots in order, so a free slot oo+TRIP-1 means oo..oo+TRIP-1 are free
      // a trip is only started with the input of a whole ordinary trip in the ring (LA rows: a lane that stopped half
      // way would need up to 7 one-symbol steps to get back to a multiple of 8), or with everything landed
      const bool fed = (ii + LA <= nf) || nf >= ninput;
      const bool fast = !careful && (oo & (MMW_TRIP - 1)) == 0 && oo + MMW_TRIP <= max_out && qw == MMW_EMPTY && fed;
      bool alive = fast;
#ifdef MMW_STATS
      if (!careful && (oo & (MMW_TRIP - 1)) == 0 && oo + MMW_TRIP <= max_out && oo < max_out && ii < ni) {
        if (qw != MMW_EMPTY) atomicAdd(&mmw_stats[8], 1ull);
        else if (!fed) atomicAdd(&mmw_stats[9], 1ull);
      }
#endif
      const float s_mu = mu, s_om = omega, s_last = last;  // the trip-start state stays live: what a replay starts from
      const int s_iib = iib, s_oo = oo;
      volatile unsigned* qb = q + cl + (oo & (MMW_Q - 1)) * MMW_CH;
      unsigned gl_o = go_b ^ (__float_as_uint(last) & SIGN), gl_m = gm_b ^ (__float_as_uint(last) & SIGN);
#pragma unroll
      for (int k = 0; k < MMW_TRIP; k++) {
        const unsigned ra = (unsigned)iib & RMASK;
        float v[8];
        // the table row of a dead lane is not fetched (its address need not be one)
        asm volatile(
            "{\n .reg .pred p;\n setp.ne.s32 p, %9, 0;\n"
            " @p ld.shared.v4.f32 {%0,%1,%2,%3}, [%8];\n"
            " @p ld.shared.v4.f32 {%4,%5,%6,%7}, [%8+%10];\n}"
            : "+f"(cf[0]), "+f"(cf[1]), "+f"(cf[2]), "+f"(cf[3]), "+f"(cf[4]), "+f"(cf[5]), "+f"(cf[6]), "+f"(cf[7])
            : "r"(ta), "r"((int)alive), "n"(TR * 16));
        // plain (weak) loads: ptxas issues them back to back; every address is a function of this step's chain, so
        // they cannot move above the fill level (nf) that vouches for them, read half a trip earlier
#pragma unroll
        for (int i = 0; i < 8; i++) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v[i]) : "r"(ring_s + ra + i * RP));
        if (k == MMW_TRIP / 2) {
          nf = mmw_ldv(pub_filled + cl);
          qw = mmw_ldq(q_lane + (unsigned)((s_oo + 2 * MMW_TRIP - 1) & (MMW_Q - 1)) * RP);
        }
        float o;
        if (order == GR_ORDER_SSE) {
          // q_i = c_i v_i + c_{i+4} v_{i+4} as pairs (q0,q1), (q2,q3); (q0+q2, q1+q3); one scalar addition
          const df_u64 p01 = df_mul2(df_pack(cf[0], cf[1]), df_pack(v[0], v[1]));
          const df_u64 p23 = df_mul2(df_pack(cf[2], cf[3]), df_pack(v[2], v[3]));
          const df_u64 p45 = df_mul2(df_pack(cf[4], cf[5]), df_pack(v[4], v[5]));
          const df_u64 p67 = df_mul2(df_pack(cf[6], cf[7]), df_pack(v[6], v[7]));
          const df_u64 q01 = df_acc2(p45, ones, p01), q23 = df_acc2(p67, ones, p23);
          float e0, e1;
          df_unpack(df_add2(q01, q23), e0, e1);
          o = __fadd_rn(e0, e1);
        } else {
          o = mmse8(cf, v, order);
        }
        const unsigned ob = __float_as_uint(o);
        const float D = __fsub_rn(fabsf(o), fabsf(last));
        const float dO = __fmul_rn(__uint_as_float(mmw_xor_and(gl_o, ob, SIGN)), D);
        const float dM = __fmul_rn(__uint_as_float(mmw_xor_and(gl_m, ob, SIGN)), D);
        const float x = __fsub_rn(__fadd_rn(omega, dO), mid);
        const float r = __fsub_rn(fabsf(__fadd_rn(x, lim)), fabsf(__fsub_rn(x, lim)));
        const float om = __fmaf_rn(0.5f, r, mid);
        const float m2 = __fadd_rn(__fadd_rn(mu, om), dM);
        const float tbf = __fadd_rd(m2, MAGIC);                         // 2^23 + floor(m2)
        const unsigned tb = __float_as_uint(tbf);
        const unsigned ub = __float_as_uint(__fmaf_rn(m2, 128.0f, MAGIC));  // 2^23 + rint(128 m2)
        // next imu = rint(128 m2) - 128 floor(m2) (the next mu is m2 - floor(m2), exactly)
        const unsigned tan = ub * (unsigned)TABROW + (tb * (0u - 128u * TABROW) + tak);
        int iibn;                                                       // MB * 256 == 0 (mod 2^32)
        asm("mad.lo.u32 %0, %1, 256, %2;" : "=r"(iibn) : "r"(tb), "r"(iib));  // (kept as ONE IMAD off tb)
        const float mu2 = __fsub_rn(m2, __fsub_rn(tbf, MAGIC));
        // the step counts if the lane was alive, rows ii..ii+7 had landed, o is not -0.0 (the one value whose sign
        // bit disagrees with slice()) and m2 is an ordinary forward step: 0 <= m2 < 2^15, one unsigned compare on
        // the bit pattern (negative values, -0, NaN and Inf all have larger patterns)
        alive = alive && iib <= fs8b && ob != SIGN && __float_as_uint(m2) < 0x47000000u;
        if (alive) qb[k * MMW_CH] = ob;  // predicated STS, immediate slot offset
        oo += alive ? 1 : 0;
        mu = mu2; omega = om; last = o; iib = iibn; ta = tan;
        gl_o = go_b ^ (ob & SIGN);
        gl_m = gm_b ^ (ob & SIGN);
      }
#ifdef MMW_STATS
      if (fast && !alive && ii < ni) atomicAdd(&mmw_stats[10], 1ull);
#endif
      // The ordinary trip end: every lane went through, at least one has input and room left, nobody is due for a
      // one-symbol step -- ONE warp reduction and one uniform branch (three votes and their divergence brackets
      // cost ~20 cycles per symbol).  oo stays a multiple of MMW_TRIP on this path.
      {
        const int iia = iib >> 8;
        const bool more = oo + MMW_TRIP <= max_out && iia < ni;
        const unsigned wf = __reduce_or_sync(0xffffffffu, (alive ? 0u : 1u) | (more ? 2u : 0u) | (iia >= FAR ? 1u : 0u));
        if (wf == 2u) {
          ii = iia;
          mmw_stv(pub_ii + cl, (LD == 1 && !more) ? ninput : iia);
#ifdef MMW_STATS
          { const long long t_mid = mmw_clk(); sc[0] += t_mid - t_trip; sc[1]++; }
#endif
          continue;
        }
      }
      if (__any_sync(0xffffffffu, !alive)) {
        if (!alive) {
          // back to the trip start, then the committed steps again, one at a time (their rows are still in the
          // ring: nothing past the trip start has been published to the loader)
          const int nc = oo - s_oo;
          mu = s_mu; omega = s_om; last = s_last;
          ii = s_iib >> 8;
          for (int j = 0; j < nc; j++) slow_step(true);
          // input was there and the lane was running: the step itself was the unusual one
          careful = careful || (fast && ii <= fs8) || __float_as_uint(last) == SIGN;
          iib = ii * 256 + cl * 4;
          ta = ta_of(mu);
        }
      }
      ii = iib >> 8;
      careful = careful || ii >= FAR;
      const bool act = oo < max_out && ii < ni;
      mmw_stv(pub_ii + cl, (LD == 1 && !act) ? ninput : ii);  // a finished lane does not hold the shared fill level back
#ifdef MMW_STATS
      const long long t_mid = mmw_clk();
      { const int w = __all_sync(0xffffffffu, alive) ? 0 : 2; sc[w] += t_mid - t_trip; sc[w + 1]++; }
#endif
      // ---- one symbol the plain way: backward steps, the tail of a call, out-of-range mu, -0.0, and the symbols
      // that bring a lane back to a multiple of MMW_TRIP after a trip that stopped half way
      if (!__any_sync(0xffffffffu, act)) break;
      const bool single = careful || (oo & (MMW_TRIP - 1)) != 0 || oo + MMW_TRIP > max_out;
      if (__any_sync(0xffffffffu, act && single)) {
        const unsigned qslot = q_lane + (unsigned)(oo & (MMW_Q - 1)) * RP;
        if (act && single && mmw_ldq(qslot) == MMW_EMPTY) {
          const int nfl = mmw_ldv(pub_filled + cl);
          const float o = slow_step(ii + 8 <= nfl && ii >= max(hi - MMW_BACK, ii0));
          unsigned obq = __float_as_uint(o);
          if (obq == MMW_EMPTY) obq = 0x7fc00000u;
          mmw_stq(qslot, obq);
          oo++;
          if (ii < 0) { ii = 0; clamped++; }
          // the ring still holds rows >= hi - BACK (the loader never overwrites rows >= pub_ii - BACK
          // and every published position is <= hi); older rows keep coming from global memory
          careful = ii < max(hi - MMW_BACK, ii0) || __float_as_uint(last) == SIGN;
          iib = ii * 256 + cl * 4;
          ta = ta_of(mu);
          mmw_stv(pub_ii + cl, min(ii, hi));
        }
      }
      if (!alive) {  // the words read ahead were for a full trip
        nf = mmw_ldv(pub_filled + cl);
        qw = mmw_ldq(q_lane + (unsigned)((oo + MMW_TRIP - 1) & (MMW_Q - 1)) * RP);
      }
#ifdef MMW_STATS
      { const long long t_end = mmw_clk(); if (t_end - t_mid > 60) { sc[4] += t_end - t_mid; sc[5]++; } }
#endif
    }
#ifdef MMW_STATS
    if (lane == 0) {
      for (int i = 0; i < 6; i++) atomicAdd(&mmw_stats[i], sc[i]);
      atomicAdd(&mmw_stats[6], (unsigned long long)(mmw_clk() - t_begin));
      atomicAdd(&mmw_stats[7], 1ull);
    }
#endif
    mmw_stv(pub_done + cl, oo + 1);
    if (valid) {
      MMChanState* sp = a.state + c;
      sp->mu = mu; sp->omega = omega; sp->last_sample = last;
      sp->next_abs = a.abs_row0 + ii;
      sp->clamped = st.clamped + clamped;
      sp->overflow = st.overflow + (ii < ni ? 1 : 0);
      if (a.state_out2) {  // the same state where the right-hand time shard reads it (no copy on the serial chain)
        MMChanState* s2 = a.state_out2 + c;
        s2->mu = mu; s2->omega = omega; s2->last_sample = last;
        s2->next_abs = sp->next_abs; s2->clamped = sp->clamped; s2->overflow = sp->overflow;
      }
      a.counts[c] = oo;
    }
    return;
  }

  // ------------------------------------------------------------------------------ POST
  mmw_post_warp<MMW_Q>(a, st, valid, c, cl, q_lane, pub_done, smap);
}

}  // namespace grb
