// digital_clock_recovery_mm_ff::general_work, batched (digital_clock_recovery_mm_ff.cc:102-139 +
// gri_mmse_fir_interpolator.cc:61-71) -- the build for a kernel that has the SMs to itself (time shards, the
// stand-alone block, a single-GPU chain whose tail no longer hides under the front).  Same roles, queue, post warps
// and one-symbol fallback paths as mm_ws_kernel (kernel_mm.cuh); what changes is how the input reaches the loop.
//
// What the measurements said (profiles/README.md, round 2): a lone warp issues one instruction per cycle and a
// dependent FP32 / integer instruction costs 4 cycles (5 across pipes), a shared-memory load ~23; ptxas spaces
// consecutive LDS of a warp 4 cycles apart; and per symbol the loop is ONE chain
//     (2 addresses) -> LDS -> interpolator (4 levels) -> timing error -> omega -> clip (4) -> mu (2) -> floor/rint
// so a symbol costs  [LDS issue + LDS latency]  +  ~17 dependent levels x 4.  Ten LDS per symbol (8 samples + the
// two halves of the interpolator row) were 40 cycles of issue alone, the plain interpolator table cost 13.6
// shared-memory wavefronts per half row (every lane has its own mu), and the per-lane cp.async loader spent 18 % of
// the shared-memory pipe and most of the issue slots of the schedulers it shares with the post warps.  Hence:
//
//  * QUAD RING.  Shared memory holds, per lane, quad[s] = {in[s], in[s+1], in[s+2], in[s+3]} (16 bytes, always
//    aligned): the 8 samples of a step are TWO LDS.128 (quad[ii], quad[ii+4]), conflict free whatever the lanes'
//    positions (lane l always touches banks 4l..4l+3).  Four LDS per symbol instead of ten.
//  * The interpolator table is replicated 8 times ([row][half][copy][4]): lane l reads copy l % 8, exactly 4
//    wavefronts per LDS.128.
//  * ASYNCHRONOUS STAGING + CONVERTER WARPS.  Warp 0 streams the CTA's 64-channel row segments (256 bytes per row)
//    into a staging ring with 16-byte cp.async (LDGSTS.128: two rows per warp instruction; cp.async.bulk was
//    measured at > 100 cycles per 256-byte copy, 0.6 TB/s for the whole chip), 16 rows per mbarrier
//    (cp.async.mbarrier.arrive); warps 0 and 1 wait on the barriers and turn rows into quads (one LDS.32 + one STS.128 per row and lane, three values of history in registers).  HBM
//    latency is absorbed by the staging depth, not by registers or by the loop.  One fill level per warp: the 32
//    channels of a warp start every call within a few rows of each other and omega is clipped to
//    +- omega_relative_limit, so they walk their columns at (nearly) the same pace; a lane that runs far ahead of
//    the slowest one waits for it (correct for any spread, fast for the ordinary one).
//  * The core is CORE 3 of kernel_mm.cuh (fewest instructions: live registers run ahead, a lane whose trip did
//    not go through replays its committed steps from the trip-start registers; packed FP32 interpolator; floor and
//    rint through the magic number 2^23 so that the next addresses are single IMADs on raw bit patterns).
#pragma once
#include "kernel_mm.cuh"

namespace grb {

#define MMQ_RING 128   // quad slots per lane (rows of look-ahead + look-back)
#define MMQ_SR 128     // staging rows (TMA destination)
#define MMQ_G 16       // rows per mbarrier
#define MMQ_TR 8       // copies of the interpolator table
#define MMQ_Q 64       // soft-symbol queue depth per lane

__device__ __forceinline__ bool mmw_mbar_try(uint64_t* bar, unsigned parity) {  // suspends up to a hardware time limit
  unsigned ok;
  asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}

static inline size_t mm_quad_smem_bytes() {
  return (size_t)(MMQ_RING + 4) * MMW_CH * 16 + (size_t)MMQ_SR * MMW_CH * 4 + MMQ_Q * MMW_CH * 4 + 4 * MMW_CH * 4 + 256 +
         129 * (size_t)(2 * MMQ_TR * 16) + (size_t)(MMQ_SR / MMQ_G) * 8 + 64;
}

template <int ORDER, int NREG>
__global__ void __maxnreg__(NREG) mm_quad_kernel(const MMArgs a) {
  extern __shared__ __align__(16) float mmw_smem[];
  constexpr int RING = MMQ_RING, SR = MMQ_SR, TR = MMQ_TR, G = MMQ_G, NB = MMQ_SR / MMQ_G;
  constexpr int TABROW = 2 * TR * 16;
  constexpr unsigned QP = MMW_CH * 16;  // bytes per quad slot (64 lanes x 16)
  constexpr unsigned RP = MMW_CH * 4;   // queue / staging row pitch in bytes
  // the quad ring comes first: its address is a link-time constant folded into the LDS immediates
  float4* quad = reinterpret_cast<float4*>(mmw_smem);                          // [RING + 4][64]; slots RING..RING+3 mirror 0..3
  float* stage = reinterpret_cast<float*>(quad + (RING + 4) * MMW_CH);         // [SR][64]
  unsigned* q = reinterpret_cast<unsigned*>(stage + SR * MMW_CH);              // [MMQ_Q][64]
  int* pub_ii = reinterpret_cast<int*>(q + MMQ_Q * MMW_CH);                    // [64] core -> converter: current input position
  int* pub_filled = pub_ii + MMW_CH;                                           // [64] converter -> core: rows < this are in the ring
  int* pub_done = pub_filled + MMW_CH;                                         // [64] core -> others: symbols produced + 1
  int* misc = pub_done + MMW_CH;                                               // [0], [1]: groups converted by warp 0 / 1
  unsigned char* smap = reinterpret_cast<unsigned char*>(misc + MMW_CH);       // [256] gr_map_bb table
  float* tab = reinterpret_cast<float*>(smap + 256);                           // [129][2][TR][4]
  uint64_t* gbar = reinterpret_cast<uint64_t*>(tab + 129 * (TABROW / 4));      // [NB]

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int role = warp < 2 ? 0 : warp < 4 ? 2 : 1;  // 0 converter (+ TMA issue on warp 0), 1 post, 2 core
  const int cl = (warp & 1) * 32 + lane;
  const int c = blockIdx.x * MMW_CH + cl;
  const bool valid = c < a.nchan;
  for (int e = threadIdx.x; e < 129 * 8; e += MMW_THREADS) {
    const float cv = __ldg(a.mmse_eff + e);
    const int j = e & 7, row = e >> 3;
#pragma unroll
    for (int rep = 0; rep < TR; rep++) tab[row * (TABROW / 4) + (j >> 2) * (TR * 4) + rep * 4 + (j & 3)] = cv;
  }
  for (int i = threadIdx.x; i < 256; i += MMW_THREADS) smap[i] = a.corr.map[i];
  for (int i = threadIdx.x; i < MMQ_Q * MMW_CH; i += MMW_THREADS) q[i] = MMW_EMPTY;

  const int ninput = (int)a.ninput;
  const int ni = ninput - 8;  // :112
  MMChanState st;
  st.mu = 0.f; st.omega = 0.f; st.last_sample = 0.f; st.slicer_avg = 0.f; st.next_abs = a.abs_row0; st.clamped = 0; st.overflow = 0;
  if (valid) st = (a.state_in ? a.state_in : a.state)[c];
  int ii0 = (int)(st.next_abs - a.abs_row0);  // may be > 0: samples already consumed
  const bool clamp0 = ii0 < 0;                // (see mm_ws_kernel: a step before the first buffered row is clamped and counted)
  if (clamp0) ii0 = 0;
  if (role == 2) {
    pub_ii[cl] = valid ? ii0 : ninput;  // a lane without a channel does not hold the shared fill level back
    pub_filled[cl] = 0;
    pub_done[cl] = valid ? 0 : 1;
  }
  if (threadIdx.x < 2) misc[threadIdx.x] = 0;
  if (threadIdx.x < NB) mbar_init(gbar + threadIdx.x, 32);  // every lane of warp 0 arrives once its copies of the group have landed
  mbar_init_fence();
  __syncthreads();

  const size_t nchan = (size_t)a.nchan;
  const float* __restrict__ col = a.in + (valid ? c : 0);
  const unsigned q_lane = (unsigned)__cvta_generic_to_shared(q + cl);

  if (role == 0) {
    // ------------------------------------------------------------------- TMA ISSUE (warp 0) + CONVERTER (warps 0, 1)
    // Rows are handled in groups of MMQ_G; group g lives in staging slot g % NB and completes on gbar[g % NB].
    // Per group a converter pays one mbarrier wait, one fence and a handful of flow-control reads, per row one
    // LDS.32 + one STS.128: a fraction of what the loop takes to consume the rows.
    const int w = warp;  // which half of the CTA's channels this converter serves
    int gmin = min(mmw_ldv(pub_ii + lane), mmw_ldv(pub_ii + lane + 32));
    gmin = __reduce_min_sync(0xffffffffu, gmin);
    const int first_g = max(gmin - MMW_BACK, 0) / G;  // the group of the first row anybody needs
    const int last_g = (ninput + G - 1) / G;            // one past the last group
    const int c0 = blockIdx.x * MMW_CH;
    const unsigned row_bytes = (unsigned)min(MMW_CH, a.nchan - c0) * 4u;
    const float* base = a.in + c0;
    int next_g = first_g;  // next group to issue (warp 0)
    int g = first_g;       // next group to convert
    float h1 = 0.f, h2 = 0.f, h3 = 0.f;  // in[r-3], in[r-2], in[r-1] of this lane's column
    const unsigned stage_lane = (unsigned)__cvta_generic_to_shared(stage + cl);
    const unsigned stage_s = (unsigned)__cvta_generic_to_shared(stage);
    const unsigned quad_lane = (unsigned)__cvta_generic_to_shared(quad + cl);
    // A group's quads are published one step late: the CTA fence (MEMBAR) waits for every shared-memory store in
    // flight, 36 cycles + ~25 per store, i.e. ~450 cycles right behind the 16 STS.128 of a group, but ~36 once they
    // have drained -- which they have by the time the next group's barrier wait returns (or before any sleep).
    bool pending = false;
    auto publish = [&]() {
      __threadfence_block();
      mmw_stv(pub_filled + cl, min(g * G, ninput));  // rows < this have landed (quads up to row - 4 are complete)
      if (lane == 0) mmw_stv(misc + w, g);
      pending = false;
    };
    auto all_done = [&]() {  // every channel of the CTA is done (output capacity reached before the input ran out)
      const bool fin = mmw_ldv(pub_done + lane) != 0 && mmw_ldv(pub_done + lane + 32) != 0;
      return __all_sync(0xffffffffu, fin) != 0;
    };
    while (g < last_g) {
      // ---- issue whatever the staging ring has room for (both converters must be done with the slot) ----------
      if (w == 0) {
        const int done_g = max(first_g, min(g, mmw_ldv(misc + 1)));  // (misc starts at 0, the groups at first_g)
        while (next_g < last_g && next_g - done_g < NB) {
          // 16-byte asynchronous copies (LDGSTS.128): a warp instruction moves two 256-byte row segments, a group
          // of G rows is G/2 instructions; every lane then makes the group's mbarrier track its copies
          const int slot_g = next_g % NB;
          const int r0 = next_g * G;
          const unsigned dst0 = stage_s + (unsigned)(slot_g * G + (lane >> 4)) * RP + (unsigned)(lane & 15) * 16u;
          const float* src0 = base + (size_t)(r0 + (lane >> 4)) * nchan + (lane & 15) * 4;
          const bool chunk_ok = (unsigned)(lane & 15) * 16u < row_bytes;
#pragma unroll
          for (int i = 0; i < G / 2; i++) {
            if (chunk_ok && r0 + 2 * i + (lane >> 4) < ninput)
              asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst0 + (unsigned)(2 * i) * RP), "l"(src0 + (size_t)(2 * i) * nchan) : "memory");
          }
          asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(gbar + slot_g)) : "memory");
          next_g++;
        }
      }
      // ---- room in the quad ring: the slot of row s = r - 3 overwrites row s - RING, which must lie more than
      //      BACK rows behind the slowest lane of this warp (finished lanes publish ninput) --------------------------
      const int r_last = g * G + G - 1;  // (a last, partial group still writes all its G slots)
      const int cur = __reduce_min_sync(0xffffffffu, mmw_ldv(pub_ii + cl));
      if (r_last - 3 - RING >= cur - MMW_BACK) {
#ifdef MMW_STATS
        if (lane == 0) atomicAdd(&mmw_stats[4], 1ull);
#endif
        if (pending) publish();
        if (all_done()) break;
        __nanosleep(150);
        continue;
      }
      // ---- the group must have landed (try_wait suspends the warp for a while: nothing to issue meanwhile, staging
      //      slots only become free when a conversion completes) ---------------------------------------------------
      if (!mmw_mbar_try(gbar + g % NB, (unsigned)((g - first_g) / NB) & 1u)) {
#ifdef MMW_STATS
        if (lane == 0) atomicAdd(&mmw_stats[5], 1ull);
#endif
        if (pending) publish();
        if (all_done()) break;
        continue;
      }
      if (pending) publish();
      const int r0 = g * G;
      const unsigned sbase = stage_lane + (unsigned)((g % NB) * G) * RP;
#pragma unroll
      for (int jb = 0; jb < G; jb += 8) {
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; j++) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v[j]) : "r"(sbase + (jb + j) * RP));
#pragma unroll
        for (int j = 0; j < 8; j++) {
          // quad[s] = {in[s], in[s+1], in[s+2], in[s+3]} is complete with row s + 3 = r0 + jb + j
          const unsigned slot = (unsigned)(r0 + jb + j - 3) & (RING - 1);
          asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(quad_lane + slot * QP), "f"(h1), "f"(h2), "f"(h3), "f"(v[j]) : "memory");
          h1 = h2; h2 = h3; h3 = v[j];
        }
      }
      // slots RING .. RING+3 mirror slots 0 .. 3 (the second LDS.128 of a step never wraps)
      if (((unsigned)(r0 - 3) & (RING - 1)) > ((unsigned)(r0 + G - 4) & (RING - 1)) || ((unsigned)(r0 - 3) & (RING - 1)) < 4) {
#pragma unroll
        for (int j = 0; j < 4; j++) {
          float x0, x1, x2, x3;
          asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x0), "=f"(x1), "=f"(x2), "=f"(x3) : "r"(quad_lane + j * QP) : "memory");
          asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(quad_lane + (RING + j) * QP), "f"(x0), "f"(x1), "f"(x2), "f"(x3) : "memory");
        }
      }
      g++;
      pending = true;
    }
    if (pending) publish();
    // nothing may still be in flight into this CTA's shared memory when it exits
    if (w == 0) {
      for (int k = max(g, first_g); k < next_g; k++) {
        while (!mmw_mbar_try(gbar + k % NB, (unsigned)((k - first_g) / NB) & 1u)) {}
      }
    }
    return;
  }

  if (role == 2) {
    // ---------------------------------------------------------------------------- CORE (see CORE 3 of mm_ws_kernel)
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
    const unsigned quad_s = (unsigned)__cvta_generic_to_shared(quad);
    constexpr unsigned RMASK = RING * QP - 16;
    constexpr int FAR = 1 << 20;  // rows beyond this go through the one-symbol path (iib stays inside 31 bits)
    // rows an ordinary trip consumes: 8 steps of ~omega (the sum of 8 steps hardly moves), + 8 of interpolator
    // look-ahead, + 4; a trip that needs more stops half way and its lane takes the one-symbol path (rare)
    const int LA = min(MMW_TRIP * ((int)(mid + lim) + 1) + 8 + 4, RING - MMW_BACK - 16);
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
        const unsigned p0 = quad_s + (unsigned)(ii & (RING - 1)) * QP + (unsigned)cl * 16u;
        asm volatile("ld.volatile.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(p0) : "memory");
        asm volatile("ld.volatile.shared.v4.f32 {%0,%1,%2,%3}, [%4+%5];" : "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "r"(p0), "n"(4 * QP) : "memory");
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
    int iib = ii * (int)QP + cl * 16;  // byte offset of (slot ii, this lane) in an unbounded ring; & RMASK = the ring slot
    unsigned ta = ta_of(mu);           // shared-memory address of this lane's copy of the interpolator row of the current mu
    bool careful = __float_as_uint(last) == SIGN;  // the lane's next step goes through the one-symbol path below
    int nf = mmw_ldv(pub_filled + cl);
    unsigned qw = mmw_ldq(q_lane + (unsigned)((oo + MMW_TRIP - 1) & (MMQ_Q - 1)) * RP);
    float cf[8], v[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { cf[i] = 0.f; v[i] = 0.f; }
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
      const int fs8b = fs8 * (int)QP + (int)QP - 1;  // ii <= fs8  <=>  iib <= fs8b  (cl * 16 < QP)
      // a trip is only started with the input of a whole ordinary trip in the ring, or with everything landed
      const bool fed = (ii + LA <= nf) || nf >= ninput;
      // the post warp empties slots in order, so a free slot oo+TRIP-1 means oo..oo+TRIP-1 are free
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
      volatile unsigned* qb = q + cl + (oo & (MMQ_Q - 1)) * MMW_CH;
      unsigned gl_o = go_b ^ (__float_as_uint(last) & SIGN), gl_m = gm_b ^ (__float_as_uint(last) & SIGN);
#pragma unroll
      for (int k = 0; k < MMW_TRIP; k++) {
        const unsigned ra = quad_s + ((unsigned)iib & RMASK);
        // the table row of a dead lane is not fetched (its address need not be one); the ring address always is one
        asm volatile(
            "{\n .reg .pred p;\n setp.ne.s32 p, %9, 0;\n"
            " @p ld.shared.v4.f32 {%0,%1,%2,%3}, [%8];\n"
            " @p ld.shared.v4.f32 {%4,%5,%6,%7}, [%8+%10];\n}"
            : "+f"(cf[0]), "+f"(cf[1]), "+f"(cf[2]), "+f"(cf[3]), "+f"(cf[4]), "+f"(cf[5]), "+f"(cf[6]), "+f"(cf[7])
            : "r"(ta), "r"((int)alive), "n"(TR * 16));
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(ra));
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4+%5];" : "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "r"(ra), "n"(4 * QP));
        if (k == MMW_TRIP - 2) {  // as late as its latency stays hidden: a fill level a quarter trip old
          nf = mmw_ldv(pub_filled + cl);
          qw = mmw_ldq(q_lane + (unsigned)((s_oo + 2 * MMW_TRIP - 1) & (MMQ_Q - 1)) * RP);
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
        int iibn;                                                       // MB * QP == 0 (mod 2^32)
        asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(iibn) : "r"(tb), "n"(QP), "r"(iib));
        const float mu2 = __fsub_rn(m2, __fsub_rn(tbf, MAGIC));
        alive = alive && iib <= fs8b && ob != SIGN && __float_as_uint(m2) < 0x47000000u;
        if (alive) qb[k * MMW_CH] = ob;  // predicated STS, immediate slot offset
        oo += alive ? 1 : 0;
        mu = mu2; omega = om; last = o; iib = iibn; ta = tan;
        gl_o = go_b ^ (ob & SIGN);
        gl_m = gm_b ^ (ob & SIGN);
      }
      // The ordinary trip end: every lane went through, at least one has input and room left -- ONE warp reduction
      // and one uniform branch.  oo stays a multiple of MMW_TRIP on this path.
      {
        const int iia = (int)((unsigned)iib / QP);
        const bool more = oo + MMW_TRIP <= max_out && iia < ni;
        const unsigned wf = __reduce_or_sync(0xffffffffu, (alive ? 0u : 1u) | (more ? 2u : 0u) | (iia >= FAR ? 1u : 0u));
        if (wf == 2u) {
          ii = iia;
          mmw_stv(pub_ii + cl, more ? iia : ninput);
#ifdef MMW_STATS
          sc[0] += mmw_clk() - t_trip; sc[1]++;
#endif
          continue;
        }
      }
#ifdef MMW_STATS
      if (fast && !alive && ii < ni) atomicAdd(&mmw_stats[10], 1ull);
#endif
      if (!alive) {
        // back to the trip start, then the committed steps again, one at a time (their rows are still in the ring:
        // nothing past the trip start has been published)
        const int nc = oo - s_oo;
        mu = s_mu; omega = s_om; last = s_last;
        ii = (int)((unsigned)s_iib / QP);
        for (int j = 0; j < nc; j++) slow_step(true);
        // input was there and the lane was running: the step itself was the unusual one
        careful = careful || (fast && ii <= fs8) || __float_as_uint(last) == SIGN;
        iib = ii * (int)QP + cl * 16;
        ta = ta_of(mu);
      }
      ii = (int)((unsigned)iib / QP);
      careful = careful || ii >= FAR;
      const bool act = oo < max_out && ii < ni;
      mmw_stv(pub_ii + cl, act ? ii : ninput);  // a finished lane does not hold the shared fill level back
      // ---- one symbol the plain way: backward steps, the tail of a call, out-of-range mu, -0.0, and the symbols
      // that bring a lane back to a multiple of MMW_TRIP after a trip that stopped half way
      if (!__any_sync(0xffffffffu, act)) break;
      const bool single = careful || (oo & (MMW_TRIP - 1)) != 0 || oo + MMW_TRIP > max_out;
      if (__any_sync(0xffffffffu, act && single)) {
        const unsigned qslot = q_lane + (unsigned)(oo & (MMQ_Q - 1)) * RP;
        if (act && single && mmw_ldq(qslot) == MMW_EMPTY) {
          const int nfl = mmw_ldv(pub_filled + cl);
          const float o = slow_step(ii + 8 <= nfl && ii >= max(hi - MMW_BACK, ii0));
          unsigned obq = __float_as_uint(o);
          if (obq == MMW_EMPTY) obq = 0x7fc00000u;
          mmw_stq(qslot, obq);
          oo++;
          if (ii < 0) { ii = 0; clamped++; }
          // the ring still holds rows >= hi - BACK (the converter never overwrites rows >= pub_ii - BACK and every
          // published position is <= hi); older rows keep coming from global memory
          careful = ii < max(hi - MMW_BACK, ii0) || __float_as_uint(last) == SIGN;
          iib = ii * (int)QP + cl * 16;
          ta = ta_of(mu);
          mmw_stv(pub_ii + cl, min(ii, hi));
        }
      }
      if (!alive) {  // the words read ahead were for a full trip
        nf = mmw_ldv(pub_filled + cl);
        qw = mmw_ldq(q_lane + (unsigned)((oo + MMW_TRIP - 1) & (MMQ_Q - 1)) * RP);
      }
#ifdef MMW_STATS
      sc[2] += mmw_clk() - t_trip; sc[3]++;
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
  mmw_post_warp<MMQ_Q>(a, st, valid, c, cl, q_lane, pub_done, smap);
}

}  // namespace grb
