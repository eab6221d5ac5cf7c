// Per-thread arithmetic of the demod tail, written once for device code (and compiled for the
// host by tests/emul, which replays single "threads" on the CPU to check the logic without a GPU;
// the product never runs these on the host).
//
// Everything here that feeds a slicer decision reproduces the reference's float arithmetic
// operation by operation: separate IEEE multiply and add (never an FMA), the reference's
// summation trees and its two lookup tables (SURVEY.md section 7 "Bit-exact demod").
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#define GR_HD __host__ __device__ __forceinline__

#ifdef __CUDA_ARCH__
#define GR_FMUL(a, b) __fmul_rn((a), (b))
#define GR_FADD(a, b) __fadd_rn((a), (b))
#define GR_FSUB(a, b) __fsub_rn((a), (b))
#define GR_FDIV(a, b) __fdiv_rn((a), (b))
#else  // host emulation: the TU is compiled with -ffp-contract=off
#define GR_FMUL(a, b) ((float)(a) * (float)(b))
#define GR_FADD(a, b) ((float)(a) + (float)(b))
#define GR_FSUB(a, b) ((float)(a) - (float)(b))
#define GR_FDIV(a, b) ((float)(a) / (float)(b))
#endif

#define GR_ORDER_GENERIC 0
#define GR_ORDER_SSE 1

namespace grb {

// ---- gr_fast_atan2f (gnuradio-core/src/lib/general/gr_fast_atan2f.cc:125-198) ---------------
// `table` = the reference's 257-entry arctangent table (build/generated/gr_tables.h).
// Branch free (a warp's lanes sit in different octants: the reference's if/else tree would make the
// warp walk every path).  Same operations on the same operands as the reference, hence the same
// bits: every octant's result is ONE rounded addition (+-K) + (+-base) with K = pi or pi/2
// (a - b == a + (-b) exactly), except the first octant pair, which returns +-base untouched.
GR_HD unsigned gr_f2u(float f) {
#ifdef __CUDA_ARCH__
  return __float_as_uint(f);
#else
  unsigned u;
  memcpy(&u, &f, 4);
  return u;
#endif
}
GR_HD float gr_u2f(unsigned u) {
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  float f;
  memcpy(&f, &u, 4);
  return f;
#endif
}
GR_HD float fast_atan2f(float y, float x, const float* __restrict__ table) {
  const float y_abs = fabsf(y), x_abs = fabsf(x);
  const bool ylt = y_abs < x_abs;                                                 // :138-141
  const float z = GR_FDIV(ylt ? y_abs : x_abs, ylt ? x_abs : y_abs);
  // `z < TAN_MAP_RES` compares the float with a DOUBLE literal 0.003921569 (:32,147).  The literal lies
  // strictly between the floats 0x3b808081 and 0x3b808082, so for a float z the test is exactly
  // z < 0x3b808082 (no FP64 instruction on the device).
  // alpha = z*256 - .5 is evaluated in double and stored to float (:151); z*256 is exact and
  // the float subtraction rounds the same exact value once, so this is bit identical.
  float alpha = GR_FSUB(GR_FMUL(z, 256.0f), 0.5f);
  int index = (int)alpha;                       // z in [0, 1] -> alpha in [-0.5, 255.5] -> index in [0, 255]
  index = index < 0 ? 0 : (index > 255 ? 255 : index);  // only NaN / garbage inputs get here out of range: stay in the table
  alpha = GR_FSUB(alpha, (float)index);
  const float t0 = table[index], t1 = table[index + 1];
  const float lerp = GR_FADD(t0, GR_FMUL(GR_FSUB(t1, t0), alpha));                // :155-157
  const float base = (z < 0.00392156932502985f) ? z : lerp;
  // octant (:159-186):  x_abs > y_abs: x >= 0 ? (y >= 0 ? base : -base) : (y >= 0 ? pi - base : base - pi)
  //                     else        : y >= 0 ? (x >= 0 ? hp - base : hp + base) : (x >= 0 ? -hp + base : -hp - base)
  const bool xa = x_abs > y_abs, xp = x >= 0.0f, yp = y >= 0.0f;
  const float K = xa ? 3.14159265358979323846f : 1.57079632679489661923f;
  const unsigned sK = yp ? 0u : 0x80000000u;                      // sign of K: + for y >= 0
  const bool neg_base = xa ? yp : (yp == xp);
  const float sum = GR_FADD(gr_u2f(gr_f2u(K) ^ sK), gr_u2f(gr_f2u(base) ^ (neg_base ? 0x80000000u : 0u)));
  const float first = gr_u2f(gr_f2u(base) ^ sK);                  // x_abs > y_abs && x >= 0: +-base itself
  const float angle = (xa && xp) ? first : sum;
  return (y == 0.0f && x == 0.0f) ? 0.0f : angle;                 // :131-132
}

// gr_quadrature_demod_cf::work (gr_quadrature_demod_cf.cc:56-59):
// product = cur * conj(prev) expanded as gcc expands std::complex<float> operator*.
GR_HD float quad_demod(float2 cur, float2 prev, float gain, const float* __restrict__ table) {
  const float cr = prev.x, ci = -prev.y;
  const float re = GR_FSUB(GR_FMUL(cur.x, cr), GR_FMUL(cur.y, ci));
  const float im = GR_FADD(GR_FMUL(cur.x, ci), GR_FMUL(cur.y, cr));
  return GR_FMUL(gain, fast_atan2f(im, re, table));
}

// ---- float dot products in the reference's summation orders ---------------------------------
// `in` is addressed as in[i * stride]; rt = reversed taps (gr_fir_XXX.h.t:51,65).
// generic: gr_fir_XXX_generic.cc.t:28-55.
GR_HD float dot_generic(const float* __restrict__ rt, int ntaps, const float* __restrict__ in, long stride) {
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  int i = 0;
  const int n = ntaps & ~3;
  for (; i < n; i += 4) {
    a0 = GR_FADD(a0, GR_FMUL(rt[i], in[(long)i * stride]));
    a1 = GR_FADD(a1, GR_FMUL(rt[i + 1], in[(long)(i + 1) * stride]));
    a2 = GR_FADD(a2, GR_FMUL(rt[i + 2], in[(long)(i + 2) * stride]));
    a3 = GR_FADD(a3, GR_FMUL(rt[i + 3], in[(long)(i + 3) * stride]));
  }
  for (; i < ntaps; i++) a0 = GR_FADD(a0, GR_FMUL(rt[i], in[(long)i * stride]));
  return GR_FADD(GR_FADD(GR_FADD(a0, a1), a2), a3);
}

// SSE: gr_fir_fff_simd.cc:99-134 + float_dotprod_sse64.S:27-108.  `al` = (absolute index of
// in[0]) mod 4.  Lane l holds absolute positions == l (mod 4); the first nblocks%4 aligned
// blocks go to accumulator 0, the rest round-robin over accumulators 0..3; the four
// accumulators are combined as (a0+a1)+(a3+a2) only when at least one full group of four
// blocks ran, then lanes as (d0+d2)+(d1+d3).  Zero-tap positions contribute +-0 and are skipped.
GR_HD float dot_sse(const float* __restrict__ rt, int ntaps, const float* __restrict__ in, long stride, int al) {
  if (ntaps == 0) return 0.0f;
  const int nblocks = (ntaps + al - 1) / 4 + 1;
  const int nrem = nblocks & 3;
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int l = 0; l < 4; l++) acc[a][l] = 0.f;
  int b = 0;
  for (; b < nrem; b++) {
#pragma unroll
    for (int l = 0; l < 4; l++) {
      const int i = b * 4 + l - al;
      if (i >= 0 && i < ntaps) acc[0][l] = GR_FADD(acc[0][l], GR_FMUL(rt[i], in[(long)i * stride]));
    }
  }
  for (; b < nblocks; b += 4) {
#pragma unroll
    for (int a = 0; a < 4; a++) {
#pragma unroll
      for (int l = 0; l < 4; l++) {
        const int i = (b + a) * 4 + l - al;
        if (i >= 0 && i < ntaps) acc[a][l] = GR_FADD(acc[a][l], GR_FMUL(rt[i], in[(long)i * stride]));
      }
    }
  }
  float d[4];
  const bool grouped = (nblocks >> 2) != 0;
#pragma unroll
  for (int l = 0; l < 4; l++)
    d[l] = grouped ? GR_FADD(GR_FADD(acc[0][l], acc[1][l]), GR_FADD(acc[3][l], acc[2][l])) : acc[0][l];
  return GR_FADD(GR_FADD(d[0], d[2]), GR_FADD(d[1], d[3]));
}

GR_HD int mod4(long a) { return (int)(a & 3); }  // two's complement: correct for negative a

// ---- 8-tap MMSE interpolator (gri_mmse_fir_interpolator.cc:61-71) ---------------------------
// c[0..7] = coefficients applied to v[0..7] (= reference taps[imu][7-i]); v[i] = in[ii+i].
// generic (gr_fir_XXX_generic.cc.t:28-55): a_l = p[l] + p[l+4], result ((a0+a1)+a2)+a3.
// SSE (float_dotprod_sse64.S): lane (i+al)&3 sums p[i] then p[i+4] (always < 4 blocks, so one
// accumulator), result (d0+d2)+(d1+d3) over LANES.  With q_i = p[i]+p[i+4] the lane of q_i is
// (i+al)&3, so the result is (q0+q2)+(q1+q3) for al = 0, 2 and (q3+q1)+(q0+q2) for al = 1, 3:
// the same value for every alignment because IEEE addition commutes.  Hence no `al` here.
GR_HD float mmse8(const float c[8], const float v[8], int order) {
  float q[4];
#pragma unroll
  for (int i = 0; i < 4; i++) q[i] = GR_FADD(GR_FMUL(c[i], v[i]), GR_FMUL(c[i + 4], v[i + 4]));
  if (order == GR_ORDER_GENERIC) return GR_FADD(GR_FADD(GR_FADD(q[0], q[1]), q[2]), q[3]);
  return GR_FADD(GR_FADD(q[0], q[2]), GR_FADD(q[1], q[3]));
}

// ---- Mueller & Mueller loop state (digital_clock_recovery_mm_ff.cc:102-139) ------------------
struct MMParams {
  float gain_omega, gain_mu, omega_mid, omega_relative_limit;
};
struct MMState {
  float mu, omega, last_sample;
};
GR_HD float slice_pm1(float x) { return x < 0.f ? -1.0f : 1.0f; }  // :89-93
GR_HD float branchless_clip(float x, float clip) {                 // gr_math.h:63-69
  const float x1 = fabsf(GR_FADD(x, clip));
  const float x2 = fabsf(GR_FSUB(x, clip));
  return GR_FMUL(0.5f, GR_FSUB(x1, x2));  // 0.5*x1 in double then to float: exact either way
}
// One loop iteration after out = interp(&in[ii], mu): returns the input advance floor(mu).
GR_HD int mm_update(MMState& s, const MMParams& p, float out) {
  const float mm_val = GR_FSUB(GR_FMUL(slice_pm1(s.last_sample), out), GR_FMUL(slice_pm1(out), s.last_sample));
  s.last_sample = out;
  s.omega = GR_FADD(s.omega, GR_FMUL(p.gain_omega, mm_val));
  s.omega = GR_FADD(p.omega_mid, branchless_clip(GR_FSUB(s.omega, p.omega_mid), p.omega_relative_limit));
  s.mu = GR_FADD(GR_FADD(s.mu, s.omega), GR_FMUL(p.gain_mu, mm_val));
  const float fl = floorf(s.mu);
  s.mu = GR_FSUB(s.mu, fl);
  return (int)fl;
}
GR_HD int mm_imu(float mu) {  // (int) rint(mu * NSTEPS), round half to even (:64)
#ifdef __CUDA_ARCH__
  return __float2int_rn(GR_FMUL(mu, 128.0f));
#else
  return (int)rintf(mu * 128.0f);
#endif
}

// ---- slicers ----------------------------------------------------------------------------------
// pager_slicer_fb::slice (gr-pager/lib/pager_slicer_fb.cc:47-69)
GR_HD unsigned char slice4(float sample, float& avg, float alpha, float beta) {
  avg = GR_FADD(GR_FMUL(avg, beta), GR_FMUL(sample, alpha));
  sample = GR_FSUB(sample, avg);
  if (sample > 0.f) return sample > 2.0f ? 3 : 2;
  return sample < -2.0f ? 0 : 1;
}
GR_HD unsigned char slice2(float x) { return x >= 0.f ? 1 : 0; }  // gr_math.h:82-88

// ---- access-code correlator step (digital_correlate_access_code_bb.cc:97-130) -----------------
struct CorrParams {
  unsigned long long access_code, mask, flag_bit;
  unsigned threshold;
};
GR_HD unsigned popc64(unsigned long long x) {
#ifdef __CUDA_ARCH__
  return (unsigned)__popcll(x);
#else
  return (unsigned)__builtin_popcountll(x);
#endif
}
GR_HD unsigned char corr_step(unsigned long long& data_reg, unsigned long long& flag_reg, const CorrParams& p,
                              unsigned bit) {
  const unsigned char t = (unsigned char)(((data_reg >> 63) & 1ull) | (((flag_reg >> 63) & 1ull) << 1));
  const unsigned nwrong = popc64((data_reg ^ p.access_code) & p.mask);
  data_reg = (data_reg << 1) | (unsigned long long)(bit & 1u);
  flag_reg <<= 1;
  if (nwrong <= p.threshold) flag_reg |= p.flag_bit;
  return t;
}

}  // namespace grb
