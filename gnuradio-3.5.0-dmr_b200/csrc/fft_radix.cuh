// In-register DFT butterflies of the shared-memory FFT engine that replaces the FFTW-backed
// gri_fft_complex (gnuradio-core/src/lib/general/gri_fft.cc:97-146): unnormalised c2c,
// DIR = -1 forward / +1 backward.  Radices {2,3,4,5,8,10,16,20}: 160 = 16*10, 4096 = 16^3,
// 8000 = 20^3 (the channelizer sizes are not powers of two, so radix-5
// butterflies are required; 10 and 20 use the Good-Thomas prime-factor map, no inner twiddles).
#pragma once
#include <cuda_runtime.h>

#ifndef GR_HD
#define GR_HD __host__ __device__ __forceinline__
#endif

namespace grb {

// Complex add / subtract / scale map onto Blackwell's packed FP32 instructions (FADD2 / FMUL2 / FFMA2 on a
// register pair = one issue slot for the real and the imaginary part), which halves the instruction count
// of the add-heavy radix-4/5 butterflies.  Host build (tests/emul): plain arithmetic.
#ifdef __CUDA_ARCH__
__device__ __forceinline__ unsigned long long f2_pack(float2 a) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a.x), "f"(a.y));
  return r;
}
__device__ __forceinline__ float2 f2_unpack(unsigned long long v) {
  float2 r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
  return r;
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) {
  unsigned long long r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2_pack(a)), "l"(f2_pack(b)));
  return f2_unpack(r);
}
__device__ __forceinline__ float2 csub(float2 a, float2 b) {
  unsigned long long r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2_pack(a)), "l"(f2_pack(b)));
  return f2_unpack(r);
}
__device__ __forceinline__ float2 cscale(float2 a, float s) {
  unsigned long long r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2_pack(a)), "l"(f2_pack(make_float2(s, s))));
  return f2_unpack(r);
}
// a * s + b
__device__ __forceinline__ float2 cfma(float2 a, float s, float2 b) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(f2_pack(a)), "l"(f2_pack(make_float2(s, s))), "l"(f2_pack(b)));
  return f2_unpack(r);
}
#else
GR_HD float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
GR_HD float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
GR_HD float2 cscale(float2 a, float s) { return make_float2(a.x * s, a.y * s); }
GR_HD float2 cfma(float2 a, float s, float2 b) { return make_float2(a.x * s + b.x, a.y * s + b.y); }
#endif
GR_HD float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
// multiply by DIR*j  (DIR=-1: -j, DIR=+1: +j)
template <int DIR> GR_HD float2 mulj(float2 a) { return DIR < 0 ? make_float2(a.y, -a.x) : make_float2(-a.y, a.x); }
// multiply by the constant e^{DIR*j*theta} given c = cos(theta), s = sin(theta)
template <int DIR> GR_HD float2 cmulc(float2 a, float c, float s) {
  return DIR < 0 ? make_float2(a.x * c + a.y * s, a.y * c - a.x * s) : make_float2(a.x * c - a.y * s, a.y * c + a.x * s);
}

// All butterflies: v[k] <- sum_n v[n] * exp(DIR * 2*pi*j * n*k / R), natural order in and out.
template <int DIR> GR_HD void fft2(float2& a, float2& b) {
  const float2 t = a;
  a = cadd(t, b);
  b = csub(t, b);
}

template <int DIR> GR_HD void fft3(float2& a, float2& b, float2& c) {
  const float2 t1 = cadd(b, c);
  const float2 m = cfma(t1, -0.5f, a);
  const float2 jd = mulj<DIR>(cscale(csub(b, c), 0.86602540378443864676f));
  a = cadd(a, t1);
  b = cadd(m, jd);
  c = csub(m, jd);
}

template <int DIR> GR_HD void fft4(float2& v0, float2& v1, float2& v2, float2& v3) {
  const float2 s0 = cadd(v0, v2), s1 = csub(v0, v2);
  const float2 s2 = cadd(v1, v3), s3 = mulj<DIR>(csub(v1, v3));
  v0 = cadd(s0, s2);
  v2 = csub(s0, s2);
  v1 = cadd(s1, s3);
  v3 = csub(s1, s3);
}

template <int DIR> GR_HD void fft5(float2& x0, float2& x1, float2& x2, float2& x3, float2& x4) {
  const float2 t1 = cadd(x1, x4), t2 = cadd(x2, x3), t3 = csub(x1, x4), t4 = csub(x2, x3);
  const float2 t5 = cadd(t1, t2);
  const float2 m1 = cfma(t5, -0.25f, x0);
  const float2 m2 = cscale(csub(t1, t2), 0.55901699437494742410f);
  const float2 a = cadd(m1, m2), b = csub(m1, m2);
  const float s1 = 0.95105651629515357212f, s2 = 0.58778525229247312917f;
  const float2 c = cfma(t3, s1, cscale(t4, s2));
  const float2 d = cfma(t3, s2, cscale(t4, -s1));
  const float2 jc = mulj<DIR>(c), jd = mulj<DIR>(d);
  x0 = cadd(x0, t5);
  x1 = cadd(a, jc);
  x4 = csub(a, jc);
  x2 = cadd(b, jd);
  x3 = csub(b, jd);
}

template <int DIR> GR_HD void fft8(float2* v) {
  // 8 = 2 x 4 Cooley-Tukey: n = 4*n1 + n2, k = k1 + 2*k2
  const float r = 0.70710678118654752440f;
#pragma unroll
  for (int n2 = 0; n2 < 4; n2++) fft2<DIR>(v[n2], v[n2 + 4]);  // v[n2 + 4*k1] = A[n2][k1]
  v[5] = cmulc<DIR>(v[5], r, r);                               // W8^1
  v[6] = mulj<DIR>(v[6]);                                      // W8^2
  v[7] = cmulc<DIR>(v[7], -r, r);                              // W8^3
  fft4<DIR>(v[0], v[1], v[2], v[3]);                           // k1 = 0: v[k2] = X[2*k2]
  fft4<DIR>(v[4], v[5], v[6], v[7]);                           // k1 = 1: v[4+k2] = X[1+2*k2]
  const float2 e1 = v[1], e2 = v[2], e3 = v[3], o0 = v[4], o1 = v[5], o2 = v[6];
  v[1] = o0; v[2] = e1; v[3] = o1; v[4] = e2; v[5] = o2; v[6] = e3;
}

template <int DIR> GR_HD void fft16(float2* v) {
  // 16 = 4 x 4: n = 4*n1 + n2, k = k1 + 4*k2
  const float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f;  // cos, sin of pi/8
  const float r = 0.70710678118654752440f;
#pragma unroll
  for (int n2 = 0; n2 < 4; n2++) fft4<DIR>(v[n2], v[n2 + 4], v[n2 + 8], v[n2 + 12]);  // v[n2+4*k1] = A[n2][k1]
  // inner twiddles W16^(n2*k1) = e^{DIR j (n2*k1) pi/8}
  v[5] = cmulc<DIR>(v[5], c1, s1);      // m = 1
  v[6] = cmulc<DIR>(v[6], r, r);        // m = 2
  v[7] = cmulc<DIR>(v[7], s1, c1);      // m = 3
  v[9] = cmulc<DIR>(v[9], r, r);        // m = 2
  v[10] = mulj<DIR>(v[10]);             // m = 4
  v[11] = cmulc<DIR>(v[11], -r, r);     // m = 6
  v[13] = cmulc<DIR>(v[13], s1, c1);    // m = 3
  v[14] = cmulc<DIR>(v[14], -r, r);     // m = 6
  v[15] = cmulc<DIR>(v[15], -c1, -s1);  // m = 9
#pragma unroll
  for (int k1 = 0; k1 < 4; k1++) fft4<DIR>(v[4 * k1], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3]);  // v[4*k1+k2] = X[k1+4*k2]
  float2 t;
#define GRB_SWAP(a, b) t = v[a]; v[a] = v[b]; v[b] = t;
  GRB_SWAP(1, 4) GRB_SWAP(2, 8) GRB_SWAP(3, 12) GRB_SWAP(6, 9) GRB_SWAP(7, 13) GRB_SWAP(11, 14)
#undef GRB_SWAP
}

template <int DIR> GR_HD void fft10(float2* v) {
  // Good-Thomas 2 x 5: n = (5*n1 + 2*n2) mod 10, k = (5*k1 + 6*k2) mod 10
  float2 a[10];
#pragma unroll
  for (int n1 = 0; n1 < 2; n1++)
#pragma unroll
    for (int n2 = 0; n2 < 5; n2++) a[n1 * 5 + n2] = v[(5 * n1 + 2 * n2) % 10];
#pragma unroll
  for (int n1 = 0; n1 < 2; n1++) fft5<DIR>(a[n1 * 5], a[n1 * 5 + 1], a[n1 * 5 + 2], a[n1 * 5 + 3], a[n1 * 5 + 4]);
#pragma unroll
  for (int k2 = 0; k2 < 5; k2++) fft2<DIR>(a[k2], a[5 + k2]);
#pragma unroll
  for (int k1 = 0; k1 < 2; k1++)
#pragma unroll
    for (int k2 = 0; k2 < 5; k2++) v[(5 * k1 + 6 * k2) % 10] = a[k1 * 5 + k2];
}

template <int DIR> GR_HD void fft20(float2* v) {
  // Good-Thomas 4 x 5: n = (5*n1 + 4*n2) mod 20, k = (5*k1 + 16*k2) mod 20
  float2 a[20];
#pragma unroll
  for (int n1 = 0; n1 < 4; n1++)
#pragma unroll
    for (int n2 = 0; n2 < 5; n2++) a[n1 * 5 + n2] = v[(5 * n1 + 4 * n2) % 20];
#pragma unroll
  for (int n1 = 0; n1 < 4; n1++) fft5<DIR>(a[n1 * 5], a[n1 * 5 + 1], a[n1 * 5 + 2], a[n1 * 5 + 3], a[n1 * 5 + 4]);
#pragma unroll
  for (int k2 = 0; k2 < 5; k2++) fft4<DIR>(a[k2], a[5 + k2], a[10 + k2], a[15 + k2]);
#pragma unroll
  for (int k1 = 0; k1 < 4; k1++)
#pragma unroll
    for (int k2 = 0; k2 < 5; k2++) v[(5 * k1 + 16 * k2) % 20] = a[k1 * 5 + k2];
}

template <int R, int DIR> GR_HD void butterfly(float2* v) {
  if (R == 2) fft2<DIR>(v[0], v[1]);
  else if (R == 3) fft3<DIR>(v[0], v[1], v[2]);
  else if (R == 4) fft4<DIR>(v[0], v[1], v[2], v[3]);
  else if (R == 5) fft5<DIR>(v[0], v[1], v[2], v[3], v[4]);
  else if (R == 8) fft8<DIR>(v);
  else if (R == 10) fft10<DIR>(v);
  else if (R == 16) fft16<DIR>(v);
  else if (R == 20) fft20<DIR>(v);
}

// v[r] *= w^r, r = 1..R-1.  Powers are produced four at a time (w^q, q = 4g+1..4g+4, each the
// previous group's value times w^4) and consumed immediately, so only ~10 twiddle registers are
// live next to the R data values; dependency depth <= 2 + R/4 multiplications (<= ~7 ulp at R=20).
template <int R> GR_HD void apply_twiddle_powers(float2* v, float2 w1) {
  if (R < 2) return;
  float2 a = w1, b = cmul(w1, w1), c = cmul(b, w1), d = cmul(b, b);
  const float2 w4 = d;
  v[1] = cmul(v[1], a);
  if (R > 2) v[2] = cmul(v[2], b);
  if (R > 3) v[3] = cmul(v[3], c);
  if (R > 4) v[4] = cmul(v[4], d);
#pragma unroll
  for (int g = 5; g < R; g += 4) {
    a = cmul(a, w4);
    v[g] = cmul(v[g], a);
    if (g + 1 < R) { b = cmul(b, w4); v[g + 1] = cmul(v[g + 1], b); }
    if (g + 2 < R) { c = cmul(c, w4); v[g + 2] = cmul(v[g + 2], c); }
    if (g + 3 < R) { d = cmul(d, w4); v[g + 3] = cmul(v[g + 3], d); }
  }
}

}  // namespace grb
