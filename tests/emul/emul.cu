// TEST INFRASTRUCTURE: replays the per-thread device functions of csrc/gr_math.cuh and
// csrc/fft_radix.cuh on the host so that their logic can be checked against the oracle on a
// machine without a GPU.  Built by tests/conftest.py with nvcc (host code only,
// -Xcompiler -ffp-contract=off).  Never part of the product library.
#include <vector>
#include <cstring>
#include <cmath>
#include "../../gnuradio-3.5.0-dmr_b200/csrc/gr_math.cuh"
#include "../../gnuradio-3.5.0-dmr_b200/csrc/fft_radix.cuh"
#include "gr_tables.h"

using namespace grb;

template <int R, int DIR>
static void host_pass(int N, int Ns, const float2* src, float2* dst) {
  const int nb = N / R;
  for (int j = 0; j < nb; j++) {
    float2 v[R];
    for (int r = 0; r < R; r++) v[r] = src[j + r * nb];
    const int k = j % Ns;
    if (Ns > 1) {
      const double ph = DIR * 2.0 * M_PI * (double)k / ((double)Ns * R);
      apply_twiddle_powers<R>(v, make_float2((float)cos(ph), (float)sin(ph)));
    }
    butterfly<R, DIR>(v);
    const int o0 = (j - k) * R + k;
    for (int r = 0; r < R; r++) dst[o0 + r * Ns] = v[r];
  }
}

template <int DIR>
static int host_fft(int N, int npass, const int* radix, const float2* in, float2* out) {
  std::vector<float2> a(in, in + N), b(N);
  int Ns = 1;
  for (int p = 0; p < npass; p++) {
    switch (radix[p]) {
      case 2: host_pass<2, DIR>(N, Ns, a.data(), b.data()); break;
      case 3: host_pass<3, DIR>(N, Ns, a.data(), b.data()); break;
      case 4: host_pass<4, DIR>(N, Ns, a.data(), b.data()); break;
      case 5: host_pass<5, DIR>(N, Ns, a.data(), b.data()); break;
      case 8: host_pass<8, DIR>(N, Ns, a.data(), b.data()); break;
      case 10: host_pass<10, DIR>(N, Ns, a.data(), b.data()); break;
      case 16: host_pass<16, DIR>(N, Ns, a.data(), b.data()); break;
      case 20: host_pass<20, DIR>(N, Ns, a.data(), b.data()); break;
      default: return -1;
    }
    Ns *= radix[p];
    a.swap(b);
  }
  memcpy(out, a.data(), sizeof(float2) * N);
  return 0;
}

static float g_atan[257], g_mmse_eff[129 * 8];
static void tables() {
  static bool done = false;
  if (done) return;
  memcpy(g_atan, GR_FAST_ATAN_TABLE_BITS, sizeof g_atan);
  float raw[129 * 8];
  memcpy(raw, GR_MMSE_TAPS_BITS, sizeof raw);
  for (int s = 0; s < 129; s++)
    for (int i = 0; i < 8; i++) g_mmse_eff[s * 8 + i] = raw[s * 8 + (7 - i)];
  done = true;
}

extern "C" {
int emul_fft(int N, int npass, const int* radix, int dir, const float* in, float* out) {
  return dir < 0 ? host_fft<-1>(N, npass, radix, (const float2*)in, (float2*)out)
                 : host_fft<1>(N, npass, radix, (const float2*)in, (float2*)out);
}
void emul_fast_atan2f(const float* y, const float* x, float* out, long n) {
  tables();
  for (long i = 0; i < n; i++) out[i] = fast_atan2f(y[i], x[i], g_atan);
}
void emul_quad_demod(float gain, const float* in /* n+1 complex */, long n, float* out) {
  tables();
  const float2* c = (const float2*)in;
  for (long i = 0; i < n; i++) out[i] = quad_demod(c[i + 1], c[i], gain, g_atan);
}
// batched-layout dot product: in addressed with `stride`, rt = reversed taps
void emul_fir_fff(const float* taps, int ntaps, const float* in, long stride, long nout, float* out, int order,
                  long abs0) {
  std::vector<float> rt(ntaps > 0 ? ntaps : 1);
  for (int i = 0; i < ntaps; i++) rt[i] = taps[ntaps - 1 - i];
  for (long o = 0; o < nout; o++) {
    const float* p = in + o * stride;
    out[o * stride] = order == GR_ORDER_SSE ? dot_sse(rt.data(), ntaps, p, stride, mod4(abs0 + o))
                                             : dot_generic(rt.data(), ntaps, p, stride);
  }
}
int emul_mm(float omega, float gain_omega, float mu, float gain_mu, float lim, const float* in, int ninput,
            float* out, int noutput, int* consumed, int order, long abs0, float* state3) {
  tables();
  MMParams p;
  p.gain_omega = gain_omega; p.gain_mu = gain_mu; p.omega_relative_limit = lim;
  const float mn = (float)(omega * (1.0 - lim)), mx = (float)(omega * (1.0 + lim));
  p.omega_mid = (float)(0.5 * (mn + mx));
  MMState s;
  s.mu = mu; s.omega = omega; s.last_sample = 0.f;
  int ii = 0, oo = 0;
  const int ni = ninput - 8;
  while (oo < noutput && ii < ni) {
    float v[8];
    for (int i = 0; i < 8; i++) v[i] = in[ii + i];
    const float o = mmse8(g_mmse_eff + 8 * mm_imu(s.mu), v, order);
    out[oo++] = o;
    ii += mm_update(s, p, o);
  }
  *consumed = ii;
  state3[0] = s.mu; state3[1] = s.omega; state3[2] = s.last_sample;
  return oo;
}
void emul_slice4(float alpha, const float* in, long n, unsigned char* out) {
  float avg = 0.f;
  const float beta = (float)(1.0 - alpha);
  for (long i = 0; i < n; i++) out[i] = slice4(in[i], avg, alpha, beta);
}
void emul_corr(unsigned long long code, unsigned long long mask, unsigned long long flag_bit, unsigned thr,
               const unsigned char* in, long n, unsigned char* out) {
  CorrParams p;
  p.access_code = code; p.mask = mask; p.flag_bit = flag_bit; p.threshold = thr;
  unsigned long long d = 0, f = 0;
  for (long i = 0; i < n; i++) out[i] = corr_step(d, f, p, in[i]);
}
}
