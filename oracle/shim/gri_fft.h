// oracle shim (test infrastructure): same class interface as the reference's FFTW wrapper
// (gnuradio-core/src/lib/general/gri_fft.h:50-75, gri_fft.cc:97-146).  FFTW3f is a
// third-party dependency that is NOT in this image and NOT vendored by the reference, so
// two stand-ins live behind execute():
//   g_grref_fft_fast == 0 (default, used for every parity check): the mathematical DFT evaluated
//       in float64 (mixed-radix Cooley-Tukey, exact O(N*p) butterflies per prime factor p) and
//       rounded to float32 once.  qa_fft.py pins FFTW to rel 4e-4; this is ~1e-7 from any FFT.
//   g_grref_fft_fast == 1 (CPU-baseline TIMING only): a float32 Stockham radix-4/2/5/3 FFT with
//       precomputed twiddles, so that the timed CPU path is not penalised by the float64
//       reference DFT.  It is still slower than FFTW's SIMD codelets; bench.py says so.
#pragma once
#include <gr_complex.h>
#include <complex>
#include <vector>
#include <stdexcept>
#include <cmath>
extern int g_grref_fft_fast;
class gri_fft_complex {
  int d_fft_size;
  bool d_forward;
  std::vector<gr_complex> d_in, d_out;
  std::vector<std::complex<double> > d_w, d_a, d_b;
  // fast path
  std::vector<int> d_radix;
  std::vector<std::vector<gr_complex> > d_tw;  // per pass: [k][r] r = 1..R-1
  std::vector<gr_complex> d_t0, d_t1;
  bool d_fast_ok;
  void rec(std::complex<double>* x, std::complex<double>* tmp, int n, int stride_w) {
    if (n == 1) return;
    int p = 2;
    while (n % p) p++;
    int m = n / p;
    for (int r = 0; r < p; r++)
      for (int i = 0; i < m; i++) tmp[r * m + i] = x[i * p + r];
    for (int r = 0; r < p; r++) rec(tmp + r * m, x, m, stride_w * p);  // x reused as scratch
    for (int k = 0; k < m; k++)
      for (int q = 0; q < p; q++) {
        std::complex<double> acc = 0;
        int kk = k + q * m;
        for (int r = 0; r < p; r++)
          acc += tmp[r * m + k] * d_w[(size_t)((long long)r * kk % n) * stride_w];
        x[kk] = acc;
      }
  }
  void fast_setup() {
    int rem = d_fft_size;
    const int cand[] = {4, 2, 5, 3};
    for (int c : cand) while (rem % c == 0) { d_radix.push_back(c); rem /= c; }
    d_fast_ok = (rem == 1) && d_fft_size > 1;
    if (!d_fast_ok) return;
    const double s = d_forward ? -1.0 : 1.0;
    int Ns = 1;
    for (size_t p = 0; p < d_radix.size(); p++) {
      const int R = d_radix[p];
      std::vector<gr_complex> tw((size_t)Ns * (R - 1));
      for (int k = 0; k < Ns; k++)
        for (int r = 1; r < R; r++) {
          const double ph = s * 2.0 * M_PI * (double)k * r / ((double)Ns * R);
          tw[(size_t)k * (R - 1) + (r - 1)] = gr_complex((float)cos(ph), (float)sin(ph));
        }
      d_tw.push_back(tw);
      Ns *= R;
    }
    d_t0.resize(d_fft_size); d_t1.resize(d_fft_size);
  }
  static inline gr_complex mulj(gr_complex a, float sgn) { return gr_complex(-sgn * a.imag(), sgn * a.real()); }
  void fast_execute() {
    const int N = d_fft_size;
    const float sgn = d_forward ? -1.f : 1.f;
    const gr_complex* src = d_in.data();
    gr_complex* dst = d_t0.data();
    int Ns = 1;
    for (size_t p = 0; p < d_radix.size(); p++) {
      const int R = d_radix[p], nb = N / R;
      if (p + 1 == d_radix.size()) dst = d_out.data();
      const gr_complex* tw = d_tw[p].data();
      for (int j = 0; j < nb; j++) {
        const int k = j % Ns;
        gr_complex v[5];
        v[0] = src[j];
        for (int r = 1; r < R; r++) {
          const gr_complex x = src[j + r * nb], w = tw[(size_t)k * (R - 1) + (r - 1)];
          v[r] = gr_complex(x.real() * w.real() - x.imag() * w.imag(), x.real() * w.imag() + x.imag() * w.real());
        }
        gr_complex* o = dst + (j - k) * R + k;
        if (R == 4) {
          const gr_complex s0 = v[0] + v[2], s1 = v[0] - v[2], s2 = v[1] + v[3], s3 = mulj(v[1] - v[3], sgn);
          o[0] = s0 + s2; o[Ns] = s1 + s3; o[2 * Ns] = s0 - s2; o[3 * Ns] = s1 - s3;
        } else if (R == 2) {
          o[0] = v[0] + v[1]; o[Ns] = v[0] - v[1];
        } else if (R == 5) {
          const gr_complex t1 = v[1] + v[4], t2 = v[2] + v[3], t3 = v[1] - v[4], t4 = v[2] - v[3], t5 = t1 + t2;
          const gr_complex m1 = v[0] - 0.25f * t5, m2 = 0.559016994f * (t1 - t2);
          const gr_complex a = m1 + m2, b = m1 - m2;
          const gr_complex c = mulj(0.951056516f * t3 + 0.587785252f * t4, sgn);
          const gr_complex d = mulj(0.587785252f * t3 - 0.951056516f * t4, sgn);
          o[0] = v[0] + t5; o[Ns] = a + c; o[4 * Ns] = a - c; o[2 * Ns] = b + d; o[3 * Ns] = b - d;
        } else {  // 3
          const gr_complex t1 = v[1] + v[2], m = v[0] - 0.5f * t1, d = mulj(0.866025404f * (v[1] - v[2]), sgn);
          o[0] = v[0] + t1; o[Ns] = m + d; o[2 * Ns] = m - d;
        }
      }
      Ns *= R;
      src = dst;
      dst = (dst == d_t0.data()) ? d_t1.data() : d_t0.data();
    }
  }
 public:
  gri_fft_complex(int fft_size, bool forward = true)
      : d_fft_size(fft_size), d_forward(forward), d_fast_ok(false) {
    if (fft_size <= 0) throw std::out_of_range("gri_fftw: invalid fft_size");
    d_in.resize(fft_size); d_out.resize(fft_size);
    d_w.resize(fft_size); d_a.resize(fft_size); d_b.resize(fft_size);
    double s = forward ? -1.0 : 1.0;
    for (int i = 0; i < fft_size; i++) {
      double ph = s * 2.0 * M_PI * (double)i / (double)fft_size;
      d_w[i] = std::complex<double>(cos(ph), sin(ph));
    }
    fast_setup();
  }
  virtual ~gri_fft_complex() {}
  gr_complex* get_inbuf() { return d_in.data(); }
  gr_complex* get_outbuf() { return d_out.data(); }
  int inbuf_length() const { return d_fft_size; }
  int outbuf_length() const { return d_fft_size; }
  void execute() {
    if (g_grref_fft_fast && d_fast_ok) { fast_execute(); return; }
    for (int i = 0; i < d_fft_size; i++) d_a[i] = std::complex<double>(d_in[i]);
    rec(d_a.data(), d_b.data(), d_fft_size, 1);
    for (int i = 0; i < d_fft_size; i++) d_out[i] = gr_complex((float)d_a[i].real(), (float)d_a[i].imag());
  }
};
