// oracle shim (test infrastructure): same class interface as the reference's FFTW wrapper
// (gnuradio-core/src/lib/general/gri_fft.h:50-75, gri_fft.cc:97-146).  FFTW3f is a
// third-party dependency that is NOT in this image and NOT vendored by the reference, so
// the arithmetic behind execute() is the mathematical DFT evaluated in float64
// (mixed-radix Cooley-Tukey, exact O(N*p) butterflies per prime factor p) and rounded
// to float32 once.  qa_fft.py pins FFTW to rel 4e-4; this is ~1e-7 from any correct FFT.
#pragma once
#include <gr_complex.h>
#include <complex>
#include <vector>
#include <stdexcept>
#include <cmath>
class gri_fft_complex {
  int d_fft_size;
  bool d_forward;
  std::vector<gr_complex> d_in, d_out;
  std::vector<std::complex<double> > d_w, d_a, d_b;
  void rec(std::complex<double>* x, std::complex<double>* tmp, int n, int stride_w) {
    if (n == 1) return;
    int p = 2;
    while (n % p) p++;
    int m = n / p;
    // decimation in time: p sub-sequences of length m
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
 public:
  gri_fft_complex(int fft_size, bool forward = true)
      : d_fft_size(fft_size), d_forward(forward) {
    if (fft_size <= 0) throw std::out_of_range("gri_fftw: invalid fft_size");
    d_in.resize(fft_size); d_out.resize(fft_size);
    d_w.resize(fft_size); d_a.resize(fft_size); d_b.resize(fft_size);
    double s = forward ? -1.0 : 1.0;
    for (int i = 0; i < fft_size; i++) {
      double ph = s * 2.0 * M_PI * (double)i / (double)fft_size;
      d_w[i] = std::complex<double>(cos(ph), sin(ph));
    }
  }
  virtual ~gri_fft_complex() {}
  gr_complex* get_inbuf() { return d_in.data(); }
  gr_complex* get_outbuf() { return d_out.data(); }
  int inbuf_length() const { return d_fft_size; }
  int outbuf_length() const { return d_fft_size; }
  void execute() {
    for (int i = 0; i < d_fft_size; i++) d_a[i] = std::complex<double>(d_in[i]);
    rec(d_a.data(), d_b.data(), d_fft_size, 1);
    for (int i = 0; i < d_fft_size; i++) d_out[i] = gr_complex((float)d_a[i].real(), (float)d_a[i].imag());
  }
};
