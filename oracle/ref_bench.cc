// ORACLE / TEST INFRASTRUCTURE ONLY (see ref_capi.cc header).
// Multi-threaded CPU timing harness over the reference's OWN block classes for the flagship
// path (bench.py --impl reference and the cpu_baseline leg):
//   gr_stream_to_streams -> gr_pfb_channelizer_ccf -> gr_vector_to_streams -> per channel:
//   gr_quadrature_demod_cf -> gr_fir_filter_fff -> digital_clock_recovery_mm_ff ->
//   pager_slicer_fb -> gr_map_bb -> gr_unpack_k_bits_bb(2) -> digital_correlate_access_code_bb
// The reference's thread-per-block scheduler cannot be built here (Boost/SWIG/Py2 absent); the
// blocks are driven directly in large chunks, which only favours the CPU side (no 32 KB buffer
// hand-offs).  Phase 1 shards the channelizer by TIME across threads (tap-history halo re-read),
// phase 2 shards the demod tail by CHANNEL -- BASELINE.md section 4.
#include <gr_block.h>
#include <gr_fir_util.h>
#include <gr_pfb_channelizer_ccf.h>
#include <gr_quadrature_demod_cf.h>
#include <gr_fir_filter_fff.h>
#include <gr_map_bb.h>
#include <gr_unpack_k_bits_bb.h>
#include <digital_clock_recovery_mm_ff.h>
#include <digital_correlate_access_code_bb.h>
#include <pager_slicer_fb.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

int g_grref_fft_fast = 0;

namespace {
double now() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
template <class B> int work1(B& blk, int nout, int nin, const void* in, void* out) {
  gr_vector_int ni(1, nin);
  gr_vector_const_void_star iv(1, in);
  gr_vector_void_star ov(1, out);
  return blk->general_work(nout, ni, iv, ov);
}
}  // namespace

extern "C" {

void grref_set_fft_fast(int on) { g_grref_fft_fast = on; }

struct grref_chain_params {
  unsigned M;
  const float* pfb_taps; int pfb_ntaps;
  float quad_gain;
  const float* rrc_taps; int rrc_ntaps;
  float omega, gain_omega, mu, gain_mu, limit, slicer_alpha;
  const int* symbol_map; int symbol_map_len;
  const char* access_code; int threshold;
};

// x: rows*M interleaved complex samples (stream start: zero history).  Returns 0 on success.
// seconds[0] = channelizer phase, seconds[1] = demod phase; *nhits = sync flags found.
int grref_bench_chain(const grref_chain_params* p, const float* x_, long rows, int nthreads, double* seconds,
                      long* nhits, float* y_keep /* optional rows*M complex out */) {
  const unsigned M = p->M;
  const gr_complex* x = (const gr_complex*)x_;
  std::vector<gr_complex> Y((size_t)rows * M);
  std::vector<float> pt(p->pfb_taps, p->pfb_taps + p->pfb_ntaps);
  const int T = (int)ceil((double)p->pfb_ntaps / (double)M);
  if (nthreads < 1) nthreads = 1;
  std::atomic<long> hits(0);
  // Block construction (flowgraph set-up: 129 MMSE filters per M&M block, M branch filters per
  // channelizer, ...) is NOT timed; only the steady-state work() calls are.
  std::vector<gr_pfb_channelizer_ccf_sptr> pfbs(nthreads);
  {
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; t++) th.emplace_back([&, t]() { pfbs[t] = gr_make_pfb_channelizer_ccf(M, pt, 1.0f); });
    for (auto& t : th) t.join();
  }
  // ---- phase 1: channelizer, time sharded --------------------------------------------------
  double t0 = now();
  {
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; t++) {
      th.emplace_back([&, t]() {
        const long r0 = rows * t / nthreads, r1 = rows * (t + 1) / nthreads;
        const long n = r1 - r0;
        if (n <= 0) return;
        gr_pfb_channelizer_ccf_sptr pfb = pfbs[t];
        // gr_stream_to_streams (general/gr_stream_to_streams.cc:57-63): item j of every M goes to stream j
        std::vector<std::vector<gr_complex> > st(M, std::vector<gr_complex>((size_t)n + T + 4));
        // (done in 32x32 tiles: the per-item memcpy scatter of the reference block thrashes the
        // cache at M = 8000; tiling only helps the CPU side)
        for (long rb = -T; rb < n; rb += 32)
          for (unsigned jb = 0; jb < M; jb += 32)
            for (long r = rb; r < std::min<long>(rb + 32, n); r++) {
              const long gr = r0 + r;
              const unsigned je = std::min<unsigned>(jb + 32, M);
              if (gr >= 0) for (unsigned j = jb; j < je; j++) st[j][(size_t)(r + T)] = x[(size_t)gr * M + j];
              else for (unsigned j = jb; j < je; j++) st[j][(size_t)(r + T)] = gr_complex(0, 0);
            }
        gr_vector_int ni(M, (int)(n + T));
        gr_vector_const_void_star iv(M);
        for (unsigned j = 0; j < M; j++) iv[j] = st[j].data();
        gr_vector_void_star ov(1, &Y[(size_t)r0 * M]);
        pfb->general_work((int)n, ni, iv, ov);  // first call after set_taps returns 0
        pfb->general_work((int)n, ni, iv, ov);
      });
    }
    for (auto& t : th) t.join();
  }
  double t1 = now();
  // per-channel block instances, constructed untimed
  std::vector<int> map(p->symbol_map, p->symbol_map + p->symbol_map_len);
  std::vector<float> rt(p->rrc_taps, p->rrc_taps + p->rrc_ntaps);
  struct Tail {
    gr_quadrature_demod_cf_sptr q; gr_fir_filter_fff_sptr fir; digital_clock_recovery_mm_ff_sptr mm;
    pager_slicer_fb_sptr s4; gr_map_bb_sptr mp; gr_unpack_k_bits_bb_sptr up; digital_correlate_access_code_bb_sptr cor;
  };
  std::vector<Tail> tails(M);
  {
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; t++)
      th.emplace_back([&, t]() {
        for (unsigned c = t; c < M; c += nthreads) {
          Tail& k = tails[c];
          k.q = gr_make_quadrature_demod_cf(p->quad_gain);
          k.fir = gr_make_fir_filter_fff(1, rt);
          k.mm = digital_make_clock_recovery_mm_ff(p->omega, p->gain_omega, p->mu, p->gain_mu, p->limit);
          k.s4 = pager_make_slicer_fb(p->slicer_alpha);
          k.mp = gr_make_map_bb(map);
          k.up = gr_make_unpack_k_bits_bb(2);
          k.cor = digital_make_correlate_access_code_bb(std::string(p->access_code), p->threshold);
        }
      });
    for (auto& t : th) t.join();
  }
  double t1b = now();
  // ---- phase 2: demod tail, channel sharded ---------------------------------------------------
  {
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; t++) {
      th.emplace_back([&, t]() {
        const int h = p->rrc_ntaps - 1;
        std::vector<gr_complex> col((size_t)rows + 1 + 4);
        std::vector<float> d((size_t)rows + h + 8), f((size_t)rows + 8), soft((size_t)rows + 8);
        std::vector<unsigned char> sl((size_t)rows + 8), db((size_t)rows + 8), bits((size_t)2 * rows + 16), cb((size_t)2 * rows + 16);
        long local_hits = 0;
        const unsigned c0 = (unsigned)((unsigned long long)M * t / nthreads), c1 = (unsigned)((unsigned long long)M * (t + 1) / nthreads);
        const unsigned CB = 16;  // channels transposed at a time (gr_vector_to_streams, tiled)
        std::vector<gr_complex> cols((size_t)CB * (rows + 1));
        for (unsigned cb0 = c0; cb0 < c1; cb0 += CB) {
         const unsigned cbn = std::min(CB, c1 - cb0);
         for (long r = 0; r < rows; r++)
           for (unsigned k = 0; k < cbn; k++) cols[(size_t)k * (rows + 1) + r + 1] = Y[(size_t)r * M + cb0 + k];
         for (unsigned k = 0; k < cbn; k++) {
          // gr_vector_to_streams (general/gr_vector_to_streams.cc:55-61)
          gr_complex* colp = &cols[(size_t)k * (rows + 1)];
          colp[0] = gr_complex(0, 0);
          std::memcpy(col.data(), colp, sizeof(gr_complex) * (rows + 1));
          Tail& tk = tails[cb0 + k];
          gr_quadrature_demod_cf_sptr q = tk.q;
          std::fill(d.begin(), d.begin() + h, 0.f);
          work1(q, (int)rows, (int)rows + 1, col.data(), d.data() + h);
          gr_fir_filter_fff_sptr fir = tk.fir;
          work1(fir, (int)rows, (int)rows + h, d.data(), f.data());
          digital_clock_recovery_mm_ff_sptr mm = tk.mm;
          int nsym = work1(mm, (int)rows, (int)rows, f.data(), soft.data());
          pager_slicer_fb_sptr s4 = tk.s4;
          work1(s4, nsym, nsym, soft.data(), sl.data());
          gr_map_bb_sptr mp = tk.mp;
          work1(mp, nsym, nsym, sl.data(), db.data());
          gr_unpack_k_bits_bb_sptr up = tk.up;
          work1(up, 2 * nsym, nsym, db.data(), bits.data());
          digital_correlate_access_code_bb_sptr cor = tk.cor;
          work1(cor, 2 * nsym, 2 * nsym, bits.data(), cb.data());
          for (int i = 0; i < 2 * nsym; i++) local_hits += (cb[i] >> 1) & 1;
         }
        }
        hits += local_hits;
      });
    }
    for (auto& t : th) t.join();
  }
  double t2 = now();
  seconds[0] = t1 - t0;
  seconds[1] = t2 - t1b;
  *nhits = hits.load();
  if (y_keep) memcpy(y_keep, Y.data(), Y.size() * sizeof(gr_complex));
  return 0;
}

}  // extern "C"
