/* GPU (B200) implementations of the GNU Radio 3.5 blocks on the channelize + DMR-demod hot path,
 * behind the reference's own block interface: same factory names and constructor arguments, same
 * setters, same work() / general_work() / forecast() / history() / output_multiple() contracts, same
 * exception types.  Each block is a thin host wrapper over one plan of libgr_cuda (gr_cuda.h); all
 * arithmetic runs in CUDA kernels -- there is no CPU fallback.
 *
 *   namespace gr_b200 {
 *     gr_make_fir_filter_ccf / gr_make_fir_filter_fff        filter/gr_fir_filter_XXX.h.t
 *     gr_make_freq_xlating_fir_filter_ccf                    filter/gr_freq_xlating_fir_filter_XXX.h.t
 *     gr_make_pfb_channelizer_ccf                            filter/gr_pfb_channelizer_ccf.h:31-34
 *     gr_make_pfb_arb_resampler_ccf                          filter/gr_pfb_arb_resampler_ccf.h:36-39
 *     gr_make_pfb_decimator_ccf                              filter/gr_pfb_decimator_ccf.h
 *     gr_make_fft_filter_ccc                                 filter/gr_fft_filter_ccc.h
 *     gr_make_fft_vcc                                        general/gr_fft_vcc.h:32-33
 *     gr_make_quadrature_demod_cf                            general/gr_quadrature_demod_cf.h
 *     digital_make_clock_recovery_mm_ff                      gr-digital/include/digital_clock_recovery_mm_ff.h:37-40
 *     pager_make_slicer_fb, digital_make_binary_slicer_fb    gr-pager/lib/pager_slicer_fb.h, gr-digital/include/digital_binary_slicer_fb.h
 *     digital_make_correlate_access_code_bb                  gr-digital/include/digital_correlate_access_code_bb.h:37-38
 *     digital_make_clock_recovery_mm_cc                      gr-digital/include/digital_clock_recovery_mm_cc.h:36-40
 *     gr_make_framer_sink_1                                  general/gr_framer_sink_1.h:33-34
 *     gr_make_map_bb, gr_make_unpack_k_bits_bb               general/gr_map_bb.h, gr_unpack_k_bits_bb.h
 *     gr_make_stream_to_streams, gr_make_vector_to_streams   general/gr_stream_to_streams.h, gr_vector_to_streams.h
 *   }
 * The names live in namespace gr_b200 so that they can be linked next to the CPU blocks; a flowgraph
 * switches over with `using namespace gr_b200;` or per block.
 *
 * Header only.  Link with -lgr_cuda.  Without GR_B200_USE_GNURADIO_RUNTIME the blocks derive from the
 * stand-in runtime of gr_b200_runtime.h (tests, non-GNU-Radio hosts); with it, from the installed
 * gr_block classes.
 */
#ifndef INCLUDED_GR_B200_BLOCKS_H
#define INCLUDED_GR_B200_BLOCKS_H

#ifdef GR_B200_USE_GNURADIO_RUNTIME
#include <gr_block.h>
#include <gr_io_signature.h>
#include <gr_sync_block.h>
#include <gr_sync_decimator.h>
#include <gr_sync_interpolator.h>
#include <gr_msg_queue.h>
#include <gr_message.h>
#include <gr_complex.h>
#define GR_B200_SPTR(T) boost::shared_ptr<T>
#define GR_B200_INITIAL_SPTR(p) gnuradio::get_initial_sptr(p)
#else
#include "gr_b200_runtime.h"
#define GR_B200_SPTR(T) std::shared_ptr<T>
#define GR_B200_INITIAL_SPTR(p) std::shared_ptr<typename std::remove_pointer<decltype(p)>::type>(p)
#endif

#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#include "gr_cuda.h"

namespace gr_b200 {

/* error code of the C ABI -> the exception the reference block throws in the same situation */
inline void throw_last_error(const char* where) {
  const std::string msg = std::string(where) + ": " + grcuda_last_error();
  switch (grcuda_last_error_code()) {
    case GRCUDA_EINVAL: throw std::invalid_argument(msg);
    case GRCUDA_ERANGE: throw std::out_of_range(msg);
    default: throw std::runtime_error(msg);
  }
}
inline int check_rc(int rc, const char* where) {
  if (rc < 0) throw_last_error(where);
  return rc;
}
inline const grcuda_complex* cin(const void* p) { return reinterpret_cast<const grcuda_complex*>(p); }
inline grcuda_complex* cout_(void* p) { return reinterpret_cast<grcuda_complex*>(p); }

/* ---- gr_fir_filter_ccf (filter/gr_fir_filter_XXX.cc.t:37-88) ------------------------------------ */
class gr_fir_filter_ccf;
typedef GR_B200_SPTR(gr_fir_filter_ccf) gr_fir_filter_ccf_sptr;
gr_fir_filter_ccf_sptr gr_make_fir_filter_ccf(int decimation, const std::vector<float>& taps);
class gr_fir_filter_ccf : public gr_sync_decimator {
  friend gr_fir_filter_ccf_sptr gr_make_fir_filter_ccf(int decimation, const std::vector<float>& taps);
  grcuda_fir_ccf* d_plan;
  gr_fir_filter_ccf(int decimation, const std::vector<float>& taps)
      : gr_sync_decimator("fir_filter_ccf", gr_make_io_signature(1, 1, sizeof(gr_complex)),
                          gr_make_io_signature(1, 1, sizeof(gr_complex)), decimation),
        d_plan(grcuda_fir_filter_ccf_create(decimation, taps.data(), (int)taps.size())) {
    if (!d_plan) throw_last_error("gr_fir_filter_ccf");
    set_history(grcuda_fir_filter_ccf_history(d_plan));     /* :51 */
  }
 public:
  ~gr_fir_filter_ccf() { grcuda_fir_filter_ccf_destroy(d_plan); }
  void set_taps(const std::vector<float>& taps) {            /* :59-64: takes effect at the next work() */
    check_rc(grcuda_fir_filter_ccf_set_taps(d_plan, taps.data(), (int)taps.size()), "set_taps");
  }
  int work(int noutput_items, gr_vector_const_void_star& input_items, gr_vector_void_star& output_items) {
    int r = check_rc(grcuda_fir_filter_ccf_work(d_plan, noutput_items, cin(input_items[0]), cout_(output_items[0])), "work");
    set_history(grcuda_fir_filter_ccf_history(d_plan));     /* :74-79: "history requirements may have changed" */
    return r;
  }
};
inline gr_fir_filter_ccf_sptr gr_make_fir_filter_ccf(int decimation, const std::vector<float>& taps) {
  return GR_B200_INITIAL_SPTR(new gr_fir_filter_ccf(decimation, taps));
}

/* ---- gr_fir_filter_fff ----------------------------------------------------------------------------- */
class gr_fir_filter_fff;
typedef GR_B200_SPTR(gr_fir_filter_fff) gr_fir_filter_fff_sptr;
gr_fir_filter_fff_sptr gr_make_fir_filter_fff(int decimation, const std::vector<float>& taps);
class gr_fir_filter_fff : public gr_sync_decimator {
  friend gr_fir_filter_fff_sptr gr_make_fir_filter_fff(int decimation, const std::vector<float>& taps);
  grcuda_fir_fff* d_plan;
  long d_nread;  /* items consumed so far: the SSE summation order depends on the absolute index */
  gr_fir_filter_fff(int decimation, const std::vector<float>& taps)
      : gr_sync_decimator("fir_filter_fff", gr_make_io_signature(1, 1, sizeof(float)), gr_make_io_signature(1, 1, sizeof(float)),
                          decimation),
        d_plan(grcuda_fir_filter_fff_create(decimation, taps.data(), (int)taps.size(), GRCUDA_ORDER_SSE)), d_nread(0) {
    if (!d_plan) throw_last_error("gr_fir_filter_fff");
    set_history(grcuda_fir_filter_fff_history(d_plan));
  }
 public:
  ~gr_fir_filter_fff() { grcuda_fir_filter_fff_destroy(d_plan); }
  void set_taps(const std::vector<float>& taps) {
    check_rc(grcuda_fir_filter_fff_set_taps(d_plan, taps.data(), (int)taps.size()), "set_taps");
  }
  int work(int noutput_items, gr_vector_const_void_star& input_items, gr_vector_void_star& output_items) {
    const long abs0 = d_nread - ((long)history() - 1);       /* absolute index of in[0] (first history item) */
    int r = check_rc(grcuda_fir_filter_fff_work(d_plan, noutput_items, (const float*)input_items[0], (float*)output_items[0], abs0),
                     "work");
    if (r > 0) d_nread += (long)r * decimation();
    set_history(grcuda_fir_filter_fff_history(d_plan));
    return r;
  }
};
inline gr_fir_filter_fff_sptr gr_make_fir_filter_fff(int decimation, const std::vector<float>& taps) {
  return GR_B200_INITIAL_SPTR(new gr_fir_filter_fff(decimation, taps));
}

/* ---- gr_freq_xlating_fir_filter_ccf (filter/gr_freq_xlating_fir_filter_XXX.cc.t:38-123) ---------- */
class gr_freq_xlating_fir_filter_ccf;
typedef GR_B200_SPTR(gr_freq_xlating_fir_filter_ccf) gr_freq_xlating_fir_filter_ccf_sptr;
gr_freq_xlating_fir_filter_ccf_sptr gr_make_freq_xlating_fir_filter_ccf(int decimation, const std::vector<float>& taps,
                                                                        double center_freq, double sampling_freq);
class gr_freq_xlating_fir_filter_ccf : public gr_sync_decimator {
  friend gr_freq_xlating_fir_filter_ccf_sptr gr_make_freq_xlating_fir_filter_ccf(int, const std::vector<float>&, double, double);
  grcuda_fxlat* d_plan;
  gr_freq_xlating_fir_filter_ccf(int decimation, const std::vector<float>& taps, double center_freq, double sampling_freq)
      : gr_sync_decimator("freq_xlating_fir_filter_ccf", gr_make_io_signature(1, 1, sizeof(gr_complex)),
                          gr_make_io_signature(1, 1, sizeof(gr_complex)), decimation),
        d_plan(grcuda_freq_xlating_fir_filter_ccf_create(decimation, taps.data(), (int)taps.size(), center_freq, sampling_freq)) {
    if (!d_plan) throw_last_error("gr_freq_xlating_fir_filter_ccf");
    set_history(grcuda_freq_xlating_fir_filter_ccf_history(d_plan));
  }
 public:
  ~gr_freq_xlating_fir_filter_ccf() { grcuda_freq_xlating_fir_filter_ccf_destroy(d_plan); }
  void set_center_freq(double center_freq) {                  /* :85-90 */
    check_rc(grcuda_freq_xlating_fir_filter_ccf_set_center_freq(d_plan, center_freq), "set_center_freq");
  }
  void set_taps(const std::vector<float>& taps) {             /* :92-97 */
    check_rc(grcuda_freq_xlating_fir_filter_ccf_set_taps(d_plan, taps.data(), (int)taps.size()), "set_taps");
  }
  int work(int noutput_items, gr_vector_const_void_star& input_items, gr_vector_void_star& output_items) {
    int r = check_rc(grcuda_freq_xlating_fir_filter_ccf_work(d_plan, noutput_items, cin(input_items[0]), cout_(output_items[0])),
                     "work");
    set_history(grcuda_freq_xlating_fir_filter_ccf_history(d_plan));
    return r;
  }
};
inline gr_freq_xlating_fir_filter_ccf_sptr gr_make_freq_xlating_fir_filter_ccf(int decimation, const std::vector<float>& taps,
                                                                               double center_freq, double sampling_freq) {
  return GR_B200_INITIAL_SPTR(new gr_freq_xlating_fir_filter_ccf(decimation, taps, center_freq, sampling_freq));
}

/* ---- gr_pfb_channelizer_ccf (filter/gr_pfb_channelizer_ccf.cc:35-200) ------------------------------ */
class gr_pfb_channelizer_ccf;
typedef GR_B200_SPTR(gr_pfb_channelizer_ccf) gr_pfb_channelizer_ccf_sptr;
gr_pfb_channelizer_ccf_sptr gr_make_pfb_channelizer_ccf(unsigned int numchans, const std::vector<float>& taps,
                                                        float oversample_rate = 1);
class gr_pfb_channelizer_ccf : public gr_block {
  friend gr_pfb_channelizer_ccf_sptr gr_make_pfb_channelizer_ccf(unsigned int, const std::vector<float>&, float);
  grcuda_pfb* d_plan;
  unsigned d_numchans;
  std::vector<float> d_taps;
  gr_pfb_channelizer_ccf(unsigned int numchans, const std::vector<float>& taps, float oversample_rate)
      : gr_block("pfb_channelizer_ccf", gr_make_io_signature(numchans, numchans, sizeof(gr_complex)),
                 gr_make_io_signature(1, 1, numchans * sizeof(gr_complex))),               /* :47-49 */
        d_plan(grcuda_pfb_channelizer_ccf_create(numchans, taps.data(), (int)taps.size(), oversample_rate)),
        d_numchans(numchans), d_taps(taps) {
    if (!d_plan) throw_last_error("gr_pfb_channelizer");     /* std::invalid_argument for a bad oversample rate (:57-60) */
    set_relative_rate(grcuda_pfb_channelizer_ccf_relative_rate(d_plan));                    /* :62 */
    set_history(grcuda_pfb_channelizer_ccf_history(d_plan));                                /* :136 */
    set_output_multiple(grcuda_pfb_channelizer_ccf_output_multiple(d_plan));                /* :89-92 */
  }
 public:
  ~gr_pfb_channelizer_ccf() { grcuda_pfb_channelizer_ccf_destroy(d_plan); }
  void set_taps(const std::vector<float>& taps) {             /* :104-139 */
    check_rc(grcuda_pfb_channelizer_ccf_set_taps(d_plan, taps.data(), (int)taps.size()), "set_taps");
    d_taps = taps;
    set_history(grcuda_pfb_channelizer_ccf_history(d_plan));
  }
  void print_taps() {                                         /* :141-152 */
    const unsigned T = (unsigned)grcuda_pfb_channelizer_ccf_taps_per_filter(d_plan);
    for (unsigned i = 0; i < d_numchans; i++) {
      printf("filter[%d]: [", i);
      for (unsigned j = 0; j < T; j++) {
        const size_t k = i + (size_t)j * d_numchans;
        printf(" %.4e", k < d_taps.size() ? d_taps[k] : 0.f);
      }
      printf("]\n\n");
    }
  }
  int general_work(int noutput_items, gr_vector_int&, gr_vector_const_void_star& input_items, gr_vector_void_star& output_items) {
    int consumed = 0;
    int r = check_rc(grcuda_pfb_channelizer_ccf_work(d_plan, noutput_items, reinterpret_cast<const grcuda_complex* const*>(input_items.data()),
                                                     cout_(output_items[0]), &consumed),
                     "general_work");
    consume_each(consumed);                                   /* :198 */
    return r;
  }
};
inline gr_pfb_channelizer_ccf_sptr gr_make_pfb_channelizer_ccf(unsigned int numchans, const std::vector<float>& taps,
                                                               float oversample_rate) {
  return GR_B200_INITIAL_SPTR(new gr_pfb_channelizer_ccf(numchans, taps, oversample_rate));
}

/* ---- gr_fft_filter_ccc (filter/gr_fft_filter_ccc.cc:46-106) ---------------------------------------------- */
class gr_fft_filter_ccc;
typedef GR_B200_SPTR(gr_fft_filter_ccc) gr_fft_filter_ccc_sptr;
gr_fft_filter_ccc_sptr gr_make_fft_filter_ccc(int decimation, const std::vector<gr_complex>& taps);
class gr_fft_filter_ccc : public gr_sync_decimator {
  friend gr_fft_filter_ccc_sptr gr_make_fft_filter_ccc(int, const std::vector<gr_complex>&);
  grcuda_fft_filter* d_plan;
  gr_fft_filter_ccc(int decimation, const std::vector<gr_complex>& taps)
      : gr_sync_decimator("fft_filter_ccc", gr_make_io_signature(1, 1, sizeof(gr_complex)),
                          gr_make_io_signature(1, 1, sizeof(gr_complex)), decimation),
        d_plan(grcuda_fft_filter_ccc_create(decimation, cin(taps.data()), (int)taps.size())) {
    if (!d_plan) throw_last_error("gr_fft_filter_ccc");
    set_history(1);                                                                          /* :58 */
    set_output_multiple(grcuda_fft_filter_ccc_output_multiple(d_plan));                      /* :66 */
  }
 public:
  ~gr_fft_filter_ccc() { grcuda_fft_filter_ccc_destroy(d_plan); }
  void set_taps(const std::vector<gr_complex>& taps) {          /* :75-79: takes effect at the next work() */
    check_rc(grcuda_fft_filter_ccc_set_taps(d_plan, cin(taps.data()), (int)taps.size()), "set_taps");
  }
  int work(int noutput_items, gr_vector_const_void_star& input_items, gr_vector_void_star& output_items) {
    int r = check_rc(grcuda_fft_filter_ccc_work(d_plan, noutput_items, cin(input_items[0]), cout_(output_items[0])), "work");
    set_output_multiple(grcuda_fft_filter_ccc_output_multiple(d_plan));                      /* :88 */
    return r;
  }
};
inline gr_fft_filter_ccc_sptr gr_make_fft_filter_ccc(int decimation, const std::vector<gr_complex>& taps) {
  return GR_B200_INITIAL_SPTR(new gr_fft_filter_ccc(decimation, taps));
}

/* ---- gr_pfb_decimator_ccf (filter/gr_pfb_decimator_ccf.cc:35-175) -------------------------------------- */
class gr_pfb_decimator_ccf;
typedef GR_B200_SPTR(gr_pfb_decimator_ccf) gr_pfb_decimator_ccf_sptr;
gr_pfb_decimator_ccf_sptr gr_make_pfb_decimator_ccf(unsigned int decim, const std::vector<float>& taps, unsigned int channel);
class gr_pfb_decimator_ccf : public gr_sync_block {
  friend gr_pfb_decimator_ccf_sptr gr_make_pfb_decimator_ccf(unsigned int, const std::vector<float>&, unsigned int);
  grcuda_pfb_decim* d_plan;
  gr_pfb_decimator_ccf(unsigned int decim, const std::vector<float>& taps, unsigned int channel)
      : gr_sync_block("pfb_decimator_ccf", gr_make_io_signature(decim, decim, sizeof(gr_complex)),
                      gr_make_io_signature(1, 1, sizeof(gr_complex))),                       /* :47-49 */
        d_plan(grcuda_pfb_decimator_ccf_create(decim, taps.data(), (int)taps.size(), channel)) {
    if (!d_plan) throw_last_error("gr_pfb_decimator_ccf");
    set_history(grcuda_pfb_decimator_ccf_history(d_plan));                                   /* :107 */
  }
 public:
  ~gr_pfb_decimator_ccf() { grcuda_pfb_decimator_ccf_destroy(d_plan); }
  void set_taps(const std::vector<float>& taps) {               /* :75-110 */
    check_rc(grcuda_pfb_decimator_ccf_set_taps(d_plan, taps.data(), (int)taps.size()), "set_taps");
    set_history(grcuda_pfb_decimator_ccf_history(d_plan));
  }
  int work(int noutput_items, gr_vector_const_void_star& input_items, gr_vector_void_star& output_items) {
    return check_rc(grcuda_pfb_decimator_ccf_work(d_plan, noutput_items,
                                                  reinterpret_cast<const grcuda_complex* const*>(input_items.data()),
                                                  cout_(output_items[0])),
                    "work");
  }
};
inline gr_pfb_decimator_ccf_sptr gr_make_pfb_decimator_ccf(unsigned int decim, const std::vector<float>& taps,
                                                           unsigned int channel) {
  return GR_B200_INITIAL_SPTR(new gr_pfb_decimator_ccf(decim, taps, channel));
}

/* ---- gr_pfb_arb_resampler_ccf (filter/gr_pfb_arb_resampler_ccf.cc:42-205) ------------------------------ */
class gr_pfb_arb_resampler_ccf;
typedef GR_B200_SPTR(gr_pfb_arb_resampler_ccf) gr_pfb_arb_resampler_ccf_sptr;
gr_pfb_arb_resampler_ccf_sptr gr_make_pfb_arb_resampler_ccf(float rate, const std::vector<float>& taps,
                                                            unsigned int filter_size = 32);
class gr_pfb_arb_resampler_ccf : public gr_block {
  friend gr_pfb_arb_resampler_ccf_sptr gr_make_pfb_arb_resampler_ccf(float, const std::vector<float>&, unsigned int);
  grcuda_pfb_arb* d_plan;
  gr_pfb_arb_resampler_ccf(float rate, const std::vector<float>& taps, unsigned int filter_size)
      : gr_block("pfb_arb_resampler_ccf", gr_make_io_signature(1, 1, sizeof(gr_complex)),
                 gr_make_io_signature(1, 1, sizeof(gr_complex))),                            /* :45-47 */
        d_plan(grcuda_pfb_arb_resampler_ccf_create(rate, taps.data(), (int)taps.size(), filter_size, 1)) {
    if (!d_plan) throw_last_error("gr_pfb_arb_resampler_ccf");
    set_relative_rate(grcuda_pfb_arb_resampler_ccf_relative_rate(d_plan));                   /* .h:162 */
    set_history(grcuda_pfb_arb_resampler_ccf_history(d_plan));                               /* :121 */
  }
 public:
  ~gr_pfb_arb_resampler_ccf() { grcuda_pfb_arb_resampler_ccf_destroy(d_plan); }
  void set_rate(float rate) {                                   /* .h:159-163 */
    check_rc(grcuda_pfb_arb_resampler_ccf_set_rate(d_plan, rate), "set_rate");
    set_relative_rate(rate);
  }
  void print_taps() {                                           /* :142-153 */
    const int T = grcuda_pfb_arb_resampler_ccf_taps_per_filter(d_plan), n = grcuda_pfb_arb_resampler_ccf_filter_size(d_plan);
    std::vector<float> t(T);
    for (int i = 0; i < n; i++) {
      grcuda_pfb_arb_resampler_ccf_get_taps(d_plan, i, 0, t.data(), T);
      printf("filter[%d]: [", i);
      for (int j = 0; j < T; j++) printf(" %.4e", t[j]);
      printf("]\n");
    }
  }
  int general_work(int noutput_items, gr_vector_int& ninput_items, gr_vector_const_void_star& input_items,
                   gr_vector_void_star& output_items) {
    int consumed = 0;
    int r = check_rc(grcuda_pfb_arb_resampler_ccf_work(d_plan, noutput_items, ninput_items[0], cin(input_items[0]),
                                                       cout_(output_items[0]), &consumed),
                     "general_work");
    consume_each(consumed);                                     /* :203 */
    return r;
  }
};
inline gr_pfb_arb_resampler_ccf_sptr gr_make_pfb_arb_resampler_ccf(float rate, const std::vector<float>& taps,
                                                                   unsigned int filter_size) {
  return GR_B200_INITIAL_SPTR(new gr_pfb_arb_resampler_ccf(rate, taps, filter_size));
}

/* ---- gr_fft_vcc (general/gr_fft_vcc.cc:34-64, gr_fft_vcc_fftw.cc:51-103) ---------------------------- */
class gr_fft_vcc;
typedef GR_B200_SPTR(gr_fft_vcc) gr_fft_vcc_sptr;
gr_fft_vcc_sptr gr_make_fft_vcc(int fft_size, bool forward, const std::vector<float>& window, bool shift = false);
class gr_fft_vcc : public gr_sync_block {
  friend gr_fft_vcc_sptr gr_make_fft_vcc(int, bool, const std::vector<float>&, bool);
  grcuda_fft* d_plan;
  gr_fft_vcc(int fft_size, bool forward, const std::vector<float>& window, bool shift)
      : gr_sync_block("fft_vcc", gr_make_io_signature(1, 1, (fft_size > 0 ? fft_size : 1) * sizeof(gr_complex)),
                      gr_make_io_signature(1, 1, (fft_size > 0 ? fft_size : 1) * sizeof(gr_complex))),
        d_plan(grcuda_fft_vcc_create(fft_size, forward, window.data(), (int)window.size(), shift)) {
    if (!d_plan) throw_last_error("gr_fft_vcc");             /* std::out_of_range for fft_size <= 0 (gri_fft.cc:104-105) */
  }
 public:
  ~gr_fft_vcc() { grcuda_fft_vcc_destroy(d_plan); }
  bool set_window(const std::vector<float>& window) {        /* gr_fft_vcc.cc:55-64 */
    return check_rc(grcuda_fft_vcc_set_window(d_plan, window.data(), (int)window.size()), "set_window") != 0;
  }
  int work(int noutput_items, gr_vector_const_void_star& input_items, gr_vector_void_star& output_items) {
    return check_rc(grcuda_fft_vcc_work(d_plan, noutput_items, cin(input_items[0]), cout_(output_items[0])), "work");
  }
};
inline gr_fft_vcc_sptr gr_make_fft_vcc(int fft_size, bool forward, const std::vector<float>& window, bool shift) {
  return GR_B200_INITIAL_SPTR(new gr_fft_vcc(fft_size, forward, window, shift));
}

/* ---- gr_quadrature_demod_cf (general/gr_quadrature_demod_cf.cc:31-62) -------------------------------- */
class gr_quadrature_demod_cf;
typedef GR_B200_SPTR(gr_quadrature_demod_cf) gr_quadrature_demod_cf_sptr;
gr_quadrature_demod_cf_sptr gr_make_quadrature_demod_cf(float gain);
class gr_quadrature_demod_cf : public gr_sync_block {
  friend gr_quadrature_demod_cf_sptr gr_make_quadrature_demod_cf(float gain);
  grcuda_quad* d_plan;
  gr_quadrature_demod_cf(float gain)
      : gr_sync_block("quadrature_demod_cf", gr_make_io_signature(1, 1, sizeof(gr_complex)), gr_make_io_signature(1, 1, sizeof(float))),
        d_plan(grcuda_quadrature_demod_cf_create(gain)) {
    if (!d_plan) throw_last_error("gr_quadrature_demod_cf");
    set_history(2);                                           /* :43: we need to look at the previous value */
  }
 public:
  ~gr_quadrature_demod_cf() { grcuda_quadrature_demod_cf_destroy(d_plan); }
  void set_gain(float gain) { check_rc(grcuda_quadrature_demod_cf_set_gain(d_plan, gain), "set_gain"); }
  float gain() const { return grcuda_quadrature_demod_cf_gain(d_plan); }
  int work(int noutput_items, gr_vector_const_void_star& input_items, gr_vector_void_star& output_items) {
    return check_rc(grcuda_quadrature_demod_cf_work(d_plan, noutput_items, cin(input_items[0]), (float*)output_items[0]), "work");
  }
};
inline gr_quadrature_demod_cf_sptr gr_make_quadrature_demod_cf(float gain) {
  return GR_B200_INITIAL_SPTR(new gr_quadrature_demod_cf(gain));
}

/* ---- digital_clock_recovery_mm_ff (gr-digital/lib/digital_clock_recovery_mm_ff.cc:36-139) ------------ */
class digital_clock_recovery_mm_ff;
typedef GR_B200_SPTR(digital_clock_recovery_mm_ff) digital_clock_recovery_mm_ff_sptr;
digital_clock_recovery_mm_ff_sptr digital_make_clock_recovery_mm_ff(float omega, float gain_omega, float mu, float gain_mu,
                                                                    float omega_relative_limit = 0.001);
class digital_clock_recovery_mm_ff : public gr_block {
  friend digital_clock_recovery_mm_ff_sptr digital_make_clock_recovery_mm_ff(float, float, float, float, float);
  grcuda_mm* d_plan;
  float d_gain_mu, d_gain_omega;
  long d_nread;
  digital_clock_recovery_mm_ff(float omega, float gain_omega, float mu, float gain_mu, float omega_relative_limit)
      : gr_block("clock_recovery_mm_ff", gr_make_io_signature(1, 1, sizeof(float)), gr_make_io_signature(1, 1, sizeof(float))),
        d_plan(grcuda_clock_recovery_mm_ff_create(1, omega, gain_omega, mu, gain_mu, omega_relative_limit, GRCUDA_ORDER_SSE)),
        d_gain_mu(gain_mu), d_gain_omega(gain_omega), d_nread(0) {
    if (!d_plan) throw_last_error("digital_clock_recovery_mm_ff");  /* std::out_of_range: omega < 1 or negative gains (:58-61) */
    set_relative_rate(1.0 / omega);                           /* :63 */
  }
  float state(int which) const {
    float mu = 0, omega = 0, last = 0;
    check_rc(grcuda_clock_recovery_mm_ff_get_state(d_plan, 0, &mu, &omega, &last), "get_state");
    return which == 0 ? mu : omega;
  }
 public:
  ~digital_clock_recovery_mm_ff() { grcuda_clock_recovery_mm_ff_destroy(d_plan); }
  void forecast(int noutput_items, gr_vector_int& ninput_items_required) {  /* :80-87 */
    const int n = grcuda_clock_recovery_mm_ff_forecast(d_plan, noutput_items);
    for (size_t i = 0; i < ninput_items_required.size(); i++) ninput_items_required[i] = n;
  }
  float mu() const { return state(0); }
  float omega() const { return state(1); }
  float gain_mu() const { return d_gain_mu; }
  float gain_omega() const { return d_gain_omega; }
  void set_gain_mu(float gain_mu) { d_gain_mu = gain_mu; check_rc(grcuda_clock_recovery_mm_ff_set_gain_mu(d_plan, gain_mu), "set_gain_mu"); }
  void set_gain_omega(float g) { d_gain_omega = g; check_rc(grcuda_clock_recovery_mm_ff_set_gain_omega(d_plan, g), "set_gain_omega"); }
  void set_mu(float mu) { check_rc(grcuda_clock_recovery_mm_ff_set_mu(d_plan, mu), "set_mu"); }
  void set_omega(float omega) { check_rc(grcuda_clock_recovery_mm_ff_set_omega(d_plan, omega), "set_omega"); }
  int general_work(int noutput_items, gr_vector_int& ninput_items, gr_vector_const_void_star& input_items,
                   gr_vector_void_star& output_items) {
    int consumed = 0;
    int r = check_rc(grcuda_clock_recovery_mm_ff_work(d_plan, noutput_items, ninput_items[0], (const float*)input_items[0],
                                                      (float*)output_items[0], &consumed, d_nread),
                     "general_work");
    d_nread += consumed;
    consume_each(consumed);                                   /* :137 */
    return r;
  }
};
inline digital_clock_recovery_mm_ff_sptr digital_make_clock_recovery_mm_ff(float omega, float gain_omega, float mu, float gain_mu,
                                                                           float omega_relative_limit) {
  return GR_B200_INITIAL_SPTR(new digital_clock_recovery_mm_ff(omega, gain_omega, mu, gain_mu, omega_relative_limit));
}

/* ---- pager_slicer_fb (gr-pager/lib/pager_slicer_fb.cc:30-84) / digital_binary_slicer_fb --------------- */
class pager_slicer_fb;
typedef GR_B200_SPTR(pager_slicer_fb) pager_slicer_fb_sptr;
pager_slicer_fb_sptr pager_make_slicer_fb(float alpha);
class pager_slicer_fb : public gr_sync_block {
  friend pager_slicer_fb_sptr pager_make_slicer_fb(float alpha);
  grcuda_slicer* d_plan;
  pager_slicer_fb(float alpha)
      : gr_sync_block("slicer_fb", gr_make_io_signature(1, 1, sizeof(float)), gr_make_io_signature(1, 1, sizeof(unsigned char))),
        d_plan(grcuda_pager_slicer_fb_create(alpha)) {
    if (!d_plan) throw_last_error("pager_slicer_fb");
  }
 public:
  ~pager_slicer_fb() { grcuda_slicer_destroy(d_plan); }
  float dc_offset() const { return grcuda_pager_slicer_fb_dc_offset(d_plan); }  /* pager_slicer_fb.h:54 */
  int work(int noutput_items, gr_vector_const_void_star& input_items, gr_vector_void_star& output_items) {
    return check_rc(grcuda_slicer_work(d_plan, noutput_items, (const float*)input_items[0], (unsigned char*)output_items[0]), "work");
  }
};
inline pager_slicer_fb_sptr pager_make_slicer_fb(float alpha) { return GR_B200_INITIAL_SPTR(new pager_slicer_fb(alpha)); }

class digital_binary_slicer_fb;
typedef GR_B200_SPTR(digital_binary_slicer_fb) digital_binary_slicer_fb_sptr;
digital_binary_slicer_fb_sptr digital_make_binary_slicer_fb();
class digital_binary_slicer_fb : public gr_sync_block {
  friend digital_binary_slicer_fb_sptr digital_make_binary_slicer_fb();
  grcuda_slicer* d_plan;
  digital_binary_slicer_fb()
      : gr_sync_block("binary_slicer_fb", gr_make_io_signature(1, 1, sizeof(float)), gr_make_io_signature(1, 1, sizeof(unsigned char))),
        d_plan(grcuda_binary_slicer_fb_create()) {
    if (!d_plan) throw_last_error("digital_binary_slicer_fb");
  }
 public:
  ~digital_binary_slicer_fb() { grcuda_slicer_destroy(d_plan); }
  int work(int noutput_items, gr_vector_const_void_star& input_items, gr_vector_void_star& output_items) {
    return check_rc(grcuda_slicer_work(d_plan, noutput_items, (const float*)input_items[0], (unsigned char*)output_items[0]), "work");
  }
};
inline digital_binary_slicer_fb_sptr digital_make_binary_slicer_fb() { return GR_B200_INITIAL_SPTR(new digital_binary_slicer_fb()); }

/* ---- digital_correlate_access_code_bb (gr-digital/lib/digital_correlate_access_code_bb.cc:36-133) ------ */
class digital_correlate_access_code_bb;
typedef GR_B200_SPTR(digital_correlate_access_code_bb) digital_correlate_access_code_bb_sptr;
digital_correlate_access_code_bb_sptr digital_make_correlate_access_code_bb(const std::string& access_code, int threshold);
class digital_correlate_access_code_bb : public gr_sync_block {
  friend digital_correlate_access_code_bb_sptr digital_make_correlate_access_code_bb(const std::string&, int);
  grcuda_corr* d_plan;
  digital_correlate_access_code_bb(const std::string& access_code, int threshold)
      : gr_sync_block("correlate_access_code_bb", gr_make_io_signature(1, 1, sizeof(char)), gr_make_io_signature(1, 1, sizeof(char))),
        d_plan(grcuda_correlate_access_code_bb_create(1, access_code.c_str(), threshold)) {
    if (!d_plan) throw_last_error("digital_correlate_access_code_bb");  /* std::out_of_range: access_code > 64 bits (:54-57) */
  }
 public:
  ~digital_correlate_access_code_bb() { grcuda_correlate_access_code_bb_destroy(d_plan); }
  bool set_access_code(const std::string& access_code) {      /* :64-85: false if longer than 64 */
    if (access_code.length() > 64) return false;
    return grcuda_correlate_access_code_bb_set_access_code(d_plan, access_code.c_str()) == GRCUDA_OK;
  }
  int work(int noutput_items, gr_vector_const_void_star& input_items, gr_vector_void_star& output_items) {
    return check_rc(grcuda_correlate_access_code_bb_work(d_plan, noutput_items, (const unsigned char*)input_items[0],
                                                         (unsigned char*)output_items[0]),
                    "work");
  }
};
inline digital_correlate_access_code_bb_sptr digital_make_correlate_access_code_bb(const std::string& access_code, int threshold) {
  return GR_B200_INITIAL_SPTR(new digital_correlate_access_code_bb(access_code, threshold));
}

/* ---- digital_clock_recovery_mm_cc (gr-digital/lib/digital_clock_recovery_mm_cc.cc:36-218) -------------- */
class digital_clock_recovery_mm_cc;
typedef GR_B200_SPTR(digital_clock_recovery_mm_cc) digital_clock_recovery_mm_cc_sptr;
digital_clock_recovery_mm_cc_sptr digital_make_clock_recovery_mm_cc(float omega, float gain_omega, float mu, float gain_mu,
                                                                    float omega_relative_limit = 0.001);
class digital_clock_recovery_mm_cc : public gr_block {
  friend digital_clock_recovery_mm_cc_sptr digital_make_clock_recovery_mm_cc(float, float, float, float, float);
  grcuda_mm_cc* d_plan;
  float d_gain_mu, d_gain_omega;
  digital_clock_recovery_mm_cc(float omega, float gain_omega, float mu, float gain_mu, float omega_relative_limit)
      : gr_block("clock_recovery_mm_cc", gr_make_io_signature(1, 1, sizeof(gr_complex)),
                 gr_make_io_signature(1, 2, sizeof(gr_complex))),   /* second output: float error signal (:55-56) */
        d_plan(grcuda_clock_recovery_mm_cc_create(1, omega, gain_omega, mu, gain_mu, omega_relative_limit)),
        d_gain_mu(gain_mu), d_gain_omega(gain_omega) {
    if (!d_plan) throw_last_error("digital_clock_recovery_mm_cc");   /* std::out_of_range (:65-68) */
    set_relative_rate(1.0 / omega);                                   /* :71 */
    set_history(3);                                                   /* :72 */
  }
  float state(int which) const {
    float mu = 0, omega = 0;
    check_rc(grcuda_clock_recovery_mm_cc_get_state(d_plan, 0, &mu, &omega), "get_state");
    return which == 0 ? mu : omega;
  }
 public:
  ~digital_clock_recovery_mm_cc() { grcuda_clock_recovery_mm_cc_destroy(d_plan); }
  void forecast(int noutput_items, gr_vector_int& ninput_items_required) {  /* :84-91 */
    const int n = grcuda_clock_recovery_mm_cc_forecast(d_plan, noutput_items);
    for (size_t i = 0; i < ninput_items_required.size(); i++) ninput_items_required[i] = n;
  }
  float mu() const { return state(0); }
  float omega() const { return state(1); }
  float gain_mu() const { return d_gain_mu; }
  float gain_omega() const { return d_gain_omega; }
  void set_verbose(bool) {}
  void set_gain_mu(float g) { d_gain_mu = g; check_rc(grcuda_clock_recovery_mm_cc_set_gain_mu(d_plan, g), "set_gain_mu"); }
  void set_gain_omega(float g) { d_gain_omega = g; check_rc(grcuda_clock_recovery_mm_cc_set_gain_omega(d_plan, g), "set_gain_omega"); }
  void set_mu(float mu) { check_rc(grcuda_clock_recovery_mm_cc_set_mu(d_plan, mu), "set_mu"); }
  void set_omega(float omega) { check_rc(grcuda_clock_recovery_mm_cc_set_omega(d_plan, omega), "set_omega"); }
  int general_work(int noutput_items, gr_vector_int& ninput_items, gr_vector_const_void_star& input_items,
                   gr_vector_void_star& output_items) {
    int consumed = 0;
    float* err = output_items.size() >= 2 ? (float*)output_items[1] : 0;   /* :131 */
    int r = check_rc(grcuda_clock_recovery_mm_cc_work(d_plan, noutput_items, ninput_items[0], cin(input_items[0]),
                                                      cout_(output_items[0]), err, &consumed),
                     "general_work");
    if (consumed > 0) consume_each(consumed);                        /* :207-214 */
    return r;
  }
};
inline digital_clock_recovery_mm_cc_sptr digital_make_clock_recovery_mm_cc(float omega, float gain_omega, float mu, float gain_mu,
                                                                           float omega_relative_limit) {
  return GR_B200_INITIAL_SPTR(new digital_clock_recovery_mm_cc(omega, gain_omega, mu, gain_mu, omega_relative_limit));
}

/* ---- gr_framer_sink_1 (general/gr_framer_sink_1.cc:75-196) ------------------------------------------------ */
class gr_framer_sink_1;
typedef GR_B200_SPTR(gr_framer_sink_1) gr_framer_sink_1_sptr;
gr_framer_sink_1_sptr gr_make_framer_sink_1(gr_msg_queue_sptr target_queue);
class gr_framer_sink_1 : public gr_sync_block {
  friend gr_framer_sink_1_sptr gr_make_framer_sink_1(gr_msg_queue_sptr target_queue);
  grcuda_framer* d_plan;
  gr_msg_queue_sptr d_target_queue;
  std::vector<grcuda_framer_msg> d_msgs;
  std::vector<unsigned char> d_payload;
  gr_framer_sink_1(gr_msg_queue_sptr target_queue)
      : gr_sync_block("framer_sink_1", gr_make_io_signature(1, 1, sizeof(unsigned char)), gr_make_io_signature(0, 0, 0)),
        d_plan(grcuda_framer_sink_1_create(1, 4096, 1 << 22)), d_target_queue(target_queue), d_msgs(4096), d_payload(1 << 22) {
    if (!d_plan) throw_last_error("gr_framer_sink_1");
  }
 public:
  ~gr_framer_sink_1() { grcuda_framer_sink_1_destroy(d_plan); }
  int work(int noutput_items, gr_vector_const_void_star& input_items, gr_vector_void_star&) {
    /* at most noutput_items / 32 packets can complete in one call: bound the chunk by the device queue */
    int done = 0;
    while (done < noutput_items) {
      const int n = noutput_items - done < 4096 * 32 ? noutput_items - done : 4096 * 32;
      check_rc(grcuda_framer_sink_1_work(d_plan, n, (const unsigned char*)input_items[0] + done), "work");
      int dropped = 0;
      const int k = check_rc(grcuda_framer_sink_1_read(d_plan, d_msgs.data(), (int)d_msgs.size(), d_payload.data(), d_payload.size(),
                                                       &dropped), "read");
      for (int i = 0; i < k; i++) {   /* gr_make_message(0, whitener_offset, 0, len) + insert_tail (:139-146, :170-178) */
        gr_message_sptr msg = gr_make_message(0, d_msgs[i].whitener_offset, 0, d_msgs[i].length);
        if (d_msgs[i].length > 0 && d_msgs[i].payload_offset >= 0)
          memcpy(msg->msg(), d_payload.data() + d_msgs[i].payload_offset, d_msgs[i].length);
        d_target_queue->insert_tail(msg);
      }
      if (dropped) fprintf(stderr, "gr_b200::gr_framer_sink_1: %d message(s) did not fit the device queue\n", dropped);
      done += n;
    }
    return noutput_items;
  }
};
inline gr_framer_sink_1_sptr gr_make_framer_sink_1(gr_msg_queue_sptr target_queue) {
  return GR_B200_INITIAL_SPTR(new gr_framer_sink_1(target_queue));
}

/* ---- gr_map_bb (general/gr_map_bb.cc:35-61) / gr_unpack_k_bits_bb (general/gr_unpack_k_bits_bb.cc:38-70) ------ */
class gr_map_bb;
typedef GR_B200_SPTR(gr_map_bb) gr_map_bb_sptr;
gr_map_bb_sptr gr_make_map_bb(const std::vector<int>& map);
class gr_map_bb : public gr_sync_block {
  friend gr_map_bb_sptr gr_make_map_bb(const std::vector<int>& map);
  grcuda_map_bb* d_plan;
  gr_map_bb(const std::vector<int>& map)
      : gr_sync_block("map_bb", gr_make_io_signature(1, 1, sizeof(unsigned char)), gr_make_io_signature(1, 1, sizeof(unsigned char))),
        d_plan(grcuda_map_bb_create(map.data(), (int)map.size())) {
    if (!d_plan) throw_last_error("gr_map_bb");
  }
 public:
  ~gr_map_bb() { grcuda_map_bb_destroy(d_plan); }
  int work(int noutput_items, gr_vector_const_void_star& input_items, gr_vector_void_star& output_items) {
    return check_rc(grcuda_map_bb_work(d_plan, noutput_items, (const unsigned char*)input_items[0], (unsigned char*)output_items[0]), "work");
  }
};
inline gr_map_bb_sptr gr_make_map_bb(const std::vector<int>& map) { return GR_B200_INITIAL_SPTR(new gr_map_bb(map)); }

class gr_unpack_k_bits_bb;
typedef GR_B200_SPTR(gr_unpack_k_bits_bb) gr_unpack_k_bits_bb_sptr;
gr_unpack_k_bits_bb_sptr gr_make_unpack_k_bits_bb(unsigned k);
class gr_unpack_k_bits_bb : public gr_sync_interpolator {
  friend gr_unpack_k_bits_bb_sptr gr_make_unpack_k_bits_bb(unsigned k);
  grcuda_unpack_k_bits* d_plan;
  gr_unpack_k_bits_bb(unsigned k)
      : gr_sync_interpolator("unpack_k_bits_bb", gr_make_io_signature(1, 1, sizeof(unsigned char)),
                             gr_make_io_signature(1, 1, sizeof(unsigned char)), k ? k : 1),
        d_plan(grcuda_unpack_k_bits_bb_create(k)) {
    if (!d_plan) throw_last_error("gr_unpack_k_bits_bb");   /* std::out_of_range("interpolation must be > 0") (:44-45) */
  }
 public:
  ~gr_unpack_k_bits_bb() { grcuda_unpack_k_bits_bb_destroy(d_plan); }
  int work(int noutput_items, gr_vector_const_void_star& input_items, gr_vector_void_star& output_items) {
    check_rc(grcuda_unpack_k_bits_bb_work(d_plan, noutput_items, (const unsigned char*)input_items[0], (unsigned char*)output_items[0]), "work");
    return noutput_items;
  }
};
inline gr_unpack_k_bits_bb_sptr gr_make_unpack_k_bits_bb(unsigned k) { return GR_B200_INITIAL_SPTR(new gr_unpack_k_bits_bb(k)); }

/* ---- gr_stream_to_streams / gr_vector_to_streams (general/gr_stream_to_streams.cc:37-66, gr_vector_to_streams.cc:37-70) */
class gr_stream_to_streams;
typedef GR_B200_SPTR(gr_stream_to_streams) gr_stream_to_streams_sptr;
gr_stream_to_streams_sptr gr_make_stream_to_streams(size_t item_size, size_t nstreams);
class gr_stream_to_streams : public gr_sync_decimator {
  friend gr_stream_to_streams_sptr gr_make_stream_to_streams(size_t item_size, size_t nstreams);
  grcuda_streams* d_plan;
  gr_stream_to_streams(size_t item_size, size_t nstreams)
      : gr_sync_decimator("stream_to_streams", gr_make_io_signature(1, 1, (int)item_size),
                          gr_make_io_signature((int)nstreams, (int)nstreams, (int)item_size), (unsigned)nstreams),
        d_plan(grcuda_stream_to_streams_create(item_size, nstreams)) {
    if (!d_plan) throw_last_error("gr_stream_to_streams");
  }
 public:
  ~gr_stream_to_streams() { grcuda_streams_destroy(d_plan); }
  int work(int noutput_items, gr_vector_const_void_star& input_items, gr_vector_void_star& output_items) {
    return check_rc(grcuda_streams_work(d_plan, noutput_items, input_items[0], output_items.data()), "work");
  }
};
inline gr_stream_to_streams_sptr gr_make_stream_to_streams(size_t item_size, size_t nstreams) {
  return GR_B200_INITIAL_SPTR(new gr_stream_to_streams(item_size, nstreams));
}

class gr_vector_to_streams;
typedef GR_B200_SPTR(gr_vector_to_streams) gr_vector_to_streams_sptr;
gr_vector_to_streams_sptr gr_make_vector_to_streams(size_t item_size, size_t nstreams);
class gr_vector_to_streams : public gr_sync_block {
  friend gr_vector_to_streams_sptr gr_make_vector_to_streams(size_t item_size, size_t nstreams);
  grcuda_streams* d_plan;
  gr_vector_to_streams(size_t item_size, size_t nstreams)
      : gr_sync_block("vector_to_streams", gr_make_io_signature(1, 1, (int)(nstreams * item_size)),
                      gr_make_io_signature((int)nstreams, (int)nstreams, (int)item_size)),
        d_plan(grcuda_vector_to_streams_create(item_size, nstreams)) {
    if (!d_plan) throw_last_error("gr_vector_to_streams");
  }
 public:
  ~gr_vector_to_streams() { grcuda_streams_destroy(d_plan); }
  int work(int noutput_items, gr_vector_const_void_star& input_items, gr_vector_void_star& output_items) {
    return check_rc(grcuda_streams_work(d_plan, noutput_items, input_items[0], output_items.data()), "work");
  }
};
inline gr_vector_to_streams_sptr gr_make_vector_to_streams(size_t item_size, size_t nstreams) {
  return GR_B200_INITIAL_SPTR(new gr_vector_to_streams(item_size, nstreams));
}

}  // namespace gr_b200

#endif /* INCLUDED_GR_B200_BLOCKS_H */
