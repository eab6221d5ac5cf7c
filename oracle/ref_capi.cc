// ORACLE / TEST INFRASTRUCTURE ONLY -- never linked into, imported by, or executed from the
// product path (libgr_cuda / grb200).  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load the library this file builds.
//
// C ABI over the reference's OWN classes, compiled unmodified from /root/reference by
// oracle/build_ref.sh into oracle/_ref/libgrref.so (git-ignored).  The classes are driven
// the way gr_block_executor::run_one_iteration drives them
// (gnuradio-core/src/lib/runtime/gr_block_executor.cc:370-371): general_work() on
// history-prefixed buffers; the harness (tests/refharness.py) lays those buffers out.
#include <gr_block.h>
#include <gr_sync_block.h>
#include <gr_fir_util.h>
#include <gr_fir_filter_ccf.h>
#include <gr_fir_filter_fff.h>
#include <gr_freq_xlating_fir_filter_ccf.h>
#include <gr_pfb_channelizer_ccf.h>
#include <gr_pfb_arb_resampler_ccf.h>
#include <gr_pfb_decimator_ccf.h>
#include <gr_fft_filter_ccc.h>
#include <gr_framer_sink_1.h>
#include <gr_fft_vcc.h>
#include <gr_quadrature_demod_cf.h>
#include <gr_math.h>
#include <gr_count_bits.h>
#include <gr_firdes.h>
#include <gr_remez.h>
#include <gr_rotator.h>
#include <gr_map_bb.h>
#include <gr_unpack_k_bits_bb.h>
#include <gri_mmse_fir_interpolator.h>
#include <digital_clock_recovery_mm_ff.h>
#include <digital_clock_recovery_mm_cc.h>
#include <digital_correlate_access_code_bb.h>
#include <digital_binary_slicer_fb.h>
#include <pager_slicer_fb.h>

#include <cstring>
#include <string>
#include <vector>
#include <thread>
#include <chrono>
#include <atomic>

int g_grref_fir_impl = 1;  // 1 = SSE classes (what gr_fir_sysconfig_x86 picks), 0 = generic

static thread_local std::string g_err;

struct grref_block {
  gr_block_sptr blk;
};

template <class F>
static grref_block* guarded(F f) {
  try {
    gr_block_sptr b = f();
    grref_block* h = new grref_block;
    h->blk = b;
    return h;
  } catch (const std::invalid_argument& e) {
    g_err = std::string("invalid_argument: ") + e.what();
  } catch (const std::out_of_range& e) {
    g_err = std::string("out_of_range: ") + e.what();
  } catch (const std::exception& e) {
    g_err = std::string("exception: ") + e.what();
  }
  return nullptr;
}

extern "C" {

const char* grref_last_error() { return g_err.c_str(); }
void grref_set_fir_impl(int sse) { g_grref_fir_impl = sse; }
int grref_get_fir_impl() { return g_grref_fir_impl; }

// ---- constructors (reference factory signatures) -------------------------------------
grref_block* grref_make_fir_filter_ccf(int decim, const float* taps, int ntaps) {
  return guarded([&] { return gr_block_sptr(gr_make_fir_filter_ccf(decim, std::vector<float>(taps, taps + ntaps))); });
}
grref_block* grref_make_fir_filter_fff(int decim, const float* taps, int ntaps) {
  return guarded([&] { return gr_block_sptr(gr_make_fir_filter_fff(decim, std::vector<float>(taps, taps + ntaps))); });
}
grref_block* grref_make_freq_xlating_fir_filter_ccf(int decim, const float* taps, int ntaps,
                                                    double center_freq, double sampling_freq) {
  return guarded([&] {
    return gr_block_sptr(gr_make_freq_xlating_fir_filter_ccf(decim, std::vector<float>(taps, taps + ntaps),
                                                             center_freq, sampling_freq));
  });
}
grref_block* grref_make_pfb_channelizer_ccf(unsigned numchans, const float* taps, int ntaps, float oversample) {
  return guarded([&] {
    return gr_block_sptr(gr_make_pfb_channelizer_ccf(numchans, std::vector<float>(taps, taps + ntaps), oversample));
  });
}
grref_block* grref_make_pfb_arb_resampler_ccf(float rate, const float* taps, int ntaps, unsigned filter_size) {
  return guarded([&] {
    return gr_block_sptr(gr_make_pfb_arb_resampler_ccf(rate, std::vector<float>(taps, taps + ntaps), filter_size));
  });
}
grref_block* grref_make_pfb_decimator_ccf(unsigned decim, const float* taps, int ntaps, unsigned channel) {
  return guarded([&] { return gr_block_sptr(gr_make_pfb_decimator_ccf(decim, std::vector<float>(taps, taps + ntaps), channel)); });
}
grref_block* grref_make_fft_filter_ccc(int decim, const float* taps_ri, int ntaps) {
  return guarded([&] {
    std::vector<gr_complex> t(ntaps);
    for (int i = 0; i < ntaps; i++) t[i] = gr_complex(taps_ri[2 * i], taps_ri[2 * i + 1]);
    return gr_block_sptr(gr_make_fft_filter_ccc(decim, t));
  });
}
int grref_fft_filter_ccc_set_taps(grref_block* h, const float* taps_ri, int ntaps) {
  gr_fft_filter_ccc* b = dynamic_cast<gr_fft_filter_ccc*>(h->blk.get());
  if (!b) return -1;
  std::vector<gr_complex> t(ntaps);
  for (int i = 0; i < ntaps; i++) t[i] = gr_complex(taps_ri[2 * i], taps_ri[2 * i + 1]);
  b->set_taps(t);
  return 0;
}
// gr_framer_sink_1: the block and the queue it posts to live together; messages are drained through the C API
struct grref_framer_ctx { gr_msg_queue_sptr q; };
static std::vector<std::pair<grref_block*, gr_msg_queue_sptr> > g_framer_queues;
grref_block* grref_make_framer_sink_1() {
  gr_msg_queue_sptr q = gr_make_msg_queue();
  grref_block* b = guarded([&] { return gr_block_sptr(gr_make_framer_sink_1(q)); });
  if (b) g_framer_queues.push_back(std::make_pair(b, q));
  return b;
}
static gr_msg_queue_sptr framer_queue(grref_block* h) {
  for (size_t i = 0; i < g_framer_queues.size(); i++)
    if (g_framer_queues[i].first == h) return g_framer_queues[i].second;
  return gr_msg_queue_sptr();
}
int grref_framer_count(grref_block* h) { gr_msg_queue_sptr q = framer_queue(h); return q ? (int)q->count() : -1; }
// pops one message: returns its length (>= 0) or -1 if the queue is empty; *arg1 = whitener offset
int grref_framer_pop(grref_block* h, unsigned char* out, int cap, double* arg1) {
  gr_msg_queue_sptr q = framer_queue(h);
  if (!q) return -1;
  gr_message_sptr m = q->delete_head_nowait();
  if (!m) return -1;
  if (arg1) *arg1 = m->arg1();
  const int n = (int)m->length();
  memcpy(out, m->msg(), (size_t)(n < cap ? n : cap));
  return n;
}
void grref_framer_forget(grref_block* h) {
  for (size_t i = 0; i < g_framer_queues.size(); i++)
    if (g_framer_queues[i].first == h) { g_framer_queues.erase(g_framer_queues.begin() + i); return; }
}
grref_block* grref_make_fft_vcc(int fft_size, int forward, const float* window, int nwin, int shift) {
  return guarded([&] {
    return gr_block_sptr(gr_make_fft_vcc(fft_size, forward != 0, std::vector<float>(window, window + nwin), shift != 0));
  });
}
grref_block* grref_make_quadrature_demod_cf(float gain) {
  return guarded([&] { return gr_block_sptr(gr_make_quadrature_demod_cf(gain)); });
}
grref_block* grref_make_clock_recovery_mm_ff(float omega, float gain_omega, float mu, float gain_mu,
                                             float omega_relative_limit) {
  return guarded([&] {
    return gr_block_sptr(digital_make_clock_recovery_mm_ff(omega, gain_omega, mu, gain_mu, omega_relative_limit));
  });
}
grref_block* grref_make_clock_recovery_mm_cc(float omega, float gain_omega, float mu, float gain_mu,
                                             float omega_relative_limit) {
  return guarded([&] {
    return gr_block_sptr(digital_make_clock_recovery_mm_cc(omega, gain_omega, mu, gain_mu, omega_relative_limit));
  });
}
grref_block* grref_make_pager_slicer_fb(float alpha) {
  return guarded([&] { return gr_block_sptr(pager_make_slicer_fb(alpha)); });
}
grref_block* grref_make_binary_slicer_fb() {
  return guarded([&] { return gr_block_sptr(digital_make_binary_slicer_fb()); });
}
grref_block* grref_make_map_bb(const int* map, int n) {
  return guarded([&] { return gr_block_sptr(gr_make_map_bb(std::vector<int>(map, map + n))); });
}
grref_block* grref_make_unpack_k_bits_bb(unsigned k) {
  return guarded([&] { return gr_block_sptr(gr_make_unpack_k_bits_bb(k)); });
}
grref_block* grref_make_correlate_access_code_bb(const char* code, int threshold) {
  return guarded([&] { return gr_block_sptr(digital_make_correlate_access_code_bb(std::string(code), threshold)); });
}
void grref_block_delete(grref_block* h) { delete h; }

// ---- generic gr_block surface --------------------------------------------------------
unsigned grref_block_history(grref_block* h) { return h->blk->history(); }
int grref_block_output_multiple(grref_block* h) { return h->blk->output_multiple(); }
double grref_block_relative_rate(grref_block* h) { return h->blk->relative_rate(); }
int grref_block_consumed(grref_block* h) { return h->blk->consumed(); }
int grref_block_forecast(grref_block* h, int noutput, int ninputs) {
  gr_vector_int req(ninputs, 0);
  h->blk->forecast(noutput, req);
  return req.empty() ? 0 : req[0];
}
int grref_block_general_work(grref_block* h, int noutput, const int* ninput_items, int nin,
                             const void* const* in, void* const* out, int nout) {
  gr_vector_int ni(ninput_items, ninput_items + nin);
  gr_vector_const_void_star iv(in, in + nin);
  gr_vector_void_star ov(out, out + nout);
  return h->blk->general_work(noutput, ni, iv, ov);
}

// ---- setters ---------------------------------------------------------------------------
int grref_fir_filter_ccf_set_taps(grref_block* h, const float* taps, int n) {
  gr_fir_filter_ccf* b = dynamic_cast<gr_fir_filter_ccf*>(h->blk.get());
  if (!b) return -1;
  b->set_taps(std::vector<float>(taps, taps + n));
  return 0;
}
int grref_fir_filter_fff_set_taps(grref_block* h, const float* taps, int n) {
  gr_fir_filter_fff* b = dynamic_cast<gr_fir_filter_fff*>(h->blk.get());
  if (!b) return -1;
  b->set_taps(std::vector<float>(taps, taps + n));
  return 0;
}
int grref_freq_xlating_set_center_freq(grref_block* h, double f) {
  gr_freq_xlating_fir_filter_ccf* b = dynamic_cast<gr_freq_xlating_fir_filter_ccf*>(h->blk.get());
  if (!b) return -1;
  b->set_center_freq(f);
  return 0;
}
int grref_freq_xlating_set_taps(grref_block* h, const float* taps, int n) {
  gr_freq_xlating_fir_filter_ccf* b = dynamic_cast<gr_freq_xlating_fir_filter_ccf*>(h->blk.get());
  if (!b) return -1;
  b->set_taps(std::vector<float>(taps, taps + n));
  return 0;
}
int grref_pfb_set_taps(grref_block* h, const float* taps, int n) {
  gr_pfb_channelizer_ccf* b = dynamic_cast<gr_pfb_channelizer_ccf*>(h->blk.get());
  if (!b) return -1;
  b->set_taps(std::vector<float>(taps, taps + n));
  return 0;
}
int grref_fft_vcc_set_window(grref_block* h, const float* w, int n) {
  gr_fft_vcc* b = dynamic_cast<gr_fft_vcc*>(h->blk.get());
  if (!b) return -1;
  return b->set_window(std::vector<float>(w, w + n)) ? 1 : 0;
}
int grref_mm_get_state(grref_block* h, float* mu, float* omega) {
  digital_clock_recovery_mm_ff* b = dynamic_cast<digital_clock_recovery_mm_ff*>(h->blk.get());
  if (!b) return -1;
  *mu = b->mu();
  *omega = b->omega();
  return 0;
}
float grref_pager_slicer_dc_offset(grref_block* h) {
  pager_slicer_fb* b = dynamic_cast<pager_slicer_fb*>(h->blk.get());
  return b ? b->dc_offset() : 0.f;
}

// ---- scalar primitives -----------------------------------------------------------------
void grref_fast_atan2f(const float* y, const float* x, float* out, long n) {
  for (long i = 0; i < n; i++) out[i] = gr_fast_atan2f(y[i], x[i]);
}
void grref_mmse_interpolate(const float* in8, const float* mu, float* out, long n) {
  // in8: n rows of 8 floats; row i 16-byte aligned phase is the caller's business
  gri_mmse_fir_interpolator interp;
  for (long i = 0; i < n; i++) out[i] = interp.interpolate(in8 + 8 * i, mu[i]);
}
void grref_mmse_taps(float* out /* 129*8 */) {
  // dumps the reference's MMSE table through its public object: feed unit impulses
  gri_mmse_fir_interpolator interp;
  for (int s = 0; s <= 128; s++)
    for (int k = 0; k < 8; k++) {
      alignas(16) float imp[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      imp[k] = 1.0f;
      out[s * 8 + k] = interp.interpolate(imp, (float)s / 128.0f);
    }
}
unsigned grref_count_bits64(unsigned long long x) { return gr_count_bits64(x); }
void grref_rotator(float incr_re, float incr_im, const float* in, float* out, long n) {
  gr_rotator r;
  r.set_phase_incr(gr_complex(incr_re, incr_im));
  const gr_complex* ci = (const gr_complex*)in;
  gr_complex* co = (gr_complex*)out;
  for (long i = 0; i < n; i++) co[i] = r.rotate(ci[i]);
}
int grref_binary_slicer(float x) { return gr_binary_slicer(x); }
float grref_branchless_clip(float x, float clip) { return gr_branchless_clip(x, clip); }

// ---- gr_firdes -------------------------------------------------------------------------
static int copy_out(const std::vector<float>& v, float* out, int cap) {
  if ((int)v.size() > cap) return -(int)v.size();
  memcpy(out, v.data(), v.size() * sizeof(float));
  return (int)v.size();
}
int grref_firdes_low_pass(double gain, double fs, double fc, double tw, int win, double beta, float* out, int cap) {
  try { return copy_out(gr_firdes::low_pass(gain, fs, fc, tw, (gr_firdes::win_type)win, beta), out, cap); }
  catch (const std::exception& e) { g_err = e.what(); return 0; }
}
int grref_firdes_low_pass_2(double gain, double fs, double fc, double tw, double atten, int win, double beta,
                            float* out, int cap) {
  try { return copy_out(gr_firdes::low_pass_2(gain, fs, fc, tw, atten, (gr_firdes::win_type)win, beta), out, cap); }
  catch (const std::exception& e) { g_err = e.what(); return 0; }
}
int grref_firdes_root_raised_cosine(double gain, double fs, double sym, double alpha, int ntaps, float* out, int cap) {
  try { return copy_out(gr_firdes::root_raised_cosine(gain, fs, sym, alpha, ntaps), out, cap); }
  catch (const std::exception& e) { g_err = e.what(); return 0; }
}
int grref_firdes_window(int win, int ntaps, double beta, float* out, int cap) {
  try { return copy_out(gr_firdes::window((gr_firdes::win_type)win, ntaps, beta), out, cap); }
  catch (const std::exception& e) { g_err = e.what(); return 0; }
}


// gr_remez (general/gr_remez.cc:792-877): returns the number of taps written, or -1 when the reference throws
int ref_remez(int order, const double* bands, int nbands2, const double* ampl, const double* weight, int nweight,
              const char* type, int grid_density, double* out) {
  try {
    std::vector<double> t = gr_remez(order, std::vector<double>(bands, bands + nbands2), std::vector<double>(ampl, ampl + nbands2),
                                     std::vector<double>(weight, weight + nweight), type, grid_density);
    for (size_t i = 0; i < t.size(); i++) out[i] = t[i];
    return (int)t.size();
  } catch (std::exception&) {
    return -1;
  }
}
}  // extern "C"
