/* ORACLE / TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C CPU restatement of the reference algorithms on the channelize + DMR-demod hot path
 * (SURVEY.md section 8a).  It is the checker for the CUDA path: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg may build, link, load or call it.
 * The product (libgr_cuda, grb200) never does.
 *
 * Parity status: PINNED.  Every function below is checked in tests/test_oracle_*.py against
 *  (1) the golden vectors / known-answer tests the reference ships for this path
 *      (qa_gr_fir_fff.cc, qa_fft.py, qa_clock_recovery_mm.py, qa_correlate_access_code.py,
 *      qa_gr_math.cc, qa_gr_rotator.cc, qa_gri_mmse_fir_interpolator.cc, qa_gr_firdes.cc), and
 *  (2) oracle/_ref/libgrref.so = the reference's own sources compiled in place.
 *
 * Buffer convention = the reference runtime's: `in` points at the first HISTORY item
 * (history-1 items before the first new item), see gr_flat_flowgraph.cc:150.
 *
 * Summation-order modes for float FIRs (the reference's result depends on which gr_fir
 * implementation gr_fir_sysconfig_x86.cc:175-201 selects):
 *   ORC_ORDER_GENERIC  gr_fir_XXX_generic.cc.t:28-81 (4 accumulators fff / 2 ccf)
 *   ORC_ORDER_SSE      float_dotprod_sse64.S:27-108 driven by gr_fir_fff_simd.cc:99-134; the
 *                      order depends on the 16-byte phase of the input pointer, which in a
 *                      page-aligned gr_buffer equals (absolute item index) mod 4.
 */
#ifndef ORACLE_H
#define ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

#define ORC_ORDER_GENERIC 0
#define ORC_ORDER_SSE 1

typedef struct { float re, im; } orc_cpx;

/* gr_fir_ccf::filterNdec (gr_fir_XXX_generic.cc.t:57-103, ccf: N_UNROLL=2). */
void orc_fir_ccf(const float* taps, int ntaps, int decim, const orc_cpx* in, long nout, orc_cpx* out);
/* gr_fir_fff::filterNdec.  abs0 = absolute stream index of in[0] (only used by ORC_ORDER_SSE). */
void orc_fir_fff(const float* taps, int ntaps, int decim, const float* in, long nout, float* out,
                 int order, long abs0);
/* gr_fir_ccc generic (complex taps), used by freq_xlating. */
void orc_fir_ccc(const orc_cpx* taps, int ntaps, int decim, const orc_cpx* in, long nout, orc_cpx* out);

/* gr_freq_xlating_fir_filter_ccf (gr_freq_xlating_fir_filter_XXX.cc.t:72-83,99-123) with the
 * gr_rotator recurrence (gr_rotator.h:40-50).  State carried in rot. */
typedef struct { orc_cpx phase, incr; unsigned counter; } orc_rotator;
void orc_rotator_init(orc_rotator* r, orc_cpx incr);
void orc_rotator_init_f(orc_rotator* r, float incr_re, float incr_im);
orc_cpx orc_rotator_rotate(orc_rotator* r, orc_cpx in);
void orc_rotate_n(orc_rotator* r, const orc_cpx* in, long n, orc_cpx* out);
void orc_freq_xlating_taps(const float* proto, int ntaps, double center_freq, double sampling_freq, int decim,
                           orc_cpx* ctaps_fwd /* ntaps, forward order as handed to set_taps */,
                           orc_cpx* phase_incr);
void orc_freq_xlating_fir_ccf(const float* proto, int ntaps, int decim, double center_freq, double sampling_freq,
                              orc_rotator* rot, const orc_cpx* in, long nout, orc_cpx* out);

/* gri_fft_complex::execute (gri_fft.cc:142-146): unnormalised DFT, sign -1 forward / +1 backward,
 * evaluated in float64 and rounded once (FFTW3f itself is third-party and absent). */
void orc_dft(const orc_cpx* in, orc_cpx* out, int n, int forward);

/* gr_pfb_channelizer_ccf::general_work (gr_pfb_channelizer_ccf.cc:155-200) incl. oversampling
 * (:81-92,169-196).  ins[j] points at stream j's first history item (history = T+1).
 * Returns noutput; *consumed = items consumed per stream. */
int orc_pfb_channelizer_ccf(int numchans, const float* taps, int ntaps, float oversample_rate,
                            const orc_cpx* const* ins, int noutput, orc_cpx* out, int* consumed);
/* metadata the constructor computes (:57-62,81-92,104-139) */
int orc_pfb_taps_per_filter(int numchans, int ntaps);
int orc_pfb_output_multiple(int numchans, float oversample_rate);
int orc_pfb_check_rate(int numchans, float oversample_rate); /* 1 ok, 0 -> std::invalid_argument */

/* gr_fft_vcc_fftw::work (gr_fft_vcc_fftw.cc:51-103). window may be NULL (nwin = 0). */
void orc_fft_vcc(int fft_size, int forward, const float* window, int nwin, int shift, const orc_cpx* in,
                 long nvec, orc_cpx* out);

/* gr_fast_atan2f (gr_fast_atan2f.cc:125-198). */
float orc_fast_atan2f(float y, float x);
/* gr_quadrature_demod_cf::work (gr_quadrature_demod_cf.cc:46-62); in[0] is the history item. */
void orc_quadrature_demod_cf(float gain, const orc_cpx* in, long nout, float* out);

/* gri_mmse_fir_interpolator::interpolate (gri_mmse_fir_interpolator.cc:61-71) */
float orc_mmse_interpolate(const float* in8, float mu, int order, long abs0);

/* digital_clock_recovery_mm_ff (digital_clock_recovery_mm_ff.cc:48-68,102-139; header :75-80). */
typedef struct {
  float mu, omega, min_omega, omega_mid, max_omega, gain_omega, gain_mu, last_sample, omega_relative_limit;
} orc_mm_state;
int orc_mm_init(orc_mm_state* s, float omega, float gain_omega, float mu, float gain_mu, float omega_relative_limit);
int orc_mm_forecast(const orc_mm_state* s, int noutput);
/* returns produced; *consumed = ii.  abs0 = absolute index of in[0]. */
int orc_mm_general_work(orc_mm_state* s, const float* in, int ninput, float* out, int noutput, int* consumed,
                        int order, long abs0);

/* pager_slicer_fb::slice (pager_slicer_fb.cc:47-69) and gr_binary_slicer (gr_math.h:82-88). */
typedef struct { float alpha, beta, avg; } orc_slicer4_state;
void orc_slicer4_init(orc_slicer4_state* s, float alpha);
void orc_slicer4(orc_slicer4_state* s, const float* in, long n, unsigned char* out);
void orc_binary_slicer(const float* in, long n, unsigned char* out);

/* gr_map_bb::work (gr_map_bb.cc:49-61), gr_unpack_k_bits_bb::work (gr_unpack_k_bits_bb.cc:53-70). */
void orc_map_bb(const int* map, int nmap, const unsigned char* in, long n, unsigned char* out);
void orc_unpack_k_bits_bb(unsigned k, const unsigned char* in, long nin, unsigned char* out);

/* digital_correlate_access_code_bb (digital_correlate_access_code_bb.cc:64-133). */
typedef struct {
  unsigned long long access_code, data_reg, flag_reg, flag_bit, mask;
  unsigned threshold;
} orc_corr_state;
int orc_corr_init(orc_corr_state* s, const char* access_code, int threshold); /* 0 ok, -1 -> out_of_range */
void orc_corr_work(orc_corr_state* s, const unsigned char* in, long n, unsigned char* out);
unsigned orc_count_bits64(unsigned long long x); /* gr_count_bits.cc:75-93 */

/* gr_pfb_arb_resampler_ccf (gr_pfb_arb_resampler_ccf.cc:42-84 ctor, :90-125 create_taps, :127-140
 * create_diff_taps, :155-205 general_work; header :159-163 set_rate).  history = taps_per_filter + 1.
 * Filters use the ccf dot product in the generic order (2 accumulators); the SSE class differs by ~1e-7. */
typedef struct {
  unsigned int_rate, dec_rate, last_filter, taps_per_filter;
  float flt_rate, acc, rate;
  int start_index, updated;
  float* taps;  /* [int_rate][taps_per_filter]  taps of filter i in FORWARD order (as handed to set_taps) */
  float* dtaps; /* same for the derivative filters */
} orc_arb_state;
int orc_arb_init(orc_arb_state* s, float rate, const float* taps, int ntaps, unsigned filter_size);
void orc_arb_set_rate(orc_arb_state* s, float rate);
void orc_arb_free(orc_arb_state* s);
/* returns produced; *consumed = what consume_each got.  in[0] = first history item. */
int orc_arb_general_work(orc_arb_state* s, const orc_cpx* in, int ninput, orc_cpx* out, int noutput, int* consumed);
/* the index recurrence alone: which input offset / filter / interpolation weight output i uses.
 * Arrays may be NULL.  Same return values and state update as general_work. */
int orc_arb_schedule(orc_arb_state* s, int ninput, int noutput, int* count_of, unsigned short* filt_of, float* acc_of,
                     int* consumed);
/* gr_pfb_decimator_ccf::work (gr_pfb_decimator_ccf.cc:44-65 ctor, :75-110 set_taps, :127-175 work): decim
 * branch filters (gr_fir_ccf, generic order here) feeding a decim-point BACKWARD DFT of which bin `channel` is
 * kept.  ins[s] = stream s from its first history item (history = taps_per_filter).  The DFT bin is
 * evaluated in float64 and rounded once, like orc_dft (FFTW is third party and absent). */
int orc_pfb_decimator_taps_per_filter(int decim, int ntaps);
void orc_pfb_decimator_ccf(int decim, const float* taps, int ntaps, unsigned channel, const orc_cpx* const* ins,
                           long noutput, orc_cpx* out);
/* gr_fft_filter_ccc / gri_fft_filter_ccc_generic (gr_fft_filter_ccc.cc:46-106, gri_fft_filter_ccc_generic.cc:62-165):
 * overlap-add FFT filter with complex taps.  fftsize = 2 * 2^ceil(log2 ntaps), nsamples = fftsize - ntaps + 1
 * (= the block's output_multiple), taps pre-scaled by 1/fftsize, tail of ntaps-1 carried between blocks.  The
 * transforms are the float64 DFT of orc_dft (FFTW absent). */
typedef struct {
  int ntaps, fftsize, nsamples, decimation;
  orc_cpx* xformed_taps; /* [fftsize] */
  orc_cpx* tail;         /* [ntaps-1] */
} orc_fftfilt_state;
int orc_fftfilt_init(orc_fftfilt_state* s, int decimation, const orc_cpx* taps, int ntaps); /* returns nsamples */
int orc_fftfilt_set_taps(orc_fftfilt_state* s, const orc_cpx* taps, int ntaps);             /* returns nsamples */
void orc_fftfilt_free(orc_fftfilt_state* s);
/* nitems outputs (multiple of nsamples) from nitems * decimation inputs; returns nitems */
int orc_fftfilt_filter(orc_fftfilt_state* s, int nitems, const orc_cpx* in, orc_cpx* out);
/* gr_framer_sink_1::work (gr_framer_sink_1.cc:93-196; header checks gr_framer_sink_1.h:88-103): consumes the
 * correlator's output bytes (bit 0 = data, bit 1 = "access code ended here"), reads a 32-bit header (two identical
 * 16-bit halves: whitener offset << 12 | payload length), then payload_len bytes MSB first, and posts a message per
 * packet.  emit(ctx, whitener_offset, payload, len) stands for d_target_queue->insert_tail(). */
typedef struct {
  int state; /* 0 sync search, 1 have sync, 2 have header */
  unsigned header;
  int headerbitlen_cnt, packetlen, whitener_offset, packetlen_cnt, byte_index;
  unsigned char packet_byte;
  unsigned char packet[4096];
} orc_framer_state;
typedef void (*orc_framer_emit)(void* ctx, int whitener_offset, const unsigned char* payload, int len);
void orc_framer_init(orc_framer_state* s);
void orc_framer_work(orc_framer_state* s, const unsigned char* in, long n, orc_framer_emit emit, void* ctx);
/* digital_clock_recovery_mm_cc (gr-digital/lib/digital_clock_recovery_mm_cc.cc:53-75 ctor, :123-218 general_work;
 * set_omega gr-digital/include/digital_clock_recovery_mm_cc.h:75-80) with gri_mmse_fir_interpolator_cc
 * (filter/gri_mmse_fir_interpolator_cc.cc:62-71: gr_fir_ccf on the 8-tap MMSE rows; generic order here).
 * err may be NULL: the reference then clips the timing error to +-1 instead of +-4 (:146 vs :181). */
typedef struct {
  float mu, omega, gain_omega, gain_mu, omega_mid, omega_relative_limit;
  orc_cpx p_2T, p_1T, p_0T, c_2T, c_1T, c_0T;
} orc_mmcc_state;
int orc_mmcc_init(orc_mmcc_state* s, float omega, float gain_omega, float mu, float gain_mu, float omega_relative_limit);
int orc_mmcc_forecast(const orc_mmcc_state* s, int noutput);
int orc_mmcc_general_work(orc_mmcc_state* s, const orc_cpx* in, int ninput, orc_cpx* out, float* err, int noutput,
                          int* consumed);
/* gr_firdes (gr_firdes.cc:57-147,601-655,720-782): tap / window design on the host. */
int orc_firdes_window(int win_type, int ntaps, double beta, float* out);
int orc_firdes_low_pass(double gain, double fs, double fc, double tw, int win_type, double beta, float* out, int cap);
int orc_firdes_low_pass_2(double gain, double fs, double fc, double tw, double atten_db, int win_type, double beta,
                          float* out, int cap);
int orc_firdes_root_raised_cosine(double gain, double fs, double sym_rate, double alpha, int ntaps, float* out);

/* tables (from build/generated/gr_tables.h) */
const float* orc_mmse_table(void);  /* [129][8], reference storage order */
const float* orc_atan_table(void);  /* [257] */

#ifdef __cplusplus
}
#endif
#endif
