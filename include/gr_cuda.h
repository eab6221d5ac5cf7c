/* libgr_cuda -- C ABI of the B200-native channelize + DMR-demod hot path.
 *
 * This header is the drop-in boundary: every entry point is what a
 * GNU Radio 3.5 block on this path would bind instead of its CPU primitive.  Plain pointers
 * and sizes only; no C++ / torch types; no exception crosses the ABI.
 *
 * Conventions
 *  - Every block family is an opaque plan handle: *_create / *_destroy / setters / getters /
 *    *_work (HOST pointers, reference buffer layout, staged through pinned double buffers) /
 *    *_work_device (DEVICE pointers, same layout, for HBM-resident pipelines).
 *  - Buffer layout = the reference runtime's: the input pointer addresses the first HISTORY
 *    item, i.e. history()-1 items before the first new item
 *    (gnuradio-core/src/lib/runtime/gr_flat_flowgraph.cc:150, gr_buffer.cc:201-214).
 *  - *_work returns the number of items produced (>= 0) or a negative GRCUDA_E* code.
 *    Setters are thread safe and take effect at the next work() boundary; the first work()
 *    after set_taps returns 0 exactly like the reference ("history requirements may have
 *    changed", gr_fir_filter_XXX.cc.t:74-79, gr_pfb_channelizer_ccf.cc:164-167).
 *  - *_work_device returns GRCUDA_OK (0) or a negative GRCUDA_E* code; it always produces exactly
 *    noutput_items (the caller owns the buffers and their sizes).  A pending setter is applied at
 *    the start of the call (after a device synchronisation), the call then proceeds: the "return 0
 *    once" of the scheduler-facing work() has no meaning for a caller that sizes its own buffers,
 *    but history() may have changed -- re-read it after a set_taps.
 *  - *_create returns NULL on error; grcuda_last_error() / grcuda_last_error_code() tell why.
 *    GRCUDA_EINVAL maps to std::invalid_argument, GRCUDA_ERANGE to std::out_of_range,
 *    GRCUDA_ECUDA to std::runtime_error in the C++ block wrappers (blocks/gr_b200_blocks.h).
 *  - `stream` arguments are cudaStream_t passed as void* (NULL = the plan's own stream).
 *  - There is NO CPU fallback: without a CUDA device every create/work fails with GRCUDA_ECUDA.
 */
#ifndef INCLUDED_GR_CUDA_H
#define INCLUDED_GR_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library itself is built with -fvisibility=hidden */
#endif

#define GRCUDA_OK 0
#define GRCUDA_EINVAL (-1) /* std::invalid_argument in the reference */
#define GRCUDA_ERANGE (-2) /* std::out_of_range in the reference */
#define GRCUDA_ECUDA (-3)  /* CUDA runtime / launch failure, or no device */
#define GRCUDA_ENOMEM (-4)
#define GRCUDA_EUNSUPPORTED (-5)

/* summation order of the float dot products in the demod tail (see oracle/oracle.h) */
#define GRCUDA_ORDER_GENERIC 0 /* gr_fir_fff_generic: 4 accumulators, ((a0+a1)+a2)+a3 */
#define GRCUDA_ORDER_SSE 1     /* float_dotprod_sse64.S order = what x86-64 GNU Radio runs */

typedef struct { float re, im; } grcuda_complex; /* == gr_complex (runtime/gr_complex.h:26) */

/* ---- library / device -------------------------------------------------------------------- */
const char* grcuda_version(void);
const char* grcuda_last_error(void);
int grcuda_last_error_code(void);
int grcuda_device_count(void);
int grcuda_set_device(int device);
int grcuda_device_synchronize(void);
/* device / pinned memory helpers for HBM-resident callers that do not use torch */
void* grcuda_malloc_device(size_t bytes);
void grcuda_free_device(void* p);
void* grcuda_malloc_pinned(size_t bytes);
void grcuda_free_pinned(void* p);
int grcuda_memcpy_h2d(void* dst, const void* src, size_t bytes, void* stream);
int grcuda_memcpy_d2h(void* dst, const void* src, size_t bytes, void* stream);
int grcuda_stream_synchronize(void* stream);
/* number of kernels this library has launched since load (bench.py's gpu_launches) */
unsigned long long grcuda_kernel_launch_count(void);

/* ---- a1  gr_fir_filter_ccf -----------------------------------------------------------------
 * replaces gr_fir_filter_ccf / gr_fir_ccf{,_simd,_x86} + fcomplex_dotprod_sse64.S
 * (gnuradio-core/src/lib/filter/gr_fir_filter_XXX.cc.t:37-88, gr_fir_ccf_simd.cc:101-141).
 * out[i] = sum_k taps[k] * x[i*D - k];  history = ntaps. */
typedef struct grcuda_fir_ccf grcuda_fir_ccf;
grcuda_fir_ccf* grcuda_fir_filter_ccf_create(int decimation, const float* taps, int ntaps);
void grcuda_fir_filter_ccf_destroy(grcuda_fir_ccf* h);
int grcuda_fir_filter_ccf_set_taps(grcuda_fir_ccf* h, const float* taps, int ntaps);
unsigned grcuda_fir_filter_ccf_history(grcuda_fir_ccf* h);
int grcuda_fir_filter_ccf_decimation(grcuda_fir_ccf* h);
int grcuda_fir_filter_ccf_work(grcuda_fir_ccf* h, int noutput_items, const grcuda_complex* in, grcuda_complex* out);
int grcuda_fir_filter_ccf_work_device(grcuda_fir_ccf* h, long noutput_items, const grcuda_complex* d_in,
                                      grcuda_complex* d_out, void* stream);

/* ---- a2  gr_fir_filter_fff -----------------------------------------------------------------
 * replaces gr_fir_filter_fff / gr_fir_fff{,_simd} + float_dotprod_sse64.S
 * (gr_fir_fff_simd.cc:99-134).  `order` selects the reference summation order reproduced
 * bit for bit; `abs_index0` = absolute stream index of in[0] (the SSE order depends on it). */
typedef struct grcuda_fir_fff grcuda_fir_fff;
grcuda_fir_fff* grcuda_fir_filter_fff_create(int decimation, const float* taps, int ntaps, int order);
void grcuda_fir_filter_fff_destroy(grcuda_fir_fff* h);
int grcuda_fir_filter_fff_set_taps(grcuda_fir_fff* h, const float* taps, int ntaps);
unsigned grcuda_fir_filter_fff_history(grcuda_fir_fff* h);
int grcuda_fir_filter_fff_work(grcuda_fir_fff* h, int noutput_items, const float* in, float* out, long abs_index0);
/* batched: nchan independent streams laid out [time][channel] (row stride = nchan floats) */
int grcuda_fir_filter_fff_work_device(grcuda_fir_fff* h, long noutput_items, int nchan, const float* d_in,
                                      float* d_out, long abs_index0, void* stream);

/* ---- a3  gr_freq_xlating_fir_filter_ccf ----------------------------------------------------
 * replaces gr_freq_xlating_fir_filter_ccf + gr_fir_ccc + gr_rotator
 * (gr_freq_xlating_fir_filter_XXX.cc.t:38-123, gr_rotator.h:40-50). */
typedef struct grcuda_fxlat grcuda_fxlat;
grcuda_fxlat* grcuda_freq_xlating_fir_filter_ccf_create(int decimation, const float* taps, int ntaps,
                                                       double center_freq, double sampling_freq);
void grcuda_freq_xlating_fir_filter_ccf_destroy(grcuda_fxlat* h);
int grcuda_freq_xlating_fir_filter_ccf_set_taps(grcuda_fxlat* h, const float* taps, int ntaps);
int grcuda_freq_xlating_fir_filter_ccf_set_center_freq(grcuda_fxlat* h, double center_freq);
unsigned grcuda_freq_xlating_fir_filter_ccf_history(grcuda_fxlat* h);
int grcuda_freq_xlating_fir_filter_ccf_work(grcuda_fxlat* h, int noutput_items, const grcuda_complex* in,
                                            grcuda_complex* out);
int grcuda_freq_xlating_fir_filter_ccf_work_device(grcuda_fxlat* h, long noutput_items, const grcuda_complex* d_in,
                                                   grcuda_complex* d_out, void* stream);

/* ---- a4/a5/a14  gr_pfb_channelizer_ccf -----------------------------------------------------
 * replaces gr_pfb_channelizer_ccf::general_work, its numchans gr_fir_ccf objects and the
 * FFTW-backed gri_fft_complex (gr_pfb_channelizer_ccf.cc:35-200, gri_fft.cc:97-146).
 * create() fails with GRCUDA_EINVAL when numchans/oversample_rate is not an integer (:57-60).
 * history = taps_per_filter + 1, output_multiple and relative_rate as the reference computes. */
typedef struct grcuda_pfb grcuda_pfb;
grcuda_pfb* grcuda_pfb_channelizer_ccf_create(unsigned numchans, const float* taps, int ntaps, float oversample_rate);
void grcuda_pfb_channelizer_ccf_destroy(grcuda_pfb* h);
int grcuda_pfb_channelizer_ccf_set_taps(grcuda_pfb* h, const float* taps, int ntaps);
unsigned grcuda_pfb_channelizer_ccf_history(grcuda_pfb* h);
int grcuda_pfb_channelizer_ccf_output_multiple(grcuda_pfb* h);
double grcuda_pfb_channelizer_ccf_relative_rate(grcuda_pfb* h);
int grcuda_pfb_channelizer_ccf_taps_per_filter(grcuda_pfb* h);
/* general_work: `in` = numchans host stream pointers (each history-prefixed), out = one stream
 * of numchans-wide vectors; *consumed = items consumed per input stream. */
int grcuda_pfb_channelizer_ccf_work(grcuda_pfb* h, int noutput_items, const grcuda_complex* const* in,
                                    grcuda_complex* out, int* consumed);
/* blks2.pfb_channelizer_ccf hier-block form (blks2impl/pfb_channelizer.py:61-75): ONE
 * interleaved wideband stream in (what gr_stream_to_streams would split), [time][channel] out.
 * `in` addresses the first history ROW ((history-1) rows of numchans items before the new rows). */
int grcuda_pfb_channelizer_ccf_work_interleaved(grcuda_pfb* h, int noutput_items, const grcuda_complex* in,
                                                grcuda_complex* out, int* consumed);
int grcuda_pfb_channelizer_ccf_work_device(grcuda_pfb* h, long noutput_items, const grcuda_complex* d_in_rows,
                                           grcuda_complex* d_out, void* stream);

/* ---- a5/a6  gr_fft_vcc ---------------------------------------------------------------------
 * replaces gr_fft_vcc_fftw::work + gri_fft_complex (gr_fft_vcc_fftw.cc:51-103, gr_fft_vcc.cc:34-64).
 * fft_size <= 0 -> GRCUDA_ERANGE (gri_fft.cc:104-105). set_window returns 0 (false) unless the
 * window is empty or fft_size long (gr_fft_vcc.cc:55-64). */
typedef struct grcuda_fft grcuda_fft;
grcuda_fft* grcuda_fft_vcc_create(int fft_size, int forward, const float* window, int nwindow, int shift);
void grcuda_fft_vcc_destroy(grcuda_fft* h);
int grcuda_fft_vcc_set_window(grcuda_fft* h, const float* window, int nwindow);
int grcuda_fft_vcc_work(grcuda_fft* h, int noutput_items, const grcuda_complex* in, grcuda_complex* out);
int grcuda_fft_vcc_work_device(grcuda_fft* h, long noutput_items, const grcuda_complex* d_in, grcuda_complex* d_out,
                               void* stream);

/* ---- a7/a8  gr_quadrature_demod_cf (+ gr_fast_atan2f) --------------------------------------
 * replaces gr_quadrature_demod_cf::work (gr_quadrature_demod_cf.cc:46-62) with the reference's
 * 257-entry table arctangent (gr_fast_atan2f.cc:125-198), bit exact.  history = 2. */
typedef struct grcuda_quad grcuda_quad;
grcuda_quad* grcuda_quadrature_demod_cf_create(float gain);
void grcuda_quadrature_demod_cf_destroy(grcuda_quad* h);
int grcuda_quadrature_demod_cf_set_gain(grcuda_quad* h, float gain);
float grcuda_quadrature_demod_cf_gain(grcuda_quad* h);
int grcuda_quadrature_demod_cf_work(grcuda_quad* h, int noutput_items, const grcuda_complex* in, float* out);
/* batched [time][channel]; d_in addresses the history row */
int grcuda_quadrature_demod_cf_work_device(grcuda_quad* h, long noutput_items, int nchan, const grcuda_complex* d_in,
                                           float* d_out, void* stream);
int grcuda_fast_atan2f_device(const float* d_y, const float* d_x, float* d_out, long n, void* stream);

/* ---- a7+a8+a2 fused: gr_quadrature_demod_cf -> gr_fir_filter_fff, batched [time][channel] ------
 * One pass over the channelizer output (gr_quadrature_demod_cf.cc:46-62 feeding
 * gr_fir_filter_XXX.cc.t:66-88 / float_dotprod_sse64.S): 8 B read + 4 B written per channel sample,
 * the discriminator output stays in shared memory.  Bit identical to running the two blocks one
 * after the other in the SSE summation order (GRCUDA_EUNSUPPORTED for GRCUDA_ORDER_GENERIC or more
 * than 129 taps).  d_in addresses the first of history() rows that precede the nrows new rows
 * (zeros at stream start); abs_row0 = absolute stream index of the first NEW row. */
int grcuda_quad_demod_fir_fff_history(grcuda_fir_fff* f);
int grcuda_quad_demod_fir_fff_work_device(grcuda_quad* q, grcuda_fir_fff* f, long nrows, int nchan,
                                          const grcuda_complex* d_in, float* d_out, long abs_row0, void* stream);

/* ---- a9/a10/a11  digital_clock_recovery_mm_ff (+ gri_mmse_fir_interpolator, slicers) --------
 * replaces digital_clock_recovery_mm_ff::general_work (digital_clock_recovery_mm_ff.cc:102-139)
 * and gri_mmse_fir_interpolator::interpolate (gri_mmse_fir_interpolator.cc:61-71), batched over
 * nchan independent channels.  create() fails with GRCUDA_ERANGE for omega < 1 or negative gains
 * (:58-61).  The loop state (mu, omega, last_sample, next input index) lives on the device per
 * channel and persists across work calls, like the block's members do. */
typedef struct grcuda_mm grcuda_mm;
grcuda_mm* grcuda_clock_recovery_mm_ff_create(int nchan, float omega, float gain_omega, float mu, float gain_mu,
                                             float omega_relative_limit, int order);
void grcuda_clock_recovery_mm_ff_destroy(grcuda_mm* h);
int grcuda_clock_recovery_mm_ff_forecast(grcuda_mm* h, int noutput_items);
/* per-channel state readback / overwrite (mirrors mu(), omega(), set_mu(), set_omega()) */
int grcuda_clock_recovery_mm_ff_get_state(grcuda_mm* h, int chan, float* mu, float* omega, float* last_sample);
int grcuda_clock_recovery_mm_ff_set_mu(grcuda_mm* h, float mu);
int grcuda_clock_recovery_mm_ff_set_omega(grcuda_mm* h, float omega);
int grcuda_clock_recovery_mm_ff_set_gain_mu(grcuda_mm* h, float gain_mu);
int grcuda_clock_recovery_mm_ff_set_gain_omega(grcuda_mm* h, float gain_omega);
/* single-stream general_work (nchan must be 1): returns produced, *consumed = items consumed.
 * abs_index0 = absolute stream index of in[0]. */
int grcuda_clock_recovery_mm_ff_work(grcuda_mm* h, int noutput_items, int ninput_items, const float* in, float* out,
                                     int* consumed, long abs_index0);
/* batched device form.  d_in: [time][channel] floats, ninput rows; row 0 has absolute index
 * abs_row0.  Channel c starts reading at its own carried position (>= abs_row0) and stops when
 * fewer than 8 look-ahead rows remain (:112,116).  Symbols are written time-major to
 * d_out[sym][channel] (row capacity max_out); d_counts[c] receives the number produced.
 * If d_slice_out != NULL the 4-level (or 2-level) slicer decision is stored alongside. */
int grcuda_clock_recovery_mm_ff_work_device(grcuda_mm* h, long ninput_rows, long abs_row0, const float* d_in,
                                            float* d_out, unsigned char* d_slice_out, int max_out, int* d_counts,
                                            void* stream);
/* slicer fused into the M&M epilogue: mode 0 none, 2 = gr_binary_slicer (gr_math.h:82-88),
 * 4 = pager_slicer_fb::slice with DC-tracking alpha (pager_slicer_fb.cc:47-69). */
int grcuda_clock_recovery_mm_ff_set_slicer(grcuda_mm* h, int levels, float alpha);
/* Sums over channels of the loop's two error counters (synchronises the device): `clamped` = steps that wanted to go
 * back before the first row the caller buffered (the reference would re-read older items of its circular buffer),
 * `overflow` = work calls that stopped at max_out before the input ran out.  Either one non-zero means the output has
 * left the reference's: give the block more look-back / capacity. */
int grcuda_clock_recovery_mm_ff_counters(grcuda_mm* h, long long* clamped, long long* overflow);
/* Which build of the clock-recovery kernel runs (all produce identical bits; tests/test_gpu_blocks.py checks them
 * against the oracle).  -1 (default) = automatic: the quad-ring kernel (kernel_mm_quad.cuh, 198 KB of shared memory,
 * one CTA per SM) when the channel count is a multiple of 4 and the input 16-byte aligned, else the per-lane loader.
 * 0 = the round-1 kernel, 10 = its successor at the same 48 registers / 47 KB, sized to co-reside with the front
 * kernels of a single-GPU chain; 11 = 64 registers; 20-22 = quad ring at 80 / 64 / 96 registers; the rest are the
 * steps in between (profiles/README.md).  GRCUDA_EINVAL for an unknown number. */
int grcuda_clock_recovery_mm_ff_set_kernel_variant(grcuda_mm* h, int variant);
#define GRCUDA_MM_VARIANTS 23

/* stand-alone slicer blocks (host pointers) */
typedef struct grcuda_slicer grcuda_slicer;
grcuda_slicer* grcuda_pager_slicer_fb_create(float alpha);
grcuda_slicer* grcuda_binary_slicer_fb_create(void);
void grcuda_slicer_destroy(grcuda_slicer* h);
float grcuda_pager_slicer_fb_dc_offset(grcuda_slicer* h);
int grcuda_slicer_work(grcuda_slicer* h, int noutput_items, const float* in, unsigned char* out);

/* ---- a12/a13  gr_map_bb + gr_unpack_k_bits_bb + digital_correlate_access_code_bb ------------
 * replaces digital_correlate_access_code_bb::work + gr_count_bits64
 * (digital_correlate_access_code_bb.cc:64-133, gr_count_bits.cc:75-93).  access_code: <= 64
 * chars, LSB of each char (else GRCUDA_ERANGE, :54-57).  Output byte = data bit delayed 64 |
 * flag << 1, identical to the reference byte stream; sync hits are also compacted to a list. */
typedef struct grcuda_corr grcuda_corr;
grcuda_corr* grcuda_correlate_access_code_bb_create(int nchan, const char* access_code, int threshold);
void grcuda_correlate_access_code_bb_destroy(grcuda_corr* h);
int grcuda_correlate_access_code_bb_set_access_code(grcuda_corr* h, const char* access_code);
int grcuda_correlate_access_code_bb_work(grcuda_corr* h, int noutput_items, const unsigned char* in,
                                         unsigned char* out);
/* batched device form fed by the slicer: symbols [sym][channel] (values 0..3, d_counts[c] valid
 * per channel) -> gr_map_bb(map, nmap) -> gr_unpack_k_bits_bb(bits_per_symbol) -> correlator.
 * d_out: [bit][channel] bytes (may be NULL to skip the byte stream).  Hits are appended to
 * d_hits as {channel, absolute bit index of the flagged byte} pairs, *d_nhits counts them. */
typedef struct { int channel; int pad; long long bit_index; } grcuda_hit;
int grcuda_correlate_access_code_bb_work_symbols_device(grcuda_corr* h, const unsigned char* d_symbols, int sym_rows,
                                                        const int* d_counts, const int* map, int nmap,
                                                        int bits_per_symbol, unsigned char* d_out, int out_rows,
                                                        grcuda_hit* d_hits, int max_hits, int* d_nhits,
                                                        void* stream);

/* ---- 8f rank 3  gr_pfb_arb_resampler_ccf ------------------------------------------------------
 * replaces gr_pfb_arb_resampler_ccf::general_work and its 2 x filter_size gr_fir_ccf objects
 * (gr_pfb_arb_resampler_ccf.cc:42-205; set_rate .h:159-163): the 12.5 kS/s -> integer samples/symbol
 * resampler the reference's own 4FSK chains put in front of clock recovery.  Constructor arguments of
 * gr_make_pfb_arb_resampler_ccf(rate, taps, filter_size) + nchan: the batched form resamples nchan
 * channels laid out [time][channel] with ONE schedule (nchan = 1 is the reference's stream).
 * history = taps_per_filter + 1; relative_rate = rate; the first work() returns 0 (:166-169).
 * `ninput_items` counts from the first history item, like the reference's ninput_items[0];
 * *consumed is what the reference passes to consume_each().  Output is bit identical to the reference
 * built with gr_fir_ccf_generic and within 1e-6 relative of the SSE class (bar: 1e-4). */
typedef struct grcuda_pfb_arb grcuda_pfb_arb;
grcuda_pfb_arb* grcuda_pfb_arb_resampler_ccf_create(float rate, const float* taps, int ntaps, unsigned filter_size,
                                                    int nchan);
void grcuda_pfb_arb_resampler_ccf_destroy(grcuda_pfb_arb* h);
int grcuda_pfb_arb_resampler_ccf_set_rate(grcuda_pfb_arb* h, float rate);
unsigned grcuda_pfb_arb_resampler_ccf_history(grcuda_pfb_arb* h);
double grcuda_pfb_arb_resampler_ccf_relative_rate(grcuda_pfb_arb* h);
int grcuda_pfb_arb_resampler_ccf_taps_per_filter(grcuda_pfb_arb* h);
int grcuda_pfb_arb_resampler_ccf_filter_size(grcuda_pfb_arb* h);
/* print_taps (:142-153) as data: taps of one polyphase filter (derivative != 0: of the derivative bank) */
int grcuda_pfb_arb_resampler_ccf_get_taps(grcuda_pfb_arb* h, int filter, int derivative, float* out, int cap);
int grcuda_pfb_arb_resampler_ccf_work(grcuda_pfb_arb* h, int noutput_items, int ninput_items, const grcuda_complex* in,
                                      grcuda_complex* out, int* consumed);
int grcuda_pfb_arb_resampler_ccf_work_device(grcuda_pfb_arb* h, int noutput_items, int ninput_items,
                                             const grcuda_complex* d_in, grcuda_complex* d_out, int* consumed,
                                             void* stream);

/* ---- 8f rank 3  gr_pfb_decimator_ccf -----------------------------------------------------------
 * replaces gr_pfb_decimator_ccf::work, its decim gr_fir_ccf objects and the decim-point gri_fft it uses to
 * de-spin ONE channel (gr_pfb_decimator_ccf.cc:44-175).  Constructor arguments of
 * gr_make_pfb_decimator_ccf(decim, taps, channel).  history = taps_per_filter (:107); one output per input
 * item of every stream; set_taps is deferred and the first work() after it returns 0 (:136-139).
 * _work takes the reference's `decim` stream pointers, _work_interleaved / _work_device the same data as rows
 * [history-1 + noutput][decim] (row m = item m of every stream = decim consecutive samples of the wideband stream). */
typedef struct grcuda_pfb_decim grcuda_pfb_decim;
grcuda_pfb_decim* grcuda_pfb_decimator_ccf_create(unsigned decim, const float* taps, int ntaps, unsigned channel);
void grcuda_pfb_decimator_ccf_destroy(grcuda_pfb_decim* h);
int grcuda_pfb_decimator_ccf_set_taps(grcuda_pfb_decim* h, const float* taps, int ntaps);
unsigned grcuda_pfb_decimator_ccf_history(grcuda_pfb_decim* h);
int grcuda_pfb_decimator_ccf_taps_per_filter(grcuda_pfb_decim* h);
int grcuda_pfb_decimator_ccf_decimation(grcuda_pfb_decim* h);
int grcuda_pfb_decimator_ccf_work(grcuda_pfb_decim* h, int noutput_items, const grcuda_complex* const* in, grcuda_complex* out);
int grcuda_pfb_decimator_ccf_work_interleaved(grcuda_pfb_decim* h, int noutput_items, const grcuda_complex* in_rows,
                                              grcuda_complex* out);
int grcuda_pfb_decimator_ccf_work_device(grcuda_pfb_decim* h, long noutput_items, const grcuda_complex* d_in_rows,
                                         grcuda_complex* d_out, void* stream);

/* ---- 8f rank 4  gr_fft_filter_ccc ---------------------------------------------------------------
 * replaces gr_fft_filter_ccc::work and gri_fft_filter_ccc_generic (gr_fft_filter_ccc.cc:46-106,
 * gri_fft_filter_ccc_generic.cc:62-165): y[n] = sum_k taps[k] x[n-k], decimated, complex taps.  Same
 * block contract: history 1 (the plan carries the last ntaps-1 samples, as the reference carries its
 * overlap-add tail), output_multiple = nsamples = fftsize - ntaps + 1, set_taps deferred to the next work()
 * which returns 0 and clears the carried state.  Device paths (path()): 0 = direct form (ntaps <= 32),
 * 1 = overlap-save with both FFTs and the product inside one CTA (fftsize <= 8192, i.e. up to 4096 taps),
 * 2 = the same kernel once per 4096-tap partition of a longer filter, accumulating (any length, cost linear in
 * ntaps).  Within 5e-6 of the float64 convolution (bar 1e-4).
 * set_path pins one (-1 = automatic); like set_taps it takes effect at the next work(), which returns 0. */
typedef struct grcuda_fft_filter grcuda_fft_filter;
grcuda_fft_filter* grcuda_fft_filter_ccc_create(int decimation, const grcuda_complex* taps, int ntaps);
void grcuda_fft_filter_ccc_destroy(grcuda_fft_filter* h);
int grcuda_fft_filter_ccc_set_taps(grcuda_fft_filter* h, const grcuda_complex* taps, int ntaps);
int grcuda_fft_filter_ccc_output_multiple(grcuda_fft_filter* h);
int grcuda_fft_filter_ccc_path(grcuda_fft_filter* h);
int grcuda_fft_filter_ccc_set_path(grcuda_fft_filter* h, int path);
int grcuda_fft_filter_ccc_decimation(grcuda_fft_filter* h);
unsigned grcuda_fft_filter_ccc_history(grcuda_fft_filter* h);
int grcuda_fft_filter_ccc_work(grcuda_fft_filter* h, int noutput_items, const grcuda_complex* in, grcuda_complex* out);
int grcuda_fft_filter_ccc_work_device(grcuda_fft_filter* h, int noutput_items, const grcuda_complex* d_in,
                                      grcuda_complex* d_out, void* stream);

/* ---- a12 as blocks of their own: gr_map_bb, gr_unpack_k_bits_bb ---------------------------------
 * replaces gr_map_bb::work (gr_map_bb.cc:35-61: identity table overwritten by the first min(256, nmap) entries) and
 * gr_unpack_k_bits_bb::work (gr_unpack_k_bits_bb.cc:38-70: k output bytes per input byte, bit k-1 first;
 * k == 0 -> GRCUDA_ERANGE like the reference's std::out_of_range).  The chain runs both fused into the tail
 * kernel; these are for flowgraphs that keep the blocks between GPU blocks. */
typedef struct grcuda_map_bb grcuda_map_bb;
grcuda_map_bb* grcuda_map_bb_create(const int* map, int nmap);
void grcuda_map_bb_destroy(grcuda_map_bb* h);
int grcuda_map_bb_work(grcuda_map_bb* h, int noutput_items, const unsigned char* in, unsigned char* out);
int grcuda_map_bb_work_device(grcuda_map_bb* h, long noutput_items, const unsigned char* d_in, unsigned char* d_out, void* stream);
typedef struct grcuda_unpack_k_bits grcuda_unpack_k_bits;
grcuda_unpack_k_bits* grcuda_unpack_k_bits_bb_create(unsigned k);
void grcuda_unpack_k_bits_bb_destroy(grcuda_unpack_k_bits* h);
unsigned grcuda_unpack_k_bits_bb_interpolation(grcuda_unpack_k_bits* h);
/* noutput_items counts OUTPUT bytes (a sync interpolator: noutput_items / k input bytes are read) */
int grcuda_unpack_k_bits_bb_work(grcuda_unpack_k_bits* h, int noutput_items, const unsigned char* in, unsigned char* out);
int grcuda_unpack_k_bits_bb_work_device(grcuda_unpack_k_bits* h, long noutput_items, const unsigned char* d_in,
                                        unsigned char* d_out, void* stream);

/* ---- a14 as blocks of their own: gr_stream_to_streams, gr_vector_to_streams ----------------------
 * replaces gr_stream_to_streams::work / gr_vector_to_streams::work (gr_stream_to_streams.cc:52-66,
 * gr_vector_to_streams.cc:53-70 -- the same loop: item i of output stream j = input item i * nstreams + j).
 * The channelizer takes the interleaved stream as it is (_work_interleaved); these are for flowgraphs that split it
 * for other consumers.  _work: `out` = nstreams host pointers, noutput_items items each.  _work_device: one device
 * buffer, stream j at d_out + j * out_stride_items items. */
typedef struct grcuda_streams grcuda_streams;
grcuda_streams* grcuda_stream_to_streams_create(size_t item_size, size_t nstreams);
grcuda_streams* grcuda_vector_to_streams_create(size_t item_size, size_t nstreams);
void grcuda_streams_destroy(grcuda_streams* h);
int grcuda_streams_nstreams(grcuda_streams* h);
int grcuda_streams_work(grcuda_streams* h, int noutput_items, const void* in, void* const* out);
int grcuda_streams_work_device(grcuda_streams* h, long noutput_items, const void* d_in, void* d_out, long out_stride_items,
                               void* stream);

/* ---- 8f rank 4  gr_framer_sink_1 -----------------------------------------------------------------
 * replaces gr_framer_sink_1::work and its state machine (gr_framer_sink_1.cc:36-72, 90-196; header checks
 * gr_framer_sink_1.h:88-103): consumes the correlator's byte stream (bit 0 data, bit 1 sync flag), reads the 32-bit
 * header after a flag (two identical 16-bit halves: whitener offset << 12 | payload length), assembles the payload
 * MSB first and posts one message per packet.  Batched over nchan independent streams, one state machine per
 * channel, state carried from call to call like the block's members.  The reference's gr_msg_queue is a device-side
 * queue (records + payload arena) drained by _read: messages come back ordered by the stream position that
 * completed them, then by channel -- the arrival order of one shared queue.  gr_message(type 0, arg1 =
 * whitener_offset, arg2 = 0, length) is the record's {whitener_offset, length}. */
typedef struct grcuda_framer grcuda_framer;
typedef struct {
  int channel;
  int whitener_offset;      /* gr_message arg1 */
  int length;               /* payload bytes */
  int seq;                  /* n-th message of its channel */
  long long payload_offset; /* into the payload buffer _read fills; -1: the device arena was full, bytes lost */
  long long end_index;      /* position in the channel's stream of the byte that completed the packet */
} grcuda_framer_msg;
grcuda_framer* grcuda_framer_sink_1_create(int nchan, int max_msgs, size_t payload_capacity);
void grcuda_framer_sink_1_destroy(grcuda_framer* h);
/* single stream (nchan == 1), host pointer; returns noutput_items (a sync block consumes all it is given) */
int grcuda_framer_sink_1_work(grcuda_framer* h, int noutput_items, const unsigned char* in);
/* batched device form: item t of channel c = d_in[t * item_stride + c * chan_stride].  The chain's correlator bytes
 * are [bit][channel] (item_stride = nchan, chan_stride = 1); item_stride == 1 selects the warp-per-channel kernel
 * for stream-major data.  Channel c has d_counts[c] * count_scale valid items (d_counts == NULL: nitems each). */
int grcuda_framer_sink_1_work_device(grcuda_framer* h, long nitems, const unsigned char* d_in, long item_stride, long chan_stride,
                                     const int* d_counts, int count_scale, void* stream);
/* messages waiting (synchronises); *dropped = messages or payloads that did not fit since the last read */
int grcuda_framer_sink_1_count(grcuda_framer* h, int* dropped);
/* drains the queue: returns the number of messages copied to msgs (payloads packed into `payload` in that order) */
int grcuda_framer_sink_1_read(grcuda_framer* h, grcuda_framer_msg* msgs, int max_msgs, unsigned char* payload,
                              size_t payload_cap, int* dropped);

/* ---- 8f rank 4  digital_clock_recovery_mm_cc -------------------------------------------------------
 * replaces digital_clock_recovery_mm_cc::general_work (digital_clock_recovery_mm_cc.cc:117-213; set_omega .h:75-80)
 * and gri_mmse_fir_interpolator_cc::interpolate (gri_mmse_fir_interpolator_cc.cc:61-71), batched over nchan
 * channels laid out [time][channel].  create() fails with GRCUDA_ERANGE for omega <= 0 or negative gains (:65-68).
 * history 3, forecast = ceil(noutput * omega + 8) + 16.  Bit identical to the reference built with the generic-order
 * gr_fir_ccf.  err_out / d_err != NULL is the block with its second output connected: the error is clipped to +-4
 * instead of +-1 (:146 vs :178), which changes the loop. */
typedef struct grcuda_mm_cc grcuda_mm_cc;
grcuda_mm_cc* grcuda_clock_recovery_mm_cc_create(int nchan, float omega, float gain_omega, float mu, float gain_mu,
                                                 float omega_relative_limit);
void grcuda_clock_recovery_mm_cc_destroy(grcuda_mm_cc* h);
int grcuda_clock_recovery_mm_cc_forecast(grcuda_mm_cc* h, int noutput_items);
int grcuda_clock_recovery_mm_cc_get_state(grcuda_mm_cc* h, int chan, float* mu, float* omega);
int grcuda_clock_recovery_mm_cc_set_mu(grcuda_mm_cc* h, float mu);
int grcuda_clock_recovery_mm_cc_set_omega(grcuda_mm_cc* h, float omega);
int grcuda_clock_recovery_mm_cc_set_gain_mu(grcuda_mm_cc* h, float gain_mu);
int grcuda_clock_recovery_mm_cc_set_gain_omega(grcuda_mm_cc* h, float gain_omega);
int grcuda_clock_recovery_mm_cc_counters(grcuda_mm_cc* h, long long* clamped, long long* overflow);
/* single-stream general_work (nchan == 1): returns produced, *consumed = what the reference passes to consume_each */
int grcuda_clock_recovery_mm_cc_work(grcuda_mm_cc* h, int noutput_items, int ninput_items, const grcuda_complex* in,
                                     grcuda_complex* out, float* err_out, int* consumed);
/* batched device form, like grcuda_clock_recovery_mm_ff_work_device: rows [time][channel], row 0 = absolute index
 * abs_row0, every channel continues at its own carried position and stops 24 rows before the end (:124) or at
 * max_out symbols; d_out [max_out][nchan], d_counts[c] = symbols produced (NULL: kept inside the plan). */
int grcuda_clock_recovery_mm_cc_work_device(grcuda_mm_cc* h, long ninput_rows, long abs_row0, const grcuda_complex* d_in,
                                            grcuda_complex* d_out, float* d_err, int max_out, int* d_counts, void* stream);

/* ---- flagship pipeline: wideband -> PFB -> batched 4FSK demod -> sync search ---------------
 * One object that owns the HBM-resident intermediates and per-channel loop state and runs
 *   pfb_channelizer_ccf -> quadrature_demod_cf -> fir_filter_fff(RRC) -> clock_recovery_mm_ff
 *   -> 4-level slicer -> map_bb -> unpack_k_bits_bb(2) -> correlate_access_code_bb
 * over successive contiguous time blocks of the wideband stream. */
typedef struct grcuda_dmr_chain grcuda_dmr_chain;
typedef struct {
  unsigned numchans;        /* M */
  const float* pfb_taps;    /* prototype filter */
  int pfb_ntaps;
  float quad_gain;          /* fs_chan / (2 pi 648) for DMR */
  const float* rrc_taps;
  int rrc_ntaps;
  float omega, gain_omega, mu, gain_mu, omega_relative_limit;
  float slicer_alpha;       /* pager_slicer_fb alpha (0 = fixed thresholds) */
  const int* symbol_map;    /* gr_map_bb table applied to the slicer decision */
  int symbol_map_len;
  const char* access_code;  /* e.g. a 48-bit DMR sync pattern */
  int threshold;
  int order;                /* GRCUDA_ORDER_* */
  int max_rows_per_block;   /* sizing of the intermediates */
  int keep_bytes;           /* 1: also produce the reference-format correlator byte stream */
} grcuda_dmr_chain_params;
grcuda_dmr_chain* grcuda_dmr_chain_create(const grcuda_dmr_chain_params* p);
void grcuda_dmr_chain_destroy(grcuda_dmr_chain* h);
/* rows of history the chain wants in front of each block's new rows (pfb taps_per_filter) */
int grcuda_dmr_chain_history_rows(grcuda_dmr_chain* h);
/* Process nrows new wideband rows (numchans samples each).  d_in addresses the first of
 * history_rows() rows that precede them (zeros at stream start, or the neighbour shard's halo).
 * Per-channel demod state is carried inside the handle from block to block. */
int grcuda_dmr_chain_process_device(grcuda_dmr_chain* h, const grcuda_complex* d_in, int nrows, void* stream);
/* the two halves of process_device, for time shards: front = channelizer + discriminator + matched
 * filter (finite memory), tail = M&M + slicer + correlator (loop state); a shard runs front on its
 * block, import_state()s its left neighbour's state, then runs tail. */
int grcuda_dmr_chain_process_front_device(grcuda_dmr_chain* h, const grcuda_complex* d_in, int nrows, void* stream);
int grcuda_dmr_chain_process_tail_device(grcuda_dmr_chain* h, void* stream);
/* process_device runs the front on `stream` and the tail on the chain's own stream, so that the tail
 * of block b overlaps the front of block b+1 (GRCUDA_CHAIN_NO_OVERLAP=1 in the environment puts both
 * on `stream`).  join() makes `stream` wait for every tail queued so far: call it before timing or
 * before consuming result_get() pointers on that stream.  read_hits() synchronises by itself. */
int grcuda_dmr_chain_join(grcuda_dmr_chain* h, void* stream);
/* same through host memory: pinned double-buffered staging (sub-blocks: H2D of i+1 overlaps compute
 * of i); sync hits of all sub-blocks accumulate for read_hits */
int grcuda_dmr_chain_process_host(grcuda_dmr_chain* h, const grcuda_complex* in, int nrows);
/* results of the last processed block (device pointers valid until the next process call) */
typedef struct {
  const grcuda_complex* d_channels; /* [nrows][M] channelizer output */
  const float* d_soft;              /* [max_sym][M] M&M output */
  const unsigned char* d_symbols;   /* [max_sym][M] slicer decisions */
  const int* d_sym_counts;          /* [M] */
  const unsigned char* d_bytes;     /* [max_sym*2][M] correlator bytes or NULL */
  const grcuda_hit* d_hits;
  const int* d_nhits;
  int max_sym;
  int nrows;
} grcuda_dmr_chain_result;
/* NOTE: process_host cuts a block into sub-blocks that reuse the symbol buffers; after it, d_soft / d_symbols /
 * d_sym_counts / d_channels / nrows describe the LAST sub-block only (hits accumulate over the whole block).  A caller
 * that wants every symbol creates the chain with keep_bytes (one sub-block) or drives process_device itself. */
int grcuda_dmr_chain_result_get(grcuda_dmr_chain* h, grcuda_dmr_chain_result* r);
/* keep_channels = 0: the channelizer output is not materialised (d_channels = NULL); the discriminator then runs inside
 * the last pass of the channelizer's FFT kernel on values that are still in registers, which removes 16 of the chain's
 * 50 B of HBM traffic per input sample.  keep_channels = 2 is the same kernel which ALSO stores the transform it computed
 * (d_channels valid; for parity tests: the tail is bit exact on exactly those values).  1 = default, two kernels.
 * The fused kernel's transform differs from the plain FFT kernel's in the last bit (different multiply-add contraction;
 * both within 1e-6 of the float64 DFT), so modes 0/2 and mode 1 agree like two FFT libraries do: dibits only differ
 * within the slicer epsilon band.  Modes 0 and 2 are bit identical to each other.
 * GRCUDA_EUNSUPPORTED (and nothing changes) when the channel count has no such kernel (it exists for 8000 and 4096
 * channels at oversample rate 1, SSE summation order).  Call it on a fresh chain or right after seek(). */
int grcuda_dmr_chain_set_keep_channels(grcuda_dmr_chain* h, int keep_channels);
int grcuda_dmr_chain_keeps_channels(grcuda_dmr_chain* h);
/* error counters of the chain (see grcuda_clock_recovery_mm_ff_counters) + sync hits that did not fit the hit list */
int grcuda_dmr_chain_counters(grcuda_dmr_chain* h, long long* clamped, long long* overflow, long long* hits_dropped);
/* copy the compacted sync hits of the last block to the host; returns their number */
int grcuda_dmr_chain_read_hits(grcuda_dmr_chain* h, grcuda_hit* hits, int max_hits);
/* smallest block process_* accepts (the carries of one block must not overlap the next) */
int grcuda_dmr_chain_min_rows(grcuda_dmr_chain* h);
/* Time-shard support.  A shard that does not start at stream row 0 creates its own
 * chain, seek()s to (first_row - warmup_rows()), processes warmup_rows() rows of halo (results
 * discarded) so that every FIR history is that of the continuous stream, then import_state()s the
 * loop state its left neighbour exported and continues bit-identically to a single-GPU run. */
int grcuda_dmr_chain_warmup_rows(grcuda_dmr_chain* h);
int grcuda_dmr_chain_seek(grcuda_dmr_chain* h, long long abs_row);
int grcuda_dmr_chain_seek_async(grcuda_dmr_chain* h, long long abs_row, void* stream); /* stream ordered, no sync */
long long grcuda_dmr_chain_tell(grcuda_dmr_chain* h);
/* shard hand-off: export / import the per-channel loop state
 * {mu, omega, last_sample, next input index, slicer avg, correlator registers} so that the next
 * time shard continues exactly where this one stopped.  Buffer = state_bytes() bytes (device). */
size_t grcuda_dmr_chain_state_bytes(grcuda_dmr_chain* h);
int grcuda_dmr_chain_export_state(grcuda_dmr_chain* h, void* d_state, void* stream);
int grcuda_dmr_chain_import_state(grcuda_dmr_chain* h, const void* d_state, void* stream);

/* ---- per-stage device timing (bench.py) -------------------------------------------------------
 * CUDA events recorded on the launch stream around each stage's kernels.  profile_read
 * synchronises, returns the accumulated milliseconds / launch counts since the last read. */
#define GRCUDA_STAGE_PFB_FIR 0
#define GRCUDA_STAGE_PFB_FFT 1
#define GRCUDA_STAGE_QUAD 2
#define GRCUDA_STAGE_RRC 3
#define GRCUDA_STAGE_MM 4
#define GRCUDA_STAGE_CORR 5
#define GRCUDA_STAGE_CARRY 6
#define GRCUDA_NSTAGES 7
int grcuda_pfb_channelizer_ccf_set_profiling(grcuda_pfb* h, int on);
int grcuda_pfb_channelizer_ccf_profile_read(grcuda_pfb* h, float ms[2], int launches[2]);
/* which build of the clock-recovery kernel the tail stage runs (grcuda_clock_recovery_mm_ff_set_kernel_variant) */
/* The tail stage as its two kernels, for a time shard (the clock-recovery kernel is the one serial chain over all blocks
 * of all ranks: nothing else sits on it).  process_tail_mm reads the loop state at d_mm_state_in (NULL: the chain's own)
 * and ALSO writes its final state to d_mm_state_out (NULL: nowhere else) -- the receive and send buffers of the state
 * ring, no copies; process_tail_corr runs the time-parallel correlator behind it (any stream), state likewise.
 * GRCUDA_EUNSUPPORTED when that correlator does not apply (keep_bytes, code shorter than 16): use process_tail_device. */
int grcuda_dmr_chain_process_tail_mm_device(grcuda_dmr_chain* h, const void* d_mm_state_in, void* d_mm_state_out, void* stream);
int grcuda_dmr_chain_process_tail_corr_device(grcuda_dmr_chain* h, const void* d_corr_state_in, void* d_corr_state_out, void* stream);
size_t grcuda_dmr_chain_mm_state_bytes(grcuda_dmr_chain* h);
size_t grcuda_dmr_chain_corr_state_bytes(grcuda_dmr_chain* h);
/* 1: the sync-hit list is NOT cleared at the start of a block: hits of successive blocks accumulate (up to max_hits;
 * bit indices are absolute, so the list still says where each hit is) until grcuda_dmr_chain_clear_hits. */
int grcuda_dmr_chain_set_accumulate_hits(grcuda_dmr_chain* h, int on);
int grcuda_dmr_chain_clear_hits(grcuda_dmr_chain* h, void* stream);
int grcuda_dmr_chain_max_hits(grcuda_dmr_chain* h);
int grcuda_dmr_chain_set_tail_variant(grcuda_dmr_chain* h, int variant);
/* 1 (default): the tail is two kernels, clock recovery + slicer, then the access-code correlator parallel over channels
 * and time; 0: one fused kernel (always used when the correlator's byte stream is kept).  Identical results. */
int grcuda_dmr_chain_set_split_correlator(grcuda_dmr_chain* h, int on);
int grcuda_dmr_chain_set_profiling(grcuda_dmr_chain* h, int on);
int grcuda_dmr_chain_profile_read(grcuda_dmr_chain* h, float ms[GRCUDA_NSTAGES], int launches[GRCUDA_NSTAGES]);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* INCLUDED_GR_CUDA_H */
