// Flagship pipeline object: wideband interleaved stream -> gr_pfb_channelizer_ccf ->
// gr_quadrature_demod_cf -> gr_fir_filter_fff (RRC) -> digital_clock_recovery_mm_ff ->
// pager_slicer_fb -> gr_map_bb -> gr_unpack_k_bits_bb(2) -> digital_correlate_access_code_bb,
// batched over all M channels, with every intermediate resident in HBM and the per-channel loop
// state carried from block to block.
//
// Data layout in HBM (row = one channel-rate time step, M channels wide):
//   Y  [YH + R][M] complex  channelizer output; first YH rows = carried tail of the previous block
//                           (fused front: YH = nrrc + 4 rows of discriminator + matched-filter history;
//                           unfused: YH = 1, the discriminator's x[i-1])
//   D  [nrrc-1 + R][M] f32  discriminator output (UNFUSED path only: generic summation order or very
//                           long matched filters); first nrrc-1 rows = carried history of the RRC FIR
//   F  2 x [KEEP + R][M] f32  matched-filter output, double buffered; first KEEP rows = carry for the
//                           M&M interpolator (copied from the other buffer's tail)
//
// Streams.  The front (channelizer, discriminator, matched filter: HBM / FP32 bound, fills the
// machine) runs on the caller's stream; the tail (M&M recursion + slicer + correlator: 125 CTAs
// running at instruction latency, see kernel_mm.cuh) runs on the chain's own tail stream and
// therefore overlaps the front of the NEXT block.  Events order the two: tail(b) waits for
// front(b); front(b+2) waits for tail(b) before it reuses that F buffer.
//   soft/sym [max_sym][M], bytes [2*max_sym][M], hits[], per-channel state arrays
#include <algorithm>
#include <cmath>

#include "common.cuh"
#include "internal.h"

using namespace grb;

namespace {
const int KEEP = 64;  // rows of F kept in front of each block (M&M look-back, see mm_kernel)
}

struct grcuda_dmr_chain {
  unsigned M = 0;
  int T = 0, nrrc = 0, max_rows = 0, max_sym = 0, max_hits = 0, keep_bytes = 0;
  int YH = 1;          // history rows in front of Y
  bool fused = false;  // discriminator + matched filter in one kernel (kernel_demod_front.cuh)
  bool fft_demod_store_y = false;  // ... and the kernel also stores the transform it computed (parity tests)
  bool fft_demod = false;  // discriminator inside the channelizer's FFT kernel (kernel_fft_demod.cuh): Y never reaches HBM
  DevBuf Dd, yl[2];        // that mode's D rows [YH + R][M] (first YH rows carried) and the last Y row of the previous / this block
  int yl_cur = 0;
  grcuda_pfb* pfb = nullptr;
  grcuda_quad* quad = nullptr;
  grcuda_fir_fff* rrc = nullptr;
  grcuda_mm* mm = nullptr;
  grcuda_corr* corr = nullptr;
  std::vector<int> symbol_map;
  DevBuf Yb, D, Fb[2], soft, sym, counts, bytes, hits, nhits, d_in_host;
  int fcur = 0;           // F buffer of the block whose front ran last
  bool pipeline = true;   // tail on its own stream (overlaps the next block's front)
  bool split_corr = true; // correlator as its own time-parallel kernel behind the clock-recovery kernel
  int tail_variant = -1;  // user's choice of clock-recovery kernel; -1: by mode (see process_tail_device)
  bool split_user = false, under_front = false;
  cudaStream_t tail_stream = nullptr;
  cudaEvent_t ev_front[2] = {nullptr, nullptr}, ev_tail[2] = {nullptr, nullptr};
  bool tail_pending[2] = {false, false};
  cudaEvent_t ev_mm = nullptr, ev_corr = nullptr;  // two-kernel tail of a time shard: clock recovery done / correlator done
  bool corr_pending = false, mm_pending = false;
  cudaEvent_t ev_state = nullptr;  // export/import_state copies (they touch what the tail kernel reads and writes)
  bool state_pending = false;
  int prev_rows = 0;      // rows of the previous block (where its F tail sits)
  cudaStream_t stream = nullptr;
  Stager stager;
  long long abs_row = 0;  // absolute channel-rate row index of the next new row
  EventProfiler prof;     // stages: 2 quad, 3 rrc, 4 mm, 5 corr, 6 carries (0/1 live in the pfb plan)
  int last_rows = 0, front_rows = 0;
  bool accumulate_hits = false;
  cudaStream_t last_stream = nullptr, copy_stream = nullptr;
  cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
  DevBuf d_stage[2];
  ~grcuda_dmr_chain() {
    if (pfb) grcuda_pfb_channelizer_ccf_destroy(pfb);
    if (quad) grcuda_quadrature_demod_cf_destroy(quad);
    if (rrc) grcuda_fir_filter_fff_destroy(rrc);
    if (mm) grcuda_clock_recovery_mm_ff_destroy(mm);
    if (corr) grcuda_correlate_access_code_bb_destroy(corr);
    if (stream) cudaStreamDestroy(stream);
    if (tail_stream) cudaStreamDestroy(tail_stream);
    if (copy_stream) cudaStreamDestroy(copy_stream);
    for (int i = 0; i < 2; i++) {
      if (ev_front[i]) cudaEventDestroy(ev_front[i]);
      if (ev_tail[i]) cudaEventDestroy(ev_tail[i]);
    }
    if (ev_state) cudaEventDestroy(ev_state);
    if (ev_mm) cudaEventDestroy(ev_mm);
    if (ev_corr) cudaEventDestroy(ev_corr);
    for (int i = 0; i < 2; i++) {
      if (ev_copied[i]) cudaEventDestroy(ev_copied[i]);
      if (ev_done[i]) cudaEventDestroy(ev_done[i]);
    }
  }
};

// The tail kernel (125 small CTAs that then run for ~1 ms at instruction latency) must get onto the SMs BEFORE
// the next block's front kernels fill them: placed late it starts late, and the SMs' register files are
// fragmented so that the big-tile kernels no longer fit next to it.  Highest stream priority makes the block
// scheduler take its CTAs first whenever both have CTAs pending.
static cudaError_t create_tail_stream(cudaStream_t* s) {
  int lo = 0, hi = 0;
  cudaError_t e = cudaDeviceGetStreamPriorityRange(&lo, &hi);
  if (e != cudaSuccess) return e;
  return cudaStreamCreateWithPriority(s, cudaStreamNonBlocking, hi);
}

extern "C" {

grcuda_dmr_chain* grcuda_dmr_chain_create(const grcuda_dmr_chain_params* p) {
  if (!p || p->numchans < 1 || p->max_rows_per_block < 1 || p->rrc_ntaps < 1) {
    set_error(GRCUDA_EINVAL, "dmr_chain: bad parameters");
    return nullptr;
  }
  grcuda_dmr_chain* h = new grcuda_dmr_chain;
  h->M = p->numchans;
  h->max_rows = p->max_rows_per_block;
  h->nrrc = p->rrc_ntaps;
  h->keep_bytes = p->keep_bytes;
  h->pfb = grcuda_pfb_channelizer_ccf_create(p->numchans, p->pfb_taps, p->pfb_ntaps, 1.0f);
  h->quad = h->pfb ? grcuda_quadrature_demod_cf_create(p->quad_gain) : nullptr;
  h->rrc = h->quad ? grcuda_fir_filter_fff_create(1, p->rrc_taps, p->rrc_ntaps, p->order) : nullptr;
  h->mm = h->rrc ? grcuda_clock_recovery_mm_ff_create((int)p->numchans, p->omega, p->gain_omega, p->mu, p->gain_mu,
                                                      p->omega_relative_limit, p->order)
                 : nullptr;
  h->corr = h->mm ? grcuda_correlate_access_code_bb_create((int)p->numchans, p->access_code, p->threshold) : nullptr;
  if (!h->corr) { delete h; return nullptr; }
  grcuda_clock_recovery_mm_ff_set_slicer(h->mm, 4, p->slicer_alpha);
  h->symbol_map.assign(p->symbol_map, p->symbol_map + p->symbol_map_len);
  h->T = grcuda_pfb_channelizer_ccf_taps_per_filter(h->pfb);
  // The fused discriminator + matched-filter kernel restates the SSE summation order only
  h->fused = p->order == GRCUDA_ORDER_SSE && h->nrrc <= demod_front_max_taps();
  h->YH = h->fused ? demod_front_history(h->nrrc) : 1;
  const size_t M = h->M, R = h->max_rows;
  // symbol capacity per block: the loop advances omega rows per symbol on average and omega never leaves
  // omega_mid +- omega_relative_limit (an ABSOLUTE band: digital_clock_recovery_mm_ff.cc:124), so a block yields at
  // most rows / (omega - limit) symbols (+ the interpolator's jitter); a configuration the buffer cannot cover is refused
  if (!(p->omega - p->omega_relative_limit >= 1.0f)) {
    set_error(GRCUDA_EINVAL, "dmr_chain: omega - omega_relative_limit = %g < 1 sample per symbol", (double)(p->omega - p->omega_relative_limit));
    delete h;
    return nullptr;
  }
  h->max_sym = (int)std::ceil((double)(KEEP + R) / (double)(p->omega - p->omega_relative_limit) * 1.25) + 64;  // x 1.25: gain_mu * mean(mm) also moves the pace
  // room for the sync hits of many blocks when the caller lets them accumulate (bench.py: a whole timed region per rank)
  h->max_hits = std::max<int>(4096, (int)std::min<size_t>(M * (size_t)h->max_sym / 8, (size_t)1 << 26));
  int rc = 0;
  rc = rc ? rc : pfb_reserve_rows(h->pfb, (long)R);  // no allocation (= device-wide sync) once blocks are flowing
  rc = rc ? rc : h->Yb.reserve((h->YH + R) * M * sizeof(float2));
  if (!h->fused) rc = rc ? rc : h->D.reserve((h->nrrc - 1 + R) * M * sizeof(float));
  rc = rc ? rc : h->Fb[0].reserve((KEEP + R) * M * sizeof(float));
  rc = rc ? rc : h->Fb[1].reserve((KEEP + R) * M * sizeof(float));
  rc = rc ? rc : h->soft.reserve((size_t)h->max_sym * M * sizeof(float));
  rc = rc ? rc : h->sym.reserve((size_t)h->max_sym * M);
  rc = rc ? rc : h->counts.reserve(M * sizeof(int));
  rc = rc ? rc : h->hits.reserve((size_t)h->max_hits * sizeof(grcuda_hit));
  rc = rc ? rc : h->nhits.reserve(sizeof(int));
  if (!rc && h->keep_bytes) rc = h->bytes.reserve((size_t)2 * h->max_sym * M);
  if (rc) { delete h; return nullptr; }
  // stream start: all histories are the zeros the reference runtime pre-loads (gr_buffer.cc:201-214)
  if (cudaMemset(h->Yb.p, 0, (size_t)h->YH * M * sizeof(float2)) != cudaSuccess ||
      (!h->fused && cudaMemset(h->D.p, 0, (size_t)(h->nrrc - 1) * M * sizeof(float) + 4) != cudaSuccess) ||
      cudaMemset(h->Fb[0].p, 0, (size_t)KEEP * M * sizeof(float)) != cudaSuccess ||
      cudaMemset(h->Fb[1].p, 0, (size_t)KEEP * M * sizeof(float)) != cudaSuccess ||
      cudaMemset(h->nhits.p, 0, sizeof(int)) != cudaSuccess ||
      cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess ||
      create_tail_stream(&h->tail_stream) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_front[0], cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_front[1], cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_tail[0], cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_tail[1], cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_state, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_mm, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_corr, cudaEventDisableTiming) != cudaSuccess) {
    set_error(GRCUDA_ECUDA, "dmr_chain: device initialisation failed: %s", cudaGetErrorString(cudaGetLastError()));
    delete h;
    return nullptr;
  }
  h->pipeline = getenv("GRCUDA_CHAIN_NO_OVERLAP") == nullptr;
  if (h->pipeline && pfb_prefer_coresident_fft(h->pfb) != GRCUDA_OK) { delete h; return nullptr; }
  // the initialisation above went through the legacy stream; every stream the work runs on is non-blocking
  if (cudaDeviceSynchronize() != cudaSuccess) {
    set_error(GRCUDA_ECUDA, "dmr_chain: device initialisation failed: %s", cudaGetErrorString(cudaGetLastError()));
    delete h;
    return nullptr;
  }
  return h;
}

void grcuda_dmr_chain_destroy(grcuda_dmr_chain* h) { delete h; }
// 0: the channelizer output is not kept (result_get's d_channels is NULL): the discriminator runs inside the
// channelizer's FFT kernel, 16 B per sample less HBM traffic.  Same symbols, same sync hits, bit for bit.
int grcuda_dmr_chain_set_keep_channels(grcuda_dmr_chain* h, int on) {
  if (h->front_rows > 0) return set_error(GRCUDA_EINVAL, "dmr_chain: set_keep_channels between process_front and process_tail");
  if (on == 1) { h->fft_demod = false; h->fft_demod_store_y = false; return GRCUDA_OK; }
  if (!h->fused || !pfb_demod_supported(h->pfb))
    return set_error(GRCUDA_EUNSUPPORTED, "dmr_chain: no fused FFT + discriminator kernel for this channel count / summation order");
  const size_t M = h->M;
  GRB_CUDA(cudaDeviceSynchronize());
  int rc;
  if ((rc = h->Dd.reserve(((size_t)h->YH + h->max_rows) * M * sizeof(float))) || (rc = h->yl[0].reserve(M * sizeof(float2))) ||
      (rc = h->yl[1].reserve(M * sizeof(float2))))
    return rc;
  if (!h->fft_demod) {
    // continue the stream where the two-kernel path left it: D history = discriminator of the carried Y rows
    // (only a fresh chain / after seek() is exact without this: the carried Y rows are zero then)
    GRB_CUDA(cudaMemset(h->Dd.p, 0, (size_t)h->YH * M * sizeof(float)));
    GRB_CUDA(cudaMemset(h->yl[0].p, 0, M * sizeof(float2)));
    GRB_CUDA(cudaMemset(h->yl[1].p, 0, M * sizeof(float2)));
    GRB_CUDA(cudaDeviceSynchronize());
  }
  h->fft_demod = true;
  h->fft_demod_store_y = on == 2;
  return GRCUDA_OK;
}
int grcuda_dmr_chain_keeps_channels(grcuda_dmr_chain* h) { return h->fft_demod ? (h->fft_demod_store_y ? 2 : 0) : 1; }
int grcuda_dmr_chain_history_rows(grcuda_dmr_chain* h) { return h->T; }

int grcuda_dmr_chain_min_rows(grcuda_dmr_chain* h) { return std::max(std::max(KEEP, h->nrrc - 1), h->YH); }
int grcuda_dmr_chain_warmup_rows(grcuda_dmr_chain* h) {
  // rows a fresh chain must process before its F rows are those of the continuous stream:
  // PFB transient (T) + discriminator / RRC history (YH + nrrc) + M&M look-back (KEEP) + look-ahead (8)
  const int w = h->T + h->YH + h->nrrc + KEEP + 8;
  return std::max(w, grcuda_dmr_chain_min_rows(h));
}
int grcuda_dmr_chain_seek(grcuda_dmr_chain* h, long long abs_row) {
  const size_t M = h->M;
  GRB_CUDA(cudaDeviceSynchronize());
  GRB_CUDA(cudaMemset(h->Yb.p, 0, (size_t)h->YH * M * sizeof(float2)));
  if (!h->fused) GRB_CUDA(cudaMemset(h->D.p, 0, (size_t)(h->nrrc - 1) * M * sizeof(float) + 4));
  if (h->Dd.p) {
    GRB_CUDA(cudaMemset(h->Dd.p, 0, (size_t)h->YH * M * sizeof(float)));
    GRB_CUDA(cudaMemset(h->yl[0].p, 0, M * sizeof(float2)));
    GRB_CUDA(cudaMemset(h->yl[1].p, 0, M * sizeof(float2)));
  }
  h->prev_rows = 0;  // the next block starts from a zero F carry
  h->abs_row = abs_row;
  return GRCUDA_OK;
}
// stream-ordered variant for the steady state of a time shard (no device-wide synchronisation)
int grcuda_dmr_chain_seek_async(grcuda_dmr_chain* h, long long abs_row, void* stream_) {
  cudaStream_t s = stream_ ? (cudaStream_t)stream_ : h->stream;
  const size_t M = h->M;
  GRB_CUDA(cudaMemsetAsync(h->Yb.p, 0, (size_t)h->YH * M * sizeof(float2), s));
  if (!h->fused) GRB_CUDA(cudaMemsetAsync(h->D.p, 0, (size_t)(h->nrrc - 1) * M * sizeof(float) + 4, s));
  if (h->Dd.p) {
    GRB_CUDA(cudaMemsetAsync(h->Dd.p, 0, (size_t)h->YH * M * sizeof(float), s));
    GRB_CUDA(cudaMemsetAsync(h->yl[0].p, 0, M * sizeof(float2), s));
    GRB_CUDA(cudaMemsetAsync(h->yl[1].p, 0, M * sizeof(float2), s));
  }
  h->prev_rows = 0;  // the next block starts from a zero F carry
  h->abs_row = abs_row;
  return GRCUDA_OK;
}
long long grcuda_dmr_chain_tell(grcuda_dmr_chain* h) { return h->abs_row; }

size_t grcuda_dmr_chain_state_bytes(grcuda_dmr_chain* h) { return mm_state_bytes(h->mm) + corr_state_bytes(h->corr); }
// the loop state is read and written by the tail kernel, which may run on the chain's tail stream:
// order the copy after the last tail, and the next tail after the copy
static int state_copy_begin(grcuda_dmr_chain* h, cudaStream_t s) {
  if (h->tail_pending[h->fcur]) GRB_CUDA(cudaStreamWaitEvent(s, h->ev_tail[h->fcur], 0));
  return GRCUDA_OK;
}
static int state_copy_end(grcuda_dmr_chain* h, cudaStream_t s) {
  GRB_CUDA(cudaEventRecord(h->ev_state, s));
  h->state_pending = true;
  return GRCUDA_OK;
}
int grcuda_dmr_chain_join(grcuda_dmr_chain* h, void* stream_) {
  cudaStream_t s = stream_ ? (cudaStream_t)stream_ : h->stream;
  for (int i = 0; i < 2; i++)
    if (h->tail_pending[i]) GRB_CUDA(cudaStreamWaitEvent(s, h->ev_tail[i], 0));
  if (h->corr_pending) GRB_CUDA(cudaStreamWaitEvent(s, h->ev_corr, 0));
  return GRCUDA_OK;
}
int grcuda_dmr_chain_export_state(grcuda_dmr_chain* h, void* d_state, void* stream_) {
  cudaStream_t s = stream_ ? (cudaStream_t)stream_ : h->stream;
  int rc0 = state_copy_begin(h, s);
  if (rc0) return rc0;
  GRB_CUDA(cudaMemcpyAsync(d_state, mm_state_ptr(h->mm), mm_state_bytes(h->mm), cudaMemcpyDeviceToDevice, s));
  GRB_CUDA(cudaMemcpyAsync((char*)d_state + mm_state_bytes(h->mm), corr_state_ptr(h->corr), corr_state_bytes(h->corr),
                           cudaMemcpyDeviceToDevice, s));
  return state_copy_end(h, s);
}
int grcuda_dmr_chain_import_state(grcuda_dmr_chain* h, const void* d_state, void* stream_) {
  cudaStream_t s = stream_ ? (cudaStream_t)stream_ : h->stream;
  int rc0 = state_copy_begin(h, s);
  if (rc0) return rc0;
  GRB_CUDA(cudaMemcpyAsync(mm_state_ptr(h->mm), d_state, mm_state_bytes(h->mm), cudaMemcpyDeviceToDevice, s));
  GRB_CUDA(cudaMemcpyAsync(corr_state_ptr(h->corr), (const char*)d_state + mm_state_bytes(h->mm), corr_state_bytes(h->corr),
                           cudaMemcpyDeviceToDevice, s));
  return state_copy_end(h, s);
}

// front stage: channelizer -> discriminator -> matched filter.  Finite-memory stages: a time shard can run this on
// its block + halo without waiting for anybody.  One stream, one Y buffer, history carried in place.
static int front_impl(grcuda_dmr_chain* h, const grcuda_complex* d_in, int nrows, cudaStream_t sA, bool coresident = false) {
  if (nrows < grcuda_dmr_chain_min_rows(h) || nrows > h->max_rows)
    return set_error(GRCUDA_EINVAL, "dmr_chain: nrows %d outside [%d, %d]", nrows, grcuda_dmr_chain_min_rows(h), h->max_rows);
  if (h->front_rows > 0) return set_error(GRCUDA_EINVAL, "dmr_chain: process_front twice without process_tail");
  cudaStream_t sB = sA;
  const size_t M = h->M;
  const long R = nrows;
  const size_t YH = h->YH;
  int rc;
  const int cur = h->fcur ^ 1;  // this block's F buffer; the other one holds the previous block
  float2* Y = h->Yb.as<float2>();
  float* D = h->D.as<float>();
  float* F = h->Fb[cur].as<float>();
  const float* Fprev = h->Fb[cur ^ 1].as<float>();
  // F[cur] was last read by the tail of the block before the previous one.  (Before the channelizer: the small
  // copy below then gives the tail kernel of the previous block, which becomes runnable at the same moment, the
  // few microseconds it needs to get all its CTAs onto empty SMs; measured: with the channelizer launched first
  // the tail kernel runs in two waves, 1.05 -> 2.15 ms.)
  if (h->tail_pending[cur]) {
    GRB_CUDA(cudaStreamWaitEvent(sB, h->ev_tail[cur], 0));
    h->tail_pending[cur] = false;
  }
  // M&M look-back carry: the last KEEP matched-filter rows of the previous block (written by its front on this
  // same stream order; its tail only reads them)
  h->prof.begin(6, sB);
  if (h->prev_rows > 0)
    GRB_CUDA(cudaMemcpyAsync(F, Fprev + (size_t)h->prev_rows * M, (size_t)KEEP * M * sizeof(float), cudaMemcpyDeviceToDevice, sB));
  else
    GRB_CUDA(cudaMemsetAsync(F, 0, (size_t)KEEP * M * sizeof(float), sB));
  h->prof.end(sB, 0);
  if (h->fft_demod) {
    // 1+2. channelizer with the discriminator in the last FFT pass: rows -> D rows YH..YH+R (4 B/sample); the
    //      channelizer output stays in registers.  3. matched filter from D.
    float* Dd = h->Dd.as<float>();
    int nt = 0;
    fir_fff_reversed_taps(h->rrc, &nt, nullptr);
    if ((rc = pfb_work_device_demod(h->pfb, R, (const float2*)d_in, Dd + YH * M, quad_gain(h->quad), h->yl[h->yl_cur].as<float2>(),
                                    h->yl[h->yl_cur ^ 1].as<float2>(), coresident, sA, h->fft_demod_store_y ? Y + YH * M : nullptr)))
      return rc;
    h->yl_cur ^= 1;
    h->prof.begin(3, sB);
    if ((rc = demod_front_launch(nullptr, F + (size_t)KEEP * M, (long)h->abs_row, (int)R, (int)M, quad_gain(h->quad),
                                 fir_fff_front_taps(h->rrc), nt, sB, Dd, fir_fff_front_taps_host(h->rrc))))
      return rc;
    h->prof.end(sB);
    h->prof.begin(6, sB);
    GRB_CUDA(cudaMemcpyAsync(Dd, Dd + (size_t)R * M, YH * M * sizeof(float), cudaMemcpyDeviceToDevice, sB));
    h->prof.end(sB, 0);
    GRB_CUDA(cudaEventRecord(h->ev_front[cur], sB));
    h->fcur = cur;
    h->front_rows = nrows;
    return GRCUDA_OK;
  }
  // 1. channelizer: [T + R][M] -> Y rows YH..YH+R
  if ((rc = grcuda_pfb_channelizer_ccf_work_device(h->pfb, R, d_in, (grcuda_complex*)(Y + YH * M), sA))) return rc;
  if (h->fused) {
    // 2+3. discriminator + matched filter in one pass: Y (8 B/sample) -> F (4 B/sample); the
    //      discriminator output only ever lives in shared memory
    int nt = 0;
    fir_fff_reversed_taps(h->rrc, &nt, nullptr);
    h->prof.begin(3, sB);
    if ((rc = demod_front_launch(Y, F + (size_t)KEEP * M, (long)h->abs_row, (int)R, (int)M, quad_gain(h->quad),
                                 fir_fff_front_taps(h->rrc), nt, sB, nullptr, fir_fff_front_taps_host(h->rrc))))
      return rc;
    h->prof.end(sB);
  } else {
    // 2. discriminator: Y rows YH-1.. -> D rows (nrrc-1)..
    h->prof.begin(2, sB);
    if ((rc = grcuda_quadrature_demod_cf_work_device(h->quad, R, (int)M, (const grcuda_complex*)(Y + (YH - 1) * M),
                                                     D + (size_t)(h->nrrc - 1) * M, sB)))
      return rc;
    h->prof.end(sB);
    // 3. matched filter: D (history-prefixed; row 0 is absolute row abs_row-(nrrc-1)) -> F rows KEEP..
    h->prof.begin(3, sB);
    if ((rc = grcuda_fir_filter_fff_work_device(h->rrc, R, (int)M, D, F + (size_t)KEEP * M, (long)(h->abs_row - (h->nrrc - 1)), sB))) return rc;
    h->prof.end(sB);
  }
  // carries of the front stages for the next block (small device-to-device copies, stream ordered;
  // nrows >= min_rows guarantees that source and destination never overlap)
  h->prof.begin(6, sB);
  GRB_CUDA(cudaMemcpyAsync(Y, Y + (size_t)R * M, YH * M * sizeof(float2), cudaMemcpyDeviceToDevice, sB));
  if (!h->fused && h->nrrc > 1)
    GRB_CUDA(cudaMemcpyAsync(D, D + (size_t)R * M, (size_t)(h->nrrc - 1) * M * sizeof(float), cudaMemcpyDeviceToDevice, sB));
  h->prof.end(sB, 0);
  GRB_CUDA(cudaEventRecord(h->ev_front[cur], sB));
  h->fcur = cur;
  h->front_rows = nrows;
  return GRCUDA_OK;
}

int grcuda_dmr_chain_process_front_device(grcuda_dmr_chain* h, const grcuda_complex* d_in, int nrows, void* stream_) {
  cudaStream_t s = stream_ ? (cudaStream_t)stream_ : h->stream;
  return front_impl(h, d_in, nrows, s);
}

// tail stage: the loops with infinite memory (M&M + DC-tracking slicer + correlator registers); a
// time shard runs it after importing its left neighbour's loop state
int grcuda_dmr_chain_process_tail_device(grcuda_dmr_chain* h, void* stream_) {
  if (h->front_rows <= 0) return set_error(GRCUDA_EINVAL, "dmr_chain: process_tail without a pending process_front");
  cudaStream_t s = stream_ ? (cudaStream_t)stream_ : h->stream;
  h->last_stream = s;
  const long R = h->front_rows;
  const int cur = h->fcur;
  int rc;
  float* F = h->Fb[cur].as<float>();
  GRB_CUDA(cudaStreamWaitEvent(s, h->ev_front[cur], 0));
  if (h->state_pending) {
    GRB_CUDA(cudaStreamWaitEvent(s, h->ev_state, 0));
    h->state_pending = false;
  }
  // 4+5. clock recovery + slicer + (dibit map -> bits -> sync correlation fused in the same kernel)
  //      over F rows [abs_row-KEEP, abs_row+R)
  if (!h->accumulate_hits) GRB_CUDA(cudaMemsetAsync(h->nhits.p, 0, sizeof(int), s));
  // Which kernels: a tail that runs UNDERNEATH the next block's front (single-GPU pipeline) is the 48-register /
  // 47 KB build with the correlator fused, which co-resides with the front kernels (measured: 1.85 ms per block
  // against 2.05 ms with the stand-alone build, whose 198 KB of shared memory keep the front off its 125 SMs); a tail
  // that has the device to itself is the quad-ring kernel followed by the time-parallel correlator (0.59 ms against
  // 0.74 ms).  Explicit choices (set_tail_variant / set_split_correlator) win.
  const bool split = h->split_user ? h->split_corr : !h->under_front;
  if (h->tail_variant < 0) grcuda_clock_recovery_mm_ff_set_kernel_variant(h->mm, h->under_front ? 10 : -1);
  h->prof.begin(4, s);
  rc = GRCUDA_EUNSUPPORTED;
  if (split && !h->keep_bytes)
    rc = mm_then_corr_launch(h->mm, h->corr, h->symbol_map.data(), (int)h->symbol_map.size(), 2, F, KEEP + R,
                             (long)(h->abs_row - KEEP), h->soft.as<float>(), h->sym.as<unsigned char>(), h->max_sym,
                             h->counts.as<int>(), (grcuda_hit*)h->hits.p, h->max_hits, h->nhits.as<int>(), s);
  if (rc == GRCUDA_EUNSUPPORTED)
    rc = mm_corr_launch(h->mm, h->corr, h->symbol_map.data(), (int)h->symbol_map.size(), 2, F, KEEP + R,
                        (long)(h->abs_row - KEEP), h->soft.as<float>(), h->sym.as<unsigned char>(), h->max_sym,
                        h->counts.as<int>(), h->keep_bytes ? h->bytes.as<unsigned char>() : nullptr,
                        (grcuda_hit*)h->hits.p, h->max_hits, h->nhits.as<int>(), s);
  if (rc) return rc;
  h->prof.end(s);
  GRB_CUDA(cudaEventRecord(h->ev_tail[cur], s));
  h->tail_pending[cur] = true;
  h->abs_row += R;
  h->last_rows = (int)R;
  h->prev_rows = (int)R;
  h->front_rows = 0;
  return GRCUDA_OK;
}

// The tail as its two kernels on two streams, for a time shard: the clock-recovery kernel is the serial
// chain over all blocks of all ranks, so nothing else sits on it -- it reads the loop state where the left neighbour's
// send landed (d_mm_state_in, nullptr: the chain's own) and writes its final state to d_mm_state_out as well (what
// the send to the right neighbour starts from); the correlator follows on its own stream with its own (small) state
// ring.  GRCUDA_EUNSUPPORTED when the time-parallel correlator does not apply (then use process_tail_device).
int grcuda_dmr_chain_process_tail_mm_device(grcuda_dmr_chain* h, const void* d_mm_state_in, void* d_mm_state_out, void* stream_) {
  if (h->front_rows <= 0) return set_error(GRCUDA_EINVAL, "dmr_chain: process_tail_mm without a pending process_front");
  if (h->keep_bytes) return GRCUDA_EUNSUPPORTED;
  cudaStream_t s = stream_ ? (cudaStream_t)stream_ : h->stream;
  const long R = h->front_rows;
  const int cur = h->fcur;
  int rc;
  float* F = h->Fb[cur].as<float>();
  GRB_CUDA(cudaStreamWaitEvent(s, h->ev_front[cur], 0));
  if (h->state_pending) { GRB_CUDA(cudaStreamWaitEvent(s, h->ev_state, 0)); h->state_pending = false; }
  // the symbol buffers are single: the previous block's correlator must have read them
  if (h->corr_pending) GRB_CUDA(cudaStreamWaitEvent(s, h->ev_corr, 0));
  if (h->tail_variant < 0) grcuda_clock_recovery_mm_ff_set_kernel_variant(h->mm, -1);
  h->prof.begin(4, s);
  if ((rc = mm_only_launch(h->mm, F, KEEP + R, (long)(h->abs_row - KEEP), h->soft.as<float>(), h->sym.as<unsigned char>(),
                           h->max_sym, h->counts.as<int>(), d_mm_state_in, d_mm_state_out, s)))
    return rc;
  h->prof.end(s);
  GRB_CUDA(cudaEventRecord(h->ev_mm, s));
  GRB_CUDA(cudaEventRecord(h->ev_tail[cur], s));  // F is only read by this kernel
  h->tail_pending[cur] = true;
  h->mm_pending = true;
  h->abs_row += R;
  h->last_rows = (int)R;
  h->prev_rows = (int)R;
  h->front_rows = 0;
  h->last_stream = s;
  return GRCUDA_OK;
}
int grcuda_dmr_chain_process_tail_corr_device(grcuda_dmr_chain* h, const void* d_corr_state_in, void* d_corr_state_out, void* stream_) {
  if (!h->mm_pending) return set_error(GRCUDA_EINVAL, "dmr_chain: process_tail_corr without a pending process_tail_mm");
  cudaStream_t s = stream_ ? (cudaStream_t)stream_ : h->stream;
  GRB_CUDA(cudaStreamWaitEvent(s, h->ev_mm, 0));
  if (!h->accumulate_hits) GRB_CUDA(cudaMemsetAsync(h->nhits.p, 0, sizeof(int), s));
  h->prof.begin(5, s);
  int rc = corr_par_launch(h->corr, h->symbol_map.data(), (int)h->symbol_map.size(), 2, h->sym.as<unsigned char>(), h->counts.as<int>(),
                           h->max_sym, (grcuda_hit*)h->hits.p, h->max_hits, h->nhits.as<int>(), d_corr_state_in, d_corr_state_out, s);
  if (rc) return rc;
  h->prof.end(s);
  GRB_CUDA(cudaEventRecord(h->ev_corr, s));
  h->corr_pending = true;
  h->mm_pending = false;
  h->last_stream = s;
  return GRCUDA_OK;
}
size_t grcuda_dmr_chain_mm_state_bytes(grcuda_dmr_chain* h) { return mm_state_bytes(h->mm); }
size_t grcuda_dmr_chain_corr_state_bytes(grcuda_dmr_chain* h) { return corr_state_bytes(h->corr); }

int grcuda_dmr_chain_process_device(grcuda_dmr_chain* h, const grcuda_complex* d_in, int nrows, void* stream_) {
  cudaStream_t s = stream_ ? (cudaStream_t)stream_ : h->stream;
  int rc = front_impl(h, d_in, nrows, s, h->pipeline);   // (pipeline: kernels sized to co-reside with the previous tail)
  if (rc) return rc;
  // the tail goes to the chain's own stream: it overlaps the front of the next block
  h->under_front = h->pipeline;
  rc = grcuda_dmr_chain_process_tail_device(h, h->pipeline ? (void*)h->tail_stream : stream_);
  h->under_front = false;
  return rc;
}

int grcuda_dmr_chain_set_tail_variant(grcuda_dmr_chain* h, int variant) {
  int rc = grcuda_clock_recovery_mm_ff_set_kernel_variant(h->mm, variant);
  if (!rc) h->tail_variant = variant;
  return rc;
}
// what would otherwise be silent: loop steps clamped at the first buffered row, blocks that hit the symbol capacity,
// and sync hits beyond the hit list's capacity (all zero for in-contract input); synchronises the device
int grcuda_dmr_chain_counters(grcuda_dmr_chain* h, long long* clamped, long long* overflow, long long* hits_dropped) {
  int rc = mm_counters(h->mm, clamped, overflow);
  if (rc) return rc;
  int n = 0;
  GRB_CUDA(cudaMemcpy(&n, h->nhits.p, sizeof(int), cudaMemcpyDeviceToHost));
  if (hits_dropped) *hits_dropped = n > h->max_hits ? n - h->max_hits : 0;
  return GRCUDA_OK;
}
int grcuda_dmr_chain_set_accumulate_hits(grcuda_dmr_chain* h, int on) {
  h->accumulate_hits = on != 0;
  return GRCUDA_OK;
}
int grcuda_dmr_chain_clear_hits(grcuda_dmr_chain* h, void* stream_) {
  cudaStream_t s = stream_ ? (cudaStream_t)stream_ : h->stream;
  GRB_CUDA(cudaMemsetAsync(h->nhits.p, 0, sizeof(int), s));
  return GRCUDA_OK;
}
int grcuda_dmr_chain_max_hits(grcuda_dmr_chain* h) { return h->max_hits; }
int grcuda_dmr_chain_set_split_correlator(grcuda_dmr_chain* h, int on) {
  h->split_corr = on != 0;
  h->split_user = true;
  return GRCUDA_OK;
}
int grcuda_dmr_chain_set_profiling(grcuda_dmr_chain* h, int on) {
  h->prof.on = on != 0;
  return grcuda_pfb_channelizer_ccf_set_profiling(h->pfb, on);
}
int grcuda_dmr_chain_profile_read(grcuda_dmr_chain* h, float* ms, int* launches) {
  float m[EventProfiler::kStages];
  int l[EventProfiler::kStages];
  h->prof.read(m, l);
  grcuda_pfb_channelizer_ccf_profile_read(h->pfb, m, l);  // fills stages 0 and 1
  for (int i = 0; i < GRCUDA_NSTAGES; i++) { if (ms) ms[i] = m[i]; if (launches) launches[i] = l[i]; }
  return GRCUDA_OK;
}

// Host entry point: pinned, double-buffered staging.  The block is cut into sub-blocks; while
// sub-block i runs on the compute stream, sub-block i+1 is DMA'd (and, for pageable callers,
// memcpy'd into pinned chunks) on the copy stream.  Hits of all sub-blocks accumulate.
int grcuda_dmr_chain_process_host(grcuda_dmr_chain* h, const grcuda_complex* in, int nrows) {
  const int minr = grcuda_dmr_chain_min_rows(h);
  if (nrows < minr || nrows > h->max_rows)
    return set_error(GRCUDA_EINVAL, "dmr_chain: nrows %d outside [%d, %d]", nrows, minr, h->max_rows);
  const size_t M = h->M;
  int nsub = h->keep_bytes ? 1 : std::max(1, std::min(8, nrows / std::max(minr, 256)));
  if (const char* e = getenv("GRCUDA_CHAIN_HOST_SUBBLOCKS")) nsub = std::max(1, std::min(atoi(e), nrows / std::max(minr, 1)));
  const int sub = (nrows + nsub - 1) / nsub;
  int rc;
  if (!h->copy_stream) {
    GRB_CUDA(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; i++) {
      GRB_CUDA(cudaEventCreateWithFlags(&h->ev_copied[i], cudaEventDisableTiming));
      GRB_CUDA(cudaEventCreateWithFlags(&h->ev_done[i], cudaEventDisableTiming));
    }
  }
  const size_t buf_bytes = (size_t)(h->T + sub) * M * sizeof(float2);
  for (int i = 0; i < 2; i++)
    if ((rc = h->d_stage[i].reserve(buf_bytes))) return rc;
  const bool user_accumulates = h->accumulate_hits;
  if (!user_accumulates) {
    cudaStream_t ts = h->pipeline ? h->tail_stream : h->stream;  // where the tails of the sub-blocks run
    GRB_CUDA(cudaMemsetAsync(h->nhits.p, 0, sizeof(int), ts));
  }
  h->accumulate_hits = true;
  int done = 0, i = 0;
  while (done < nrows) {
    int n = std::min(sub, nrows - done);
    if (nrows - (done + n) > 0 && nrows - (done + n) < minr) n = nrows - done;  // fold a short tail
    const int b = i & 1;
    if (n > sub && (rc = h->d_stage[b].reserve((size_t)(h->T + n) * M * sizeof(float2)))) { h->accumulate_hits = user_accumulates; return rc; }
    if (i >= 2) GRB_CUDA(cudaStreamWaitEvent(h->copy_stream, h->ev_done[b], 0));  // buffer b free again
    // host rows [done, done + T + n) : T history rows + n new rows of this sub-block
    if ((rc = h->stager.h2d(h->d_stage[b].p, (const char*)in + (size_t)done * M * sizeof(float2),
                            (size_t)(h->T + n) * M * sizeof(float2), h->copy_stream))) { h->accumulate_hits = user_accumulates; return rc; }
    GRB_CUDA(cudaEventRecord(h->ev_copied[b], h->copy_stream));
    GRB_CUDA(cudaStreamWaitEvent(h->stream, h->ev_copied[b], 0));
    if ((rc = grcuda_dmr_chain_process_device(h, (const grcuda_complex*)h->d_stage[b].p, n, h->stream))) { h->accumulate_hits = user_accumulates; return rc; }
    GRB_CUDA(cudaEventRecord(h->ev_done[b], h->stream));
    done += n;
    i++;
  }
  h->accumulate_hits = user_accumulates;
  GRB_CUDA(cudaStreamSynchronize(h->stream));
  if (h->last_stream) GRB_CUDA(cudaStreamSynchronize(h->last_stream));
  return GRCUDA_OK;
}

int grcuda_dmr_chain_result_get(grcuda_dmr_chain* h, grcuda_dmr_chain_result* r) {
  r->d_channels = (h->fft_demod && !h->fft_demod_store_y) ? nullptr : (const grcuda_complex*)(h->Yb.as<float2>() + (size_t)h->YH * h->M);
  r->d_soft = h->soft.as<float>();
  r->d_symbols = h->sym.as<unsigned char>();
  r->d_sym_counts = h->counts.as<int>();
  r->d_bytes = h->keep_bytes ? h->bytes.as<unsigned char>() : nullptr;
  r->d_hits = (const grcuda_hit*)h->hits.p;
  r->d_nhits = h->nhits.as<int>();
  r->max_sym = h->max_sym;
  r->nrows = h->last_rows;
  return GRCUDA_OK;
}

int grcuda_dmr_chain_read_hits(grcuda_dmr_chain* h, grcuda_hit* hits, int max_hits) {
  int n = 0;
  cudaStream_t s = h->last_stream ? h->last_stream : h->stream;  // ordered after the last process call
  GRB_CUDA(cudaMemcpyAsync(&n, h->nhits.p, sizeof(int), cudaMemcpyDeviceToHost, s));
  GRB_CUDA(cudaStreamSynchronize(s));
  n = std::min(n, h->max_hits);
  const int m = std::min(n, max_hits);
  if (m > 0) {
    GRB_CUDA(cudaMemcpyAsync(hits, h->hits.p, (size_t)m * sizeof(grcuda_hit), cudaMemcpyDeviceToHost, s));
    GRB_CUDA(cudaStreamSynchronize(s));
  }
  return n;
}

}  // extern "C"
