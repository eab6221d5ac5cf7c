// Non-ABI accessors between the translation units of libgr_cuda (hidden visibility).
#pragma once
#include <cstddef>
#include <vector>
#include <cuda_runtime.h>
#include "../../include/gr_cuda.h"

namespace grb {
void* mm_state_ptr(grcuda_mm* h);
size_t mm_state_bytes(grcuda_mm* h);
int mm_counters(grcuda_mm* h, long long* clamped, long long* overflow);
void* corr_state_ptr(grcuda_corr* h);
size_t corr_state_bytes(grcuda_corr* h);
int mm_corr_launch(grcuda_mm* mm, grcuda_corr* corr, const int* map, int nmap, int bits_per_symbol, const float* d_in,
                   long ninput, long abs_row0, float* d_soft, unsigned char* d_sym, int max_out, int* d_counts,
                   unsigned char* d_bytes, grcuda_hit* d_hits, int max_hits, int* d_nhits, cudaStream_t s);
int mm_then_corr_launch(grcuda_mm* mm, grcuda_corr* corr, const int* map, int nmap, int bits_per_symbol, const float* d_in,
                        long ninput, long abs_row0, float* d_soft, unsigned char* d_sym, int max_out, int* d_counts,
                        grcuda_hit* d_hits, int max_hits, int* d_nhits, cudaStream_t s);
int mm_only_launch(grcuda_mm* mm, const float* d_in, long ninput, long abs_row0, float* d_soft, unsigned char* d_sym, int max_out,
                   int* d_counts, const void* d_state_in, void* d_state_out, cudaStream_t s);
int corr_par_launch(grcuda_corr* corr, const int* map, int nmap, int bits_per_symbol, const unsigned char* d_sym, const int* d_counts,
                    int max_out, grcuda_hit* d_hits, int max_hits, int* d_nhits, const void* d_state_in, void* d_state_out,
                    cudaStream_t s);
// fused quadrature_demod_cf + fir_filter_fff (SSE order) on [time][channel] data (demod_front.cu)
int demod_front_max_taps();
int demod_front_history(int ntaps);
std::vector<float> demod_front_tap_table(const float* rt, int ntaps);  // host image of the kernel's tap store
int demod_front_launch(const float2* y, float* f, long abs_row0, int nrows, int M, float gain, const float* d_tp, int ntaps,
                       cudaStream_t s, const float* dsrc = nullptr, const float* h_tp = nullptr);
const float* fir_fff_front_taps_host(grcuda_fir_fff* h);  // host image of the same table (kernel-parameter taps)
// channelizer with the discriminator inside the FFT kernel (kernel_fft_demod.cuh): rows in -> D rows out (float), the
// channelizer output itself never reaches HBM.  prev_y: [M] channelizer output of the row before the first one
// (zeros at stream start); last_y receives that of the last row.  GRCUDA_EUNSUPPORTED when the plan has no such kernel.
int pfb_demod_supported(grcuda_pfb* h);
int pfb_work_device_demod(grcuda_pfb* h, long nrows, const float2* d_in_rows, float* d_D, float gain, const float2* prev_y,
                          float2* last_y, bool coresident, cudaStream_t s, float2* y_out = nullptr);
const float* fir_fff_front_taps(grcuda_fir_fff* h);  // device copy of demod_front_tap_table (nullptr: too many taps)
// reversed taps / order / gain of the stand-alone plans (host copies)
const float* fir_fff_reversed_taps(grcuda_fir_fff* h, int* ntaps, int* order);
float quad_gain(grcuda_quad* h);
int pfb_reserve_rows(grcuda_pfb* h, long rows);
int pfb_prefer_coresident_fft(grcuda_pfb* h);
}  // namespace grb
