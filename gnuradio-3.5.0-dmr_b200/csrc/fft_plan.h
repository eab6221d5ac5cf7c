// Host-side FFT plan (the GPU counterpart of gri_fft_complex's FFTW plan,
// gnuradio-core/src/lib/general/gri_fft.cc:97-146): picks radices, builds twiddle tables,
// launches the batched shared-memory engine of fft_engine.cuh.
#pragma once
#include <cuda_runtime.h>

namespace grb {

struct FftPlan;
// dir: -1 forward, +1 backward (unnormalised).  Returns nullptr with the error set.
// coresident: prefer the register-capped build of the plan (where one exists): slower on an empty
// machine, but its CTAs fit on an SM next to a CTA of the clock-recovery kernel, which the chain
// runs concurrently with the next block's front.
FftPlan* fft_plan_create(int n, int dir, bool coresident = false);
void fft_plan_destroy(FftPlan* p);
// rows: [nrows][n] complex.  window: n device floats or nullptr.  in_rot/out_rot: element
// rotations implementing ifftshift on load / fftshift on store.
int fft_plan_exec(FftPlan* p, const float2* d_in, float2* d_out, long nrows, const float* d_window, int in_rot,
                  int out_rot, cudaStream_t stream);
const char* fft_plan_describe(FftPlan* p);
// The backward transform with gr_quadrature_demod_cf fused into its last pass (kernel_fft_demod.cuh): rows in,
// discriminator rows (float) out; the transform itself is never stored.  Exists for the three-pass plans whose last
// pass gives every thread the same channels in every row (8000 = 20^3, 4096 = 16^3).
bool fft_plan_demod_supported(FftPlan* p);
int fft_plan_exec_demod(FftPlan* p, const float2* d_in, float* d_d, long nrows, float gain, const float* d_atan_table,
                        const float2* d_prev_y, float2* d_last_y, bool coresident, cudaStream_t stream, float2* d_y_out = nullptr);

}  // namespace grb
