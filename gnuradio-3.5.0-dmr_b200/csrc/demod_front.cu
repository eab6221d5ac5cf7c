// Host launcher of the fused discriminator + matched-filter kernel (kernel_demod_front.cuh).
#include <algorithm>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "internal.h"
#include "kernel_demod_front.cuh"

namespace grb {

typedef void (*df_kernel_t)(const DemodFrontArgs);

template <int RHO> static df_kernel_t pick_qm(int qm) {
  switch (qm) {
    case 0: return demod_front_kernel<RHO, 0>;
    case 1: return demod_front_kernel<RHO, 1>;
    case 2: return demod_front_kernel<RHO, 2>;
    default: return demod_front_kernel<RHO, 3>;
  }
}

int demod_front_max_taps() { return 4 * DF_MAXB - 7; }
int demod_front_history(int ntaps) { return ntaps + 4; }  // rows of Y that must precede the block

// host image of the kernel's tap store: tp[al][p] = rt[p - al] (rt = REVERSED taps), zero padded
std::vector<float> demod_front_tap_table(const float* rt, int ntaps) {
  std::vector<float> tp((size_t)4 * DF_MAXB * 4, 0.f);
  for (int al = 0; al < 4; al++)
    for (int p = 0; p < DF_MAXB * 4; p++) {
      const int i = p - al;
      if (i >= 0 && i < ntaps) tp[(size_t)al * DF_MAXB * 4 + p] = rt[i];
    }
  return tp;
}

// y: [hist + nrows][M] complex with hist = demod_front_history(ntaps); d_tp = device copy of
// demod_front_tap_table().
// dsrc != nullptr: [hist + nrows][M] floats, the discriminator output made elsewhere (row numbering as y); y unused.
int demod_front_launch(const float2* y, float* f, long abs_row0, int nrows, int M, float gain, const float* d_tp, int ntaps,
                       cudaStream_t s, const float* dsrc, const float* h_tp) {
  if (ntaps < 1 || ntaps > demod_front_max_taps() || !d_tp) return set_error(GRCUDA_EUNSUPPORTED, "demod_front: %d taps", ntaps);
  if (nrows <= 0) return GRCUDA_OK;
  DeviceTables tabs;
  int rc = get_tables(&tabs);
  if (rc) return rc;
  DemodFrontArgs a;
  a.dsrc = dsrc; a.y = y; a.f = f; a.abs_row0 = abs_row0; a.nrows = nrows; a.M = M; a.hist = demod_front_history(ntaps);
  a.gain = gain; a.one = 1.0f; a.atan_table = tabs.atan; a.ntaps = ntaps; a.tp = d_tp;
  const int n1 = ntaps - 1, rho = n1 & 3;
  a.q = n1 >> 2;
  df_kernel_t k;
  switch (rho) {
    case 0: k = pick_qm<0>(a.q & 3); break;
    case 1: k = pick_qm<1>(a.q & 3); break;
    case 2: k = pick_qm<2>(a.q & 3); break;
    default: k = pick_qm<3>(a.q & 3); break;
  }
  const int J = a.q + 1 + (rho > 0 ? 1 : 0);
  memset(a.tpc, 0, sizeof a.tpc);
  if (J == 8 && rho == 0 && h_tp) {
    k = demod_front_kernel<0, 3, 8>;
    for (int al = 0; al < 4; al++)
      for (int b = 0; b < 8; b++) memcpy(&a.tpc[al][b][0], h_tp + (size_t)al * DF_MAXB * 4 + b * 4, 16);
  }   // ntaps = 29: the chain's matched filter (11 symbols at 2.6 samples + 1)
  const size_t smem = ((size_t)4 * DF_MAXB * 4 + 260 + (size_t)(DF_RT + 4 * (J - 1)) * 32) * sizeof(float);
  if (smem > 48 * 1024) GRB_CUDA(raise_dynamic_smem((const void*)k, (size_t)smem));
  const long A0 = (abs_row0 >> 2) << 2;
  const long ntiles = (abs_row0 + nrows - A0 + DF_RT - 1) / DF_RT;
  dim3 grid((M + 31) / 32, (unsigned)ntiles);
  k<<<grid, DF_THREADS, smem, s>>>(a);
  GRB_LAUNCH_CHECK();
  return GRCUDA_OK;
}

}  // namespace grb

// ---- C ABI: the fused pair as a stand-alone batched op (include/gr_cuda.h) -----------------------
extern "C" {

int grcuda_quad_demod_fir_fff_history(grcuda_fir_fff* f) {
  int nt = 0;
  grb::fir_fff_reversed_taps(f, &nt, nullptr);
  return grb::demod_front_history(nt);
}

int grcuda_quad_demod_fir_fff_work_device(grcuda_quad* q, grcuda_fir_fff* f, long nrows, int nchan,
                                          const grcuda_complex* d_in, float* d_out, long abs_row0, void* stream) {
  int nt = 0, order = 0;
  grb::fir_fff_reversed_taps(f, &nt, &order);
  if (order != GRCUDA_ORDER_SSE)
    return grb::set_error(GRCUDA_EUNSUPPORTED, "quad_demod_fir_fff: only the SSE summation order is fused");
  if (nrows > 0x7fffffffL || nchan < 1) return grb::set_error(GRCUDA_EINVAL, "quad_demod_fir_fff: bad shape");
  return grb::demod_front_launch((const float2*)d_in, d_out, abs_row0, (int)nrows, nchan, grb::quad_gain(q),
                                 grb::fir_fff_front_taps(f), nt, (cudaStream_t)stream, nullptr, grb::fir_fff_front_taps_host(f));
}

}  // extern "C"

