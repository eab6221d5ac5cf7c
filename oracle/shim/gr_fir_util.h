// oracle shim (test infrastructure): stands in for gr_fir_util / gr_fir_sysconfig_x86
// (gnuradio-core/src/lib/filter/gr_fir_sysconfig_x86.cc:175-201), which on x86-64 with
// SSE returns the *_sse subclasses.  g_grref_fir_impl selects the same SSE classes (1,
// default) or the portable *_generic ones (0) so both reference code paths are witnesses.
#pragma once
#include <vector>
#include <gr_complex.h>
#include <gr_fir_ccf_generic.h>
#include <gr_fir_fff_generic.h>
#include <gr_fir_ccc_generic.h>
#include <gr_fir_ccf_x86.h>
#include <gr_fir_fff_x86.h>
#include <gr_fir_ccc_x86.h>
extern int g_grref_fir_impl;
struct gr_fir_util {
  static gr_fir_ccf* create_gr_fir_ccf(const std::vector<float>& t) {
    if (g_grref_fir_impl) return new gr_fir_ccf_sse(t);
    return new gr_fir_ccf_generic(t);
  }
  static gr_fir_fff* create_gr_fir_fff(const std::vector<float>& t) {
    if (g_grref_fir_impl) return new gr_fir_fff_sse(t);
    return new gr_fir_fff_generic(t);
  }
  static gr_fir_ccc* create_gr_fir_ccc(const std::vector<gr_complex>& t) {
    if (g_grref_fir_impl) return new gr_fir_ccc_sse(t);
    return new gr_fir_ccc_generic(t);
  }
};
