// oracle shim (test infrastructure): gr_sync_interpolator
// (gnuradio-core/src/lib/runtime/gr_sync_interpolator.cc): 1:N work(), consume r/interp.
#pragma once
#include <gr_sync_block.h>
class gr_sync_interpolator : public gr_sync_block {
  unsigned d_interpolation;
 protected:
  gr_sync_interpolator(const std::string& name, gr_io_signature_sptr in, gr_io_signature_sptr out,
                       unsigned interpolation)
      : gr_sync_block(name, in, out), d_interpolation(interpolation) {
    set_relative_rate(1.0 * interpolation);
    set_output_multiple(interpolation);
  }
 public:
  unsigned interpolation() const { return d_interpolation; }
  int general_work(int noutput_items, gr_vector_int&, gr_vector_const_void_star& in,
                   gr_vector_void_star& out) {
    int r = work(noutput_items, in, out);
    if (r > 0) consume_each(r / d_interpolation);
    return r;
  }
};
