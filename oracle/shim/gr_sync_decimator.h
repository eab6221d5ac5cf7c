// oracle shim (test infrastructure): gr_sync_decimator
// (gnuradio-core/src/lib/runtime/gr_sync_decimator.cc:58-68): consume r*decimation.
#pragma once
#include <gr_sync_block.h>
class gr_sync_decimator : public gr_sync_block {
  unsigned d_decimation;
 protected:
  gr_sync_decimator(const std::string& name, gr_io_signature_sptr in, gr_io_signature_sptr out,
                    unsigned decimation)
      : gr_sync_block(name, in, out), d_decimation(decimation) {
    set_relative_rate(1.0 / decimation);
  }
 public:
  unsigned decimation() const { return d_decimation; }
  int general_work(int noutput_items, gr_vector_int&, gr_vector_const_void_star& in,
                   gr_vector_void_star& out) {
    int r = work(noutput_items, in, out);
    if (r > 0) consume_each(r * d_decimation);
    return r;
  }
};
