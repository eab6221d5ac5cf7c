// oracle shim (test infrastructure): gr_sync_block
// (gnuradio-core/src/lib/runtime/gr_sync_block.cc:58-68): 1:1 work(), auto-consume.
#pragma once
#include <gr_block.h>
class gr_sync_block : public gr_block {
 protected:
  gr_sync_block(const std::string& name, gr_io_signature_sptr in, gr_io_signature_sptr out)
      : gr_block(name, in, out) { set_fixed_rate(true); }
 public:
  virtual int work(int noutput_items, gr_vector_const_void_star& input_items,
                   gr_vector_void_star& output_items) = 0;
  int general_work(int noutput_items, gr_vector_int&, gr_vector_const_void_star& in,
                   gr_vector_void_star& out) {
    int r = work(noutput_items, in, out);
    if (r > 0) consume_each(r);
    return r;
  }
};
