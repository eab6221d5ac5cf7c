// oracle shim (test infrastructure): minimal stand-in for the reference runtime's
// gr_io_signature (gnuradio-core/src/lib/runtime/gr_io_signature.h) -- records the
// declared stream shapes only.
#pragma once
#include <memory>
#include <vector>
class gr_io_signature {
  int d_min, d_max, d_size;
 public:
  gr_io_signature(int mn, int mx, int sz) : d_min(mn), d_max(mx), d_size(sz) {}
  int min_streams() const { return d_min; }
  int max_streams() const { return d_max; }
  int sizeof_stream_item(int) const { return d_size; }
};
typedef std::shared_ptr<gr_io_signature> gr_io_signature_sptr;
inline gr_io_signature_sptr gr_make_io_signature(int mn, int mx, int sz) {
  return gr_io_signature_sptr(new gr_io_signature(mn, mx, sz));
}
// two differently sized streams (runtime/gr_io_signature.h gr_make_io_signature2): the shim keeps the first size
inline gr_io_signature_sptr gr_make_io_signature2(int mn, int mx, int sz1, int) {
  return gr_io_signature_sptr(new gr_io_signature(mn, mx, sz1));
}
