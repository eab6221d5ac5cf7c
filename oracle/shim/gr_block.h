// oracle shim (test infrastructure): the slice of gr_block
// (gnuradio-core/src/lib/runtime/gr_block.h:58-326) that the hot-path blocks touch.
// The scheduler is NOT reproduced; the harness calls general_work() directly with
// history-prefixed buffers, exactly what gr_block_executor hands a block.
#pragma once
#include <gr_core_api.h>
#include <gr_types.h>
#include <gr_io_signature.h>
#include <string>
#include <stdexcept>
#include <cmath>
#include <memory>

namespace gnuradio {
template <class T> std::shared_ptr<T> get_initial_sptr(T* p) { return std::shared_ptr<T>(p); }
}

class gr_block {
 public:
  enum { WORK_CALLED_PRODUCE = -2, WORK_DONE = -1 };
  virtual ~gr_block() {}
  std::string name() const { return d_name; }
  unsigned history() const { return d_history; }
  void set_history(unsigned h) { d_history = h; }
  void set_relative_rate(double r) { d_relative_rate = r; }
  double relative_rate() const { return d_relative_rate; }
  void set_output_multiple(int m) { d_output_multiple = m; }
  int output_multiple() const { return d_output_multiple; }
  void set_fixed_rate(bool f) { d_fixed_rate = f; }
  bool fixed_rate() const { return d_fixed_rate; }
  void consume_each(int n) { d_consumed = n; }
  void consume(int, int n) { d_consumed = n; }
  int consumed() const { return d_consumed; }
  gr_io_signature_sptr input_signature() const { return d_in; }
  gr_io_signature_sptr output_signature() const { return d_out; }
  // default forecast: gnuradio-core/src/lib/runtime/gr_block.cc:50-56
  virtual void forecast(int noutput_items, gr_vector_int& req) {
    for (size_t i = 0; i < req.size(); i++) req[i] = noutput_items + history() - 1;
  }
  virtual int general_work(int noutput_items, gr_vector_int& ninput_items,
                           gr_vector_const_void_star& input_items,
                           gr_vector_void_star& output_items) = 0;
 protected:
  gr_block(const std::string& name, gr_io_signature_sptr in, gr_io_signature_sptr out)
      : d_name(name), d_in(in), d_out(out) {}
 private:
  std::string d_name;
  gr_io_signature_sptr d_in, d_out;
  unsigned d_history = 1;
  double d_relative_rate = 1.0;
  int d_output_multiple = 1;
  bool d_fixed_rate = false;
  int d_consumed = 0;
};
typedef std::shared_ptr<gr_block> gr_block_sptr;
