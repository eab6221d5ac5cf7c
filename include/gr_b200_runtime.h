/* Minimal stand-in for the part of the GNU Radio 3.5 runtime a block sees, so that the GPU blocks
 * in gr_b200_blocks.h (and their tests) compile without Boost / SWIG / the scheduler.
 *
 * It mirrors, member for member, what the hot-path blocks use of
 *   gnuradio-core/src/lib/runtime/gr_block.h:58-326      (gr_block: history, forecast, general_work,
 *                                                         consume, output_multiple, relative_rate)
 *   gnuradio-core/src/lib/runtime/gr_sync_block.cc:58-68  (work() + consume_each)
 *   gnuradio-core/src/lib/runtime/gr_sync_decimator.cc:58-68
 *   gnuradio-core/src/lib/runtime/gr_io_signature.h, gr_types.h:33-37, gr_complex.h:26
 *
 * Inside a real GNU Radio 3.5 tree, define GR_B200_USE_GNURADIO_RUNTIME before including
 * gr_b200_blocks.h: the blocks then derive from the installed gr_block / gr_sync_block /
 * gr_sync_decimator and this header is not used.
 */
#ifndef INCLUDED_GR_B200_RUNTIME_H
#define INCLUDED_GR_B200_RUNTIME_H

#include <cmath>
#include <complex>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

typedef std::complex<float> gr_complex;                     /* runtime/gr_complex.h:26 */
typedef std::vector<int> gr_vector_int;                     /* runtime/gr_types.h:33-37 */
typedef std::vector<const void*> gr_vector_const_void_star;
typedef std::vector<void*> gr_vector_void_star;

class gr_io_signature {
  int d_min, d_max, d_size;
 public:
  gr_io_signature(int min_streams, int max_streams, int sizeof_stream_item)
      : d_min(min_streams), d_max(max_streams), d_size(sizeof_stream_item) {
    if (min_streams < 0 || (max_streams != -1 && max_streams < min_streams))
      throw std::invalid_argument("gr_io_signature(2)");  /* gr_io_signature.cc:74-76 */
  }
  int min_streams() const { return d_min; }
  int max_streams() const { return d_max; }
  int sizeof_stream_item(int) const { return d_size; }
};
typedef std::shared_ptr<gr_io_signature> gr_io_signature_sptr;
inline gr_io_signature_sptr gr_make_io_signature(int min_streams, int max_streams, int sizeof_stream_item) {
  return gr_io_signature_sptr(new gr_io_signature(min_streams, max_streams, sizeof_stream_item));
}

class gr_block {
 public:
  enum { WORK_CALLED_PRODUCE = -2, WORK_DONE = -1 };  /* gr_block.h:63-66 */
  virtual ~gr_block() {}
  const std::string& name() const { return d_name; }
  gr_io_signature_sptr input_signature() const { return d_in_sig; }
  gr_io_signature_sptr output_signature() const { return d_out_sig; }

  unsigned history() const { return d_history; }            /* gr_block.h:83-84 */
  void set_history(unsigned history) { d_history = history; }
  bool fixed_rate() const { return d_fixed_rate; }           /* :92 */
  virtual void forecast(int noutput_items, gr_vector_int& ninput_items_required) {  /* gr_block.cc:50-56 */
    for (size_t i = 0; i < ninput_items_required.size(); i++) ninput_items_required[i] = noutput_items + history() - 1;
  }
  virtual bool start() { return true; }
  virtual bool stop() { return true; }
  virtual int general_work(int noutput_items, gr_vector_int& ninput_items, gr_vector_const_void_star& input_items,
                           gr_vector_void_star& output_items) = 0;
  void set_output_multiple(int multiple) {                   /* gr_block.cc:59-65 */
    if (multiple < 1) throw std::invalid_argument("gr_block::set_output_multiple");
    d_output_multiple = multiple;
  }
  int output_multiple() const { return d_output_multiple; }
  void consume(int which_input, int how_many_items) {        /* gr_block.cc:79-90 (records instead of moving a gr_buffer_reader) */
    if ((size_t)which_input >= d_consumed.size()) d_consumed.resize(which_input + 1, 0);
    d_consumed[which_input] += how_many_items;
  }
  void consume_each(int how_many_items) {
    d_consume_each += how_many_items;
  }
  void set_relative_rate(double relative_rate) {             /* gr_block.cc:102-108 */
    if (relative_rate < 0.0) throw std::invalid_argument("gr_block::set_relative_rate");
    d_relative_rate = relative_rate;
  }
  double relative_rate() const { return d_relative_rate; }

  /* what a scheduler (or a test harness) reads back after general_work: items consumed on input i */
  int b200_consumed(int which_input) const {
    return d_consume_each + ((size_t)which_input < d_consumed.size() ? d_consumed[which_input] : 0);
  }
  void b200_reset_consumed() { d_consumed.assign(d_consumed.size(), 0); d_consume_each = 0; }

 protected:
  gr_block(const std::string& name, gr_io_signature_sptr input_signature, gr_io_signature_sptr output_signature)
      : d_name(name), d_in_sig(input_signature), d_out_sig(output_signature), d_output_multiple(1), d_relative_rate(1.0),
        d_history(1), d_fixed_rate(false), d_consume_each(0) {}
  void set_fixed_rate(bool fixed_rate) { d_fixed_rate = fixed_rate; }

 private:
  std::string d_name;
  gr_io_signature_sptr d_in_sig, d_out_sig;
  int d_output_multiple;
  double d_relative_rate;
  unsigned d_history;
  bool d_fixed_rate;
  std::vector<int> d_consumed;
  int d_consume_each;
};
typedef std::shared_ptr<gr_block> gr_block_sptr;

class gr_sync_block : public gr_block {                       /* runtime/gr_sync_block.cc:30-68 */
 protected:
  gr_sync_block(const std::string& name, gr_io_signature_sptr in, gr_io_signature_sptr out) : gr_block(name, in, out) {
    set_fixed_rate(true);
  }
 public:
  virtual int work(int noutput_items, gr_vector_const_void_star& input_items, gr_vector_void_star& output_items) = 0;
  void forecast(int noutput_items, gr_vector_int& ninput_items_required) {
    for (size_t i = 0; i < ninput_items_required.size(); i++) ninput_items_required[i] = noutput_items + history() - 1;
  }
  int general_work(int noutput_items, gr_vector_int&, gr_vector_const_void_star& input_items, gr_vector_void_star& output_items) {
    int r = work(noutput_items, input_items, output_items);
    if (r > 0) consume_each(r);
    return r;
  }
};

class gr_sync_decimator : public gr_sync_block {              /* runtime/gr_sync_decimator.cc:30-68 */
  unsigned d_decimation;
 protected:
  gr_sync_decimator(const std::string& name, gr_io_signature_sptr in, gr_io_signature_sptr out, unsigned decimation)
      : gr_sync_block(name, in, out) {
    set_decimation(decimation);
  }
 public:
  unsigned decimation() const { return d_decimation; }
  void set_decimation(unsigned decimation) {
    d_decimation = decimation;
    set_relative_rate(1.0 / decimation);
  }
  void forecast(int noutput_items, gr_vector_int& ninput_items_required) {
    for (size_t i = 0; i < ninput_items_required.size(); i++)
      ninput_items_required[i] = noutput_items * decimation() + history() - 1;
  }
  int general_work(int noutput_items, gr_vector_int&, gr_vector_const_void_star& input_items, gr_vector_void_star& output_items) {
    int r = work(noutput_items, input_items, output_items);
    if (r > 0) consume_each(r * decimation());
    return r;
  }
};

class gr_sync_interpolator : public gr_sync_block {           /* runtime/gr_sync_interpolator.cc:30-70 */
  unsigned d_interpolation;
 protected:
  gr_sync_interpolator(const std::string& name, gr_io_signature_sptr in, gr_io_signature_sptr out, unsigned interpolation)
      : gr_sync_block(name, in, out) {
    set_interpolation(interpolation);
  }
 public:
  unsigned interpolation() const { return d_interpolation; }
  void set_interpolation(unsigned interpolation) {
    d_interpolation = interpolation;
    set_relative_rate(1.0 * interpolation);
    set_output_multiple(interpolation);
  }
  void forecast(int noutput_items, gr_vector_int& ninput_items_required) {
    for (size_t i = 0; i < ninput_items_required.size(); i++)
      ninput_items_required[i] = noutput_items / interpolation() + history() - 1;
  }
  int general_work(int noutput_items, gr_vector_int&, gr_vector_const_void_star& input_items, gr_vector_void_star& output_items) {
    int r = work(noutput_items, input_items, output_items);
    if (r > 0) consume_each(r / interpolation());
    return r;
  }
};

/* gr_message / gr_msg_queue as far as gr_framer_sink_1 uses them (runtime/gr_message.h:35-92, gr_msg_queue.h:34-88):
 * gr_make_message(type, arg1, arg2, length), msg(), insert_tail(), delete_head_nowait(), count(). */
class gr_message;
typedef std::shared_ptr<gr_message> gr_message_sptr;
class gr_message {
  long d_type;
  double d_arg1, d_arg2;
  std::vector<unsigned char> d_buf;
  friend gr_message_sptr gr_make_message(long type, double arg1, double arg2, size_t length);
  gr_message(long type, double arg1, double arg2, size_t length) : d_type(type), d_arg1(arg1), d_arg2(arg2), d_buf(length) {}
 public:
  long type() const { return d_type; }
  double arg1() const { return d_arg1; }
  double arg2() const { return d_arg2; }
  unsigned char* msg() { return d_buf.data(); }
  size_t length() const { return d_buf.size(); }
  std::string to_string() const { return std::string(d_buf.begin(), d_buf.end()); }
};
inline gr_message_sptr gr_make_message(long type, double arg1 = 0, double arg2 = 0, size_t length = 0) {
  return gr_message_sptr(new gr_message(type, arg1, arg2, length));
}
class gr_msg_queue {
  std::vector<gr_message_sptr> d_q;
  size_t d_head;
 public:
  gr_msg_queue() : d_head(0) {}
  void insert_tail(gr_message_sptr msg) { d_q.push_back(msg); }
  gr_message_sptr delete_head_nowait() { return d_head < d_q.size() ? d_q[d_head++] : gr_message_sptr(); }
  bool empty_p() const { return d_head >= d_q.size(); }
  unsigned count() const { return (unsigned)(d_q.size() - d_head); }
};
typedef std::shared_ptr<gr_msg_queue> gr_msg_queue_sptr;
inline gr_msg_queue_sptr gr_make_msg_queue(unsigned limit = 0) { (void)limit; return gr_msg_queue_sptr(new gr_msg_queue()); }

#endif /* INCLUDED_GR_B200_RUNTIME_H */
