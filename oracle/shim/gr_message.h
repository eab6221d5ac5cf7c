/* ORACLE / TEST INFRASTRUCTURE.  Stand-in for runtime/gr_message.h (the real one pulls in gruel/thread.h -> Boost):
 * just what gr_framer_sink_1.cc uses -- gr_make_message(type, arg1, arg2, length), msg(), length(), arg1(). */
#ifndef ORACLE_SHIM_GR_MESSAGE_H
#define ORACLE_SHIM_GR_MESSAGE_H
#include <boost/shared_ptr.hpp>
#include <vector>
#include <cstddef>
class gr_message;
typedef boost::shared_ptr<gr_message> gr_message_sptr;
class gr_message {
  long d_type;
  double d_arg1, d_arg2;
  std::vector<unsigned char> d_buf;
 public:
  gr_message(long type, double arg1, double arg2, size_t length) : d_type(type), d_arg1(arg1), d_arg2(arg2), d_buf(length) {}
  long type() const { return d_type; }
  double arg1() const { return d_arg1; }
  double arg2() const { return d_arg2; }
  unsigned char* msg() { return d_buf.data(); }
  size_t length() const { return d_buf.size(); }
};
inline gr_message_sptr gr_make_message(long type = 0, double arg1 = 0, double arg2 = 0, size_t length = 0) {
  return gr_message_sptr(new gr_message(type, arg1, arg2, length));
}
#endif
