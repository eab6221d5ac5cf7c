/* ORACLE / TEST INFRASTRUCTURE.  Stand-in for runtime/gr_msg_queue.h: an unbounded, single-threaded FIFO with the
 * calls gr_framer_sink_1.cc and the harness make (insert_tail, delete_head_nowait, count, empty_p). */
#ifndef ORACLE_SHIM_GR_MSG_QUEUE_H
#define ORACLE_SHIM_GR_MSG_QUEUE_H
#include <gr_message.h>
#include <deque>
class gr_msg_queue;
typedef boost::shared_ptr<gr_msg_queue> gr_msg_queue_sptr;
class gr_msg_queue {
  std::deque<gr_message_sptr> d_q;
 public:
  void insert_tail(gr_message_sptr m) { d_q.push_back(m); }
  gr_message_sptr delete_head_nowait() {
    if (d_q.empty()) return gr_message_sptr();
    gr_message_sptr m = d_q.front();
    d_q.pop_front();
    return m;
  }
  unsigned int count() const { return (unsigned int)d_q.size(); }
  bool empty_p() const { return d_q.empty(); }
};
inline gr_msg_queue_sptr gr_make_msg_queue(unsigned int = 0) { return gr_msg_queue_sptr(new gr_msg_queue); }
#endif
