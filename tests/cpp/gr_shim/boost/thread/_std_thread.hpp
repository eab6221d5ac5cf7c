#pragma once
#include <condition_variable>
#include <mutex>
#include <thread>
namespace boost {
using std::thread; using std::condition_variable;
struct mutex : std::mutex { typedef std::unique_lock<mutex> scoped_lock; };
template <class M> using unique_lock = std::unique_lock<M>;
}
