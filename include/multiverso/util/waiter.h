// Waiter: the completion latch of one table request -- armed with the number of per-server
// partitions, released by the replies (reference: include/multiverso/util/waiter.h:9-33).
// WaitFor() adds a bounded wait so that a stalled request can be reported (which table, which
// request, how many servers still owe a reply) instead of hanging silently: a dead server or a BSP
// schedule with unequal step counts shows up in the log of the waiting rank.
#ifndef MULTIVERSO_UTIL_WAITER_H_
#define MULTIVERSO_UTIL_WAITER_H_
#include <chrono>
#include <condition_variable>
#include <mutex>

namespace multiverso {

class Waiter {
 public:
  explicit Waiter(int num_wait = 1) : outstanding_(num_wait) {}

  // Blocks until every expected Notify() has arrived.
  void Wait() {
    std::unique_lock<std::mutex> lk(mu_);
    cv_.wait(lk, [&] { return outstanding_ <= 0; });
  }
  // Same with a time limit; false when replies are still outstanding after `seconds`.
  bool WaitFor(double seconds) {
    std::unique_lock<std::mutex> lk(mu_);
    return cv_.wait_for(lk, std::chrono::duration<double>(seconds), [&] { return outstanding_ <= 0; });
  }
  void Notify() {
    std::lock_guard<std::mutex> lk(mu_);
    if (--outstanding_ <= 0) cv_.notify_all();
  }
  // (Re)arm with the number of replies to expect; zero completes the request at once.
  void Reset(int num_wait) {
    std::lock_guard<std::mutex> lk(mu_);
    outstanding_ = num_wait;
    if (outstanding_ <= 0) cv_.notify_all();
  }
  int outstanding() {
    std::lock_guard<std::mutex> lk(mu_);
    return outstanding_;
  }
  bool Done() { return outstanding() <= 0; }

 private:
  std::mutex mu_;
  std::condition_variable cv_;
  int outstanding_;
};

}  // namespace multiverso
#endif
