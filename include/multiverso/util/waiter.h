// Waiter: counting latch (include/multiverso/util/waiter.h:9-33).
#ifndef MULTIVERSO_UTIL_WAITER_H_
#define MULTIVERSO_UTIL_WAITER_H_
#include <condition_variable>
#include <mutex>

namespace multiverso {

class Waiter {
 public:
  explicit Waiter(int num_wait = 1) : num_wait_(num_wait) {}
  void Wait() {
    std::unique_lock<std::mutex> lk(mu_);
    cv_.wait(lk, [&] { return num_wait_ <= 0; });
  }
  void Notify() {
    std::lock_guard<std::mutex> lk(mu_);
    if (--num_wait_ <= 0) cv_.notify_all();
  }
  void Reset(int num_wait) {
    std::lock_guard<std::mutex> lk(mu_);
    num_wait_ = num_wait;
    if (num_wait_ <= 0) cv_.notify_all();
  }
  bool Done() {
    std::lock_guard<std::mutex> lk(mu_);
    return num_wait_ <= 0;
  }

 private:
  std::mutex mu_;
  std::condition_variable cv_;
  int num_wait_;
};

}  // namespace multiverso
#endif
