// MtQueue<T>: mutex + condvar MPMC queue with move semantics (counterpart of
// include/multiverso/util/mt_queue.h:18-147; built on std::deque).
#ifndef MULTIVERSO_UTIL_MT_QUEUE_H_
#define MULTIVERSO_UTIL_MT_QUEUE_H_
#include <condition_variable>
#include <deque>
#include <mutex>

namespace multiverso {

template <typename T>
class MtQueue {
 public:
  void Push(T item) {
    {
      std::lock_guard<std::mutex> lk(mu_);
      q_.push_back(std::move(item));
    }
    cv_.notify_one();
  }
  // Blocks until an item is available or Exit() was called. false => queue exited and empty.
  bool Pop(T& out) {
    std::unique_lock<std::mutex> lk(mu_);
    cv_.wait(lk, [&] { return !q_.empty() || exit_; });
    if (q_.empty()) return false;
    out = std::move(q_.front());
    q_.pop_front();
    return true;
  }
  bool TryPop(T& out) {
    std::lock_guard<std::mutex> lk(mu_);
    if (q_.empty()) return false;
    out = std::move(q_.front());
    q_.pop_front();
    return true;
  }
  // Copy of the head without removing it.
  bool Front(T& out) {
    std::unique_lock<std::mutex> lk(mu_);
    cv_.wait(lk, [&] { return !q_.empty() || exit_; });
    if (q_.empty()) return false;
    out = q_.front();
    return true;
  }
  size_t Size() const {
    std::lock_guard<std::mutex> lk(mu_);
    return q_.size();
  }
  bool Empty() const { return Size() == 0; }
  void Exit() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      exit_ = true;
    }
    cv_.notify_all();
  }
  bool Alive() const {
    std::lock_guard<std::mutex> lk(mu_);
    return !exit_;
  }

 private:
  mutable std::mutex mu_;
  std::condition_variable cv_;
  std::deque<T> q_;
  bool exit_ = false;
};

}  // namespace multiverso
#endif
