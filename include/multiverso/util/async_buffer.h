// ASyncBuffer<B>: double-buffer prefetcher (include/multiverso/util/async_buffer.h:10-116).
// A background thread runs fill(buffer); Get() returns the ready buffer and starts
// prefetching into the other one.
#ifndef MULTIVERSO_UTIL_ASYNC_BUFFER_H_
#define MULTIVERSO_UTIL_ASYNC_BUFFER_H_
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>

namespace multiverso {

template <typename BufferType>
class ASyncBuffer {
 public:
  ASyncBuffer(BufferType* b0, BufferType* b1, std::function<void(BufferType*)> fill)
      : fill_(std::move(fill)) {
    buf_[0] = b0;
    buf_[1] = b1;
    thread_ = std::thread([this] { Main(); });
    Kick(0);
  }
  ~ASyncBuffer() { Join(); }
  // Ready buffer; kicks the prefetch of the other buffer.
  BufferType* Get() {
    std::unique_lock<std::mutex> lk(mu_);
    cv_.wait(lk, [&] { return ready_; });
    ready_ = false;
    int cur = filling_;
    lk.unlock();
    Kick(1 - cur);
    return buf_[cur];
  }
  void Join() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      if (stop_) return;
      stop_ = true;
    }
    cv_.notify_all();
    if (thread_.joinable()) thread_.join();
  }

 private:
  void Kick(int which) {
    {
      std::lock_guard<std::mutex> lk(mu_);
      filling_ = which;
      request_ = true;
    }
    cv_.notify_all();
  }
  void Main() {
    for (;;) {
      std::unique_lock<std::mutex> lk(mu_);
      cv_.wait(lk, [&] { return request_ || stop_; });
      if (stop_) return;
      request_ = false;
      int which = filling_;
      lk.unlock();
      fill_(buf_[which]);
      lk.lock();
      ready_ = true;
      lk.unlock();
      cv_.notify_all();
    }
  }
  BufferType* buf_[2];
  std::function<void(BufferType*)> fill_;
  std::thread thread_;
  std::mutex mu_;
  std::condition_variable cv_;
  int filling_ = 0;
  bool request_ = false, ready_ = false, stop_ = false;
};

}  // namespace multiverso
#endif
