#include "multiverso/util/parallel_for.h"

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "multiverso/util/configure.h"

namespace multiverso {

namespace {

// One ParallelFor call. Work is handed out in small grains from a shared cursor (dynamic
// scheduling): a helper that is slow to wake up -- or a vCPU that is not running at all -- simply
// takes fewer grains, and a helper that wakes up after everything is done never makes the caller
// wait (completion is counted in items, not in helpers). The state outlives late helpers through
// the shared_ptr; `body` is only dereferenced while grains remain, i.e. before the caller returns.
struct Loop {
  const std::function<void(int64_t, int64_t)>* body;
  int64_t n, grain;
  std::atomic<int64_t> cursor{0};
  std::atomic<int64_t> done{0};
  std::mutex mu;
  std::condition_variable cv;

  void Work() {
    for (;;) {
      const int64_t b = cursor.fetch_add(grain, std::memory_order_relaxed);
      if (b >= n) return;
      const int64_t e = std::min(n, b + grain);
      (*body)(b, e);
      if (done.fetch_add(e - b, std::memory_order_acq_rel) + (e - b) == n) {
        std::lock_guard<std::mutex> lk(mu);
        cv.notify_all();
      }
    }
  }
  void WaitAll() {
    if (done.load(std::memory_order_acquire) == n) return;
    std::unique_lock<std::mutex> lk(mu);
    cv.wait(lk, [&] { return done.load(std::memory_order_acquire) == n; });
  }
};

// Idle pool threads block on a condition variable (no spinning: the caller, the worker actor and
// the server actor of one process all use this pool right after one another).
class Pool {
 public:
  static Pool& Get() {
    static Pool* pool = new Pool();   // leaked on purpose: workers may outlive static destruction
    return *pool;
  }
  int capacity() const { return static_cast<int>(threads_.size()) + 1; }
  void Enlist(const std::shared_ptr<Loop>& loop, int helpers) {
    {
      std::lock_guard<std::mutex> lk(mu_);
      for (int i = 0; i < helpers; ++i) queue_.push_back(loop);
    }
    if (helpers == 1) cv_.notify_one();
    else cv_.notify_all();
  }

 private:
  Pool() {
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    for (unsigned i = 1; i < hw; ++i) threads_.emplace_back([this] { Main(); });
    for (auto& t : threads_) t.detach();
  }
  void Main() {
    for (;;) {
      std::shared_ptr<Loop> loop;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return !queue_.empty(); });
        loop = std::move(queue_.front());
        queue_.pop_front();
      }
      loop->Work();
    }
  }
  std::mutex mu_;
  std::condition_variable cv_;
  std::deque<std::shared_ptr<Loop>> queue_;
  std::vector<std::thread> threads_;
};

}  // namespace

int ParallelForCapacity() { return Pool::Get().capacity(); }

void ParallelFor(int64_t n, int threads, const std::function<void(int64_t, int64_t)>& body) {
  if (n <= 0) return;
  Pool& pool = Pool::Get();
  const int width = static_cast<int>(std::min<int64_t>(std::min(threads, pool.capacity()), n));
  if (width <= 1) {
    body(0, n);
    return;
  }
  auto loop = std::make_shared<Loop>();
  loop->body = &body;
  loop->n = n;
  loop->grain = std::max<int64_t>(1, n / (static_cast<int64_t>(width) * 8));   // ~8 grains per participant
  pool.Enlist(loop, width - 1);
  loop->Work();        // the caller is a participant
  loop->WaitAll();
}

MV_DECLARE_int(omp_threads);

void ParallelMemcpy(void* dst, const void* src, size_t bytes) {
  constexpr size_t kParallelBytes = 8u << 20;
  if (bytes < kParallelBytes) {
    if (bytes) std::memcpy(dst, src, bytes);
    return;
  }
  constexpr int64_t kPiece = 256 << 10;   // copy in 256 KiB pieces; ParallelFor groups them into grains
  const int64_t pieces = static_cast<int64_t>((bytes + kPiece - 1) / kPiece);
  ParallelFor(pieces, std::max(1, MV_CONFIG(omp_threads)), [=](int64_t lo, int64_t hi) {
    const size_t b = static_cast<size_t>(lo) * kPiece;
    const size_t e = std::min(bytes, static_cast<size_t>(hi) * kPiece);
    std::memcpy(static_cast<char*>(dst) + b, static_cast<const char*>(src) + b, e - b);
  });
}

}  // namespace multiverso
