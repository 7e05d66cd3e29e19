// Timer: Start(), elapse() in milliseconds (include/multiverso/util/timer.h:9-24).
#ifndef MULTIVERSO_UTIL_TIMER_H_
#define MULTIVERSO_UTIL_TIMER_H_
#include <chrono>

namespace multiverso {
class Timer {
 public:
  Timer() { Start(); }
  void Start() { start_ = Clock::now(); }
  double elapse() const {
    return std::chrono::duration<double, std::milli>(Clock::now() - start_).count();
  }

 private:
  using Clock = std::chrono::steady_clock;
  Clock::time_point start_;
};
}  // namespace multiverso
#endif
