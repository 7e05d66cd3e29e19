// Dashboard / Monitor: named cumulative timers (include/multiverso/dashboard.h:16-76).
// Thread-safe per monitor here (the reference's are plain doubles).
#ifndef MULTIVERSO_DASHBOARD_H_
#define MULTIVERSO_DASHBOARD_H_
#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include "multiverso/util/timer.h"

namespace multiverso {

class Monitor {
 public:
  explicit Monitor(const std::string& name);
  void Begin() { timer_.Start(); }
  void End() { Add(timer_.elapse()); }
  void Add(double ms) {
    std::lock_guard<std::mutex> lk(mu_);
    elapse_ += ms;
    ++count_;
  }
  // Readers take the same lock as Add(): monitors are read (Display) while actors still update them.
  double average() const {
    std::lock_guard<std::mutex> lk(mu_);
    return count_ ? elapse_ / count_ : 0.0;
  }
  const std::string& name() const { return name_; }
  double elapse() const {
    std::lock_guard<std::mutex> lk(mu_);
    return elapse_;
  }
  long long count() const {
    std::lock_guard<std::mutex> lk(mu_);
    return count_;
  }
  std::string info_string() const;

 private:
  std::string name_;
  Timer timer_;
  mutable std::mutex mu_;
  double elapse_ = 0.0;
  long long count_ = 0;
};

class Dashboard {
 public:
  static void AddMonitor(const std::string& name, Monitor* m);
  static std::string Watch(const std::string& name);
  static void Display();
  static void Reset();

 private:
  static std::map<std::string, Monitor*>& record();
  static std::mutex& mu();
};

// Function-static monitor + a scope-local timer (safe when several threads run the scope).
#define MONITOR_BEGIN(name)                               \
  static ::multiverso::Monitor g_##name##_monitor(#name); \
  ::multiverso::Timer g_##name##_timer;
#define MONITOR_END(name) g_##name##_monitor.Add(g_##name##_timer.elapse());

}  // namespace multiverso
#endif
