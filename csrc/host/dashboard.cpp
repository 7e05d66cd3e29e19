#include "multiverso/dashboard.h"
#include <sstream>
#include "multiverso/util/log.h"

namespace multiverso {

Monitor::Monitor(const std::string& name) : name_(name) { Dashboard::AddMonitor(name, this); }

std::string Monitor::info_string() const {
  long long count;
  double elapse;
  {
    std::lock_guard<std::mutex> lk(mu_);
    count = count_;
    elapse = elapse_;
  }
  std::ostringstream ss;
  ss << "[Monitor] " << name_ << ": count = " << count << " elapse = " << elapse
     << "ms average = " << (count ? elapse / count : 0.0) << "ms";
  return ss.str();
}

std::map<std::string, Monitor*>& Dashboard::record() {
  static auto* r = new std::map<std::string, Monitor*>();
  return *r;
}
std::mutex& Dashboard::mu() {
  static auto* m = new std::mutex();
  return *m;
}
void Dashboard::AddMonitor(const std::string& name, Monitor* m) {
  std::lock_guard<std::mutex> lk(mu());
  record()[name] = m;
}
std::string Dashboard::Watch(const std::string& name) {
  std::lock_guard<std::mutex> lk(mu());
  auto it = record().find(name);
  return it == record().end() ? std::string("[Monitor] ") + name + ": not found" : it->second->info_string();
}
void Dashboard::Display() {
  std::lock_guard<std::mutex> lk(mu());
  Log::Info("--------------Show dashboard monitor information--------------");
  for (auto& kv : record()) Log::Info("%s", kv.second->info_string().c_str());
  Log::Info("---------------------------------------------------------------");
}
void Dashboard::Reset() {
  std::lock_guard<std::mutex> lk(mu());
  record().clear();
}

}  // namespace multiverso
