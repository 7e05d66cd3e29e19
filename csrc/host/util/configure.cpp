// Flag registry implementation (see include/multiverso/util/configure.h).
#include "multiverso/util/configure.h"
#include <cstdlib>
#include <cstring>
#include "multiverso/util/log.h"

namespace multiverso {
namespace config {

Registry& Registry::Get() {
  static Registry* r = new Registry();   // leaked on purpose: flags outlive static dtors
  return *r;
}

bool Registry::Has(const std::string& name) {
  std::lock_guard<std::mutex> lk(mu_);
  return entries_.count(name) != 0;
}

bool Registry::SetFromString(const std::string& name, const std::string& value) {
  std::lock_guard<std::mutex> lk(mu_);
  auto it = entries_.find(name);
  if (it == entries_.end()) return false;
  Value& v = it->second.value;
  if (std::holds_alternative<std::string>(v)) {
    v = value;
    return true;
  }
  if (std::holds_alternative<bool>(v)) {
    if (value == "true" || value == "1" || value == "True" || value == "TRUE") v = true;
    else if (value == "false" || value == "0" || value == "False" || value == "FALSE") v = false;
    else return false;
    return true;
  }
  char* end = nullptr;
  if (std::holds_alternative<int>(v)) {
    long x = strtol(value.c_str(), &end, 10);
    if (end == value.c_str() || *end != '\0') return false;
    v = static_cast<int>(x);
    return true;
  }
  double d = strtod(value.c_str(), &end);   // parsed from the VALUE (reference: whole arg, Q10)
  if (end == value.c_str() || *end != '\0') return false;
  v = d;
  return true;
}

void Registry::PrintHelp() {
  std::lock_guard<std::mutex> lk(mu_);
  for (auto& kv : entries_) Log::Info("  -%s : %s", kv.first.c_str(), kv.second.text.c_str());
}

}  // namespace config

void ParseCMDFlags(int* argc, char* argv[]) {
  if (argc == nullptr || argv == nullptr) return;
  int kept = 0;
  for (int i = 0; i < *argc; ++i) {
    const char* arg = argv[i];
    bool consumed = false;
    if (arg != nullptr && arg[0] == '-') {
      const char* eq = strchr(arg, '=');
      if (eq != nullptr) {
        const char* key = arg + 1;
        if (*key == '-') ++key;
        std::string name(key, eq - key), value(eq + 1);
        consumed = config::Registry::Get().SetFromString(name, value);
      }
    }
    if (!consumed) argv[kept++] = argv[i];
  }
  *argc = kept;
}

template <typename T>
void SetCMDFlag(const std::string& name, const T& value) {
  if (!config::Registry::Get().Set<T>(name, value))
    Log::Fatal("SetCMDFlag: flag '%s' is not defined with that type", name.c_str());
}
template void SetCMDFlag<int>(const std::string&, const int&);
template void SetCMDFlag<bool>(const std::string&, const bool&);
template void SetCMDFlag<double>(const std::string&, const double&);
template void SetCMDFlag<std::string>(const std::string&, const std::string&);

}  // namespace multiverso
