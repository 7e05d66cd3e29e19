// Flag registry: MV_DEFINE_* / MV_CONFIG_* / ParseCMDFlags / SetCMDFlag.
// Counterpart of include/multiverso/util/configure.h:13-115; one type-erased registry
// (std::variant) instead of one singleton per type; fixes SURVEY Q10 (args containing '-'
// but no '=' are left alone; doubles are parsed from the value).
#ifndef MULTIVERSO_UTIL_CONFIGURE_H_
#define MULTIVERSO_UTIL_CONFIGURE_H_
#include <map>
#include <mutex>
#include <string>
#include <variant>

namespace multiverso {
namespace config {

using Value = std::variant<int, bool, double, std::string>;

struct Entry {
  Value value;
  std::string text;
};

class Registry {
 public:
  static Registry& Get();
  // Returns a stable pointer to the stored value (entries are never erased).
  template <typename T>
  T* Define(const std::string& name, const T& dflt, const std::string& text) {
    std::lock_guard<std::mutex> lk(mu_);
    auto it = entries_.find(name);
    if (it == entries_.end()) it = entries_.emplace(name, Entry{Value(dflt), text}).first;
    return std::get_if<T>(&it->second.value);
  }
  bool Has(const std::string& name);
  template <typename T>
  bool Set(const std::string& name, const T& v) {
    std::lock_guard<std::mutex> lk(mu_);
    auto it = entries_.find(name);
    if (it == entries_.end()) return false;
    if (auto p = std::get_if<T>(&it->second.value)) {
      *p = v;
      return true;
    }
    return false;
  }
  // Parse "value" according to the flag's declared type. false if unknown flag / bad value.
  bool SetFromString(const std::string& name, const std::string& value);
  void PrintHelp();

 private:
  std::mutex mu_;
  std::map<std::string, Entry> entries_;
};

}  // namespace config

// Consumes recognised "-key=value" arguments and compacts argv (configure.cpp:9-54).
void ParseCMDFlags(int* argc, char* argv[]);

template <typename T>
void SetCMDFlag(const std::string& name, const T& value);

#define MV_DEFINE_FLAG_(type, name, dflt, text) \
  type* MV_FLAG_PTR_##name = ::multiverso::config::Registry::Get().Define<type>(#name, dflt, text)
#define MV_DECLARE_FLAG_(type, name) extern type* MV_FLAG_PTR_##name
#define MV_DEFINE_int(name, dflt, text) MV_DEFINE_FLAG_(int, name, dflt, text)
#define MV_DEFINE_bool(name, dflt, text) MV_DEFINE_FLAG_(bool, name, dflt, text)
#define MV_DEFINE_double(name, dflt, text) MV_DEFINE_FLAG_(double, name, dflt, text)
#define MV_DEFINE_string(name, dflt, text) MV_DEFINE_FLAG_(std::string, name, std::string(dflt), text)
#define MV_DECLARE_int(name) MV_DECLARE_FLAG_(int, name)
#define MV_DECLARE_bool(name) MV_DECLARE_FLAG_(bool, name)
#define MV_DECLARE_double(name) MV_DECLARE_FLAG_(double, name)
#define MV_DECLARE_string(name) MV_DECLARE_FLAG_(std::string, name)
#define MV_CONFIG(name) (*MV_FLAG_PTR_##name)

}  // namespace multiverso
#endif
