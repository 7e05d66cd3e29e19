#include "configure.h"

#include <cstdio>
#include <cstdlib>
#include <functional>
#include <map>

#include "multiverso/io/io.h"

namespace logreg {

namespace {
std::string Trim(const std::string& s) {
  const size_t b = s.find_first_not_of(" \t\r\n");
  if (b == std::string::npos) return "";
  return s.substr(b, s.find_last_not_of(" \t\r\n") - b + 1);
}
bool ToBool(const std::string& v) { return v == "1" || v == "true" || v == "True" || v == "TRUE" || v == "yes"; }
}  // namespace

bool Configure::Set(const std::string& key, const std::string& value) {
  using Setter = std::function<void(const std::string&)>;
  auto i32 = [](int* p) { return Setter([p](const std::string& v) { *p = static_cast<int>(atof(v.c_str())); }); };
  auto i64 = [](int64_t* p) { return Setter([p](const std::string& v) { *p = static_cast<int64_t>(atof(v.c_str())); }); };
  auto f64 = [](double* p) { return Setter([p](const std::string& v) { *p = atof(v.c_str()); }); };
  auto str = [](std::string* p) { return Setter([p](const std::string& v) { *p = v; }); };
  auto flg = [](bool* p) { return Setter([p](const std::string& v) { *p = ToBool(v); }); };
  const std::map<std::string, Setter> table = {
      {"input_size", i64(&input_size)},
      {"output_size", i32(&output_size)},
      {"sparse", flg(&sparse)},
      {"train_epoch", i32(&train_epoch)},
      {"minibatch_size", i32(&minibatch_size)},
      {"read_buffer_size", i32(&read_buffer_size)},
      {"regular_coef", f64(&regular_coef)},
      {"learning_rate", f64(&learning_rate)},
      {"learning_rate_coef", f64(&learning_rate_coef)},
      {"alpha", f64(&alpha)},
      {"beta", f64(&beta)},
      {"lambda1", f64(&lambda1)},
      {"lambda2", f64(&lambda2)},
      {"init_model_file", str(&init_model_file)},
      {"train_file", str(&train_file)},
      {"reader_type", str(&reader_type)},
      {"test_file", str(&test_file)},
      {"output_model_file", str(&output_model_file)},
      {"output_file", str(&output_file)},
      {"use_ps", flg(&use_ps)},
      {"pipeline", flg(&pipeline)},
      {"sync_frequency", i32(&sync_frequency)},
      {"updater_type", str(&updater_type)},
      {"objective_type", str(&objective_type)},
      {"regular_type", str(&regular_type)},
      {"show_time_per_sample", i64(&show_time_per_sample)},
  };
  auto it = table.find(key);
  if (it == table.end()) return false;
  it->second(value);
  return true;
}

bool Configure::Load(const std::string& path) {
  multiverso::TextReader reader(multiverso::URI(path), 1 << 16);
  if (!reader.Good()) {
    fprintf(stderr, "logreg: cannot open the config file %s\n", path.c_str());
    return false;
  }
  std::string line;
  while (reader.GetLine(line)) {
    line = Trim(line.substr(0, line.find('#')));
    const size_t eq = line.find('=');
    if (line.empty() || eq == std::string::npos) continue;
    const std::string key = Trim(line.substr(0, eq)), value = Trim(line.substr(eq + 1));
    if (!Set(key, value)) fprintf(stderr, "logreg: unknown config key '%s' ignored\n", key.c_str());
  }
  if (input_size <= 0) {
    fprintf(stderr, "logreg: input_size must be set\n");
    return false;
  }
  if (output_size < 1) output_size = 1;
  if (minibatch_size < 1) minibatch_size = 1;
  if (sync_frequency < 1) sync_frequency = 1;
  return true;
}

}  // namespace logreg
