// Configuration of the native LogisticRegression application: "key=value" lines, the 26 keys
// and defaults of the reference (Applications/LogisticRegression/src/configure.h:19-97,
// configure.cpp:32-82). '#' starts a comment; unknown keys are reported and ignored.
#ifndef MVAPP_LOGREG_CONFIGURE_H_
#define MVAPP_LOGREG_CONFIGURE_H_
#include <cstdint>
#include <string>

namespace logreg {

struct Configure {
  int64_t input_size = 0;          // number of features (the bias column is added on top)
  int output_size = 1;             // 1 = binary / regression, > 1 = classes
  bool sparse = false;             // libsvm-style input, sparse gradients / sparse PS tables
  int train_epoch = 1;
  int minibatch_size = 20;
  int read_buffer_size = 2048;     // samples buffered by the async reader
  double regular_coef = 0.0005;
  double learning_rate = 0.8;
  double learning_rate_coef = 1e6;
  double alpha = 0.005, beta = 1.0, lambda1 = 15.0, lambda2 = 0.0;   // FTRL
  std::string init_model_file;
  std::string train_file;          // ';'-separated list
  std::string reader_type = "default";      // default | weight | bsparse
  std::string test_file;
  std::string output_model_file = "logreg.model";
  std::string output_file = "logreg.output";
  bool use_ps = false;
  bool pipeline = true;            // double-buffered pulls (use_ps only)
  int sync_frequency = 1;          // pull every this many minibatches (use_ps only)
  std::string updater_type = "default";     // default | sgd | ftrl
  std::string objective_type = "default";   // default (linear) | sigmoid | softmax | ftrl
  std::string regular_type = "default";     // default (none) | L1 | L2
  int64_t show_time_per_sample = 10000;

  // false (with a message on stderr) when the file cannot be read or input_size is missing
  bool Load(const std::string& path);
  bool Set(const std::string& key, const std::string& value);
  bool ftrl() const { return objective_type == "ftrl" || updater_type == "ftrl"; }
};

}  // namespace logreg
#endif
