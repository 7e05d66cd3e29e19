// Command line of the native `wordembedding` application: the reference's 21 "-flag value"
// pairs with its defaults (Applications/WordEmbedding/src/util.cpp:6-56).
#ifndef MVAPP_WORDEMBEDDING_OPTION_H_
#define MVAPP_WORDEMBEDDING_OPTION_H_
#include <cstdint>
#include <string>

namespace wordembedding {

struct Option {
  std::string train_file, read_vocab_file, output_file, sw_file, endpoints_file;
  bool hs = false;
  bool output_binary = false;
  bool cbow = true;
  bool stopwords = false;
  bool use_adagrad = false;
  bool is_pipeline = true;
  double sample = 0;
  int64_t data_block_size = 1000000;            // corpus bytes per data block
  int64_t max_preload_data_size = 8000000000LL; // bound of the loader's block queue, bytes
  int embeding_size = 100;
  int thread_cnt = 1;
  int window_size = 5;
  int negative_num = 5;
  int min_count = 5;
  int epoch = 1;
  float init_learning_rate = 0.025f;
  int64_t total_words = 0;                      // filled from the dictionary

  // Consumes "-flag value" pairs; unknown arguments are ignored (the "-key=value" ones belong
  // to MV_Init). Returns false when a flag misses its value.
  bool Parse(int argc, char* argv[]);
  static void PrintUsage();
  void Print() const;
};

}  // namespace wordembedding
#endif
