#include "option.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>

#include "multiverso/util/log.h"

namespace wordembedding {

bool Option::Parse(int argc, char* argv[]) {
  using Setter = std::function<void(const char*)>;
  auto flag = [](bool* dst) { return Setter([dst](const char* v) { *dst = atoi(v) != 0; }); };
  auto num = [](int* dst) { return Setter([dst](const char* v) { *dst = atoi(v); }); };
  auto big = [](int64_t* dst) { return Setter([dst](const char* v) { *dst = atoll(v); }); };
  auto str = [](std::string* dst) { return Setter([dst](const char* v) { *dst = v; }); };
  const std::map<std::string, Setter> table = {
      {"-size", num(&embeding_size)},
      {"-train_file", str(&train_file)},
      {"-endpoints_file", str(&endpoints_file)},
      {"-read_vocab", str(&read_vocab_file)},
      {"-binary", flag(&output_binary)},
      {"-cbow", flag(&cbow)},
      {"-alpha", [this](const char* v) { init_learning_rate = static_cast<float>(atof(v)); }},
      {"-output", str(&output_file)},
      {"-window", num(&window_size)},
      {"-sample", [this](const char* v) { sample = atof(v); }},
      {"-hs", flag(&hs)},
      {"-data_block_size", big(&data_block_size)},
      {"-max_preload_data_size", big(&max_preload_data_size)},
      {"-negative", num(&negative_num)},
      {"-threads", num(&thread_cnt)},
      {"-min_count", num(&min_count)},
      {"-epoch", num(&epoch)},
      {"-stopwords", flag(&stopwords)},
      {"-sw_file", str(&sw_file)},
      {"-use_adagrad", flag(&use_adagrad)},
      {"-is_pipeline", flag(&is_pipeline)},
  };
  for (int i = 1; i < argc; ++i) {
    auto it = table.find(argv[i]);
    if (it == table.end()) continue;
    if (i + 1 >= argc) {
      fprintf(stderr, "flag %s needs a value\n", argv[i]);
      return false;
    }
    it->second(argv[++i]);
  }
  if (hs) negative_num = 0;   // the two output layers are alternatives
  return true;
}

void Option::PrintUsage() {
  puts("usage: wordembedding -train_file <corpus> [-read_vocab <vocab>] -output <file> [flags] [-mvflag=value ...]");
  puts("  -size <int>                   embedding dimension (100)");
  puts("  -train_file <file>            training corpus, one sentence per line");
  puts("  -read_vocab <file>            vocabulary 'word freq' per line (word_count output); built from the corpus if absent");
  puts("  -output <file>                embeddings in word2vec format, written by rank 0");
  puts("  -binary <0|1>                 binary output (0)");
  puts("  -cbow <0|1>                   continuous bag of words (1) or skip-gram (0)");
  puts("  -alpha <float>                initial learning rate (0.025)");
  puts("  -window <int>                 maximal skip length between words (5)");
  puts("  -sample <float>               sub-sampling threshold of frequent words, 0 = off (0)");
  puts("  -hs <0|1>                     hierarchical softmax (0)");
  puts("  -negative <int>               negative samples per target (5)");
  puts("  -threads <int>                trainer threads (1)");
  puts("  -min_count <int>              drop words rarer than this (5)");
  puts("  -epoch <int>                  passes over the corpus (1)");
  puts("  -data_block_size <bytes>      corpus bytes per data block (1000000)");
  puts("  -max_preload_data_size <b>    bound of the loader queue (8000000000)");
  puts("  -stopwords <0|1> -sw_file <f> drop the words listed in the file");
  puts("  -use_adagrad <0|1>            AdaGrad with G^2 kept in two extra tables (0)");
  puts("  -is_pipeline <0|1>            prefetch the next block's parameters while training (1)");
  puts("  -endpoints_file <file>        machine list for the explicit-endpoint bootstrap");
}

void Option::Print() const {
  multiverso::Log::Info(
      "wordembedding: train_file=%s vocab=%s output=%s binary=%d size=%d %s %s window=%d sample=%g "
      "threads=%d min_count=%d epoch=%d alpha=%g block=%lld adagrad=%d pipeline=%d\n",
      train_file.c_str(), read_vocab_file.c_str(), output_file.c_str(), int(output_binary),
      embeding_size, cbow ? "cbow" : "skip-gram",
      hs ? "hs" : ("negative=" + std::to_string(negative_num)).c_str(), window_size, sample,
      thread_cnt, min_count, epoch, init_learning_rate, static_cast<long long>(data_block_size),
      int(use_adagrad), int(is_pipeline));
}

}  // namespace wordembedding
