// word_count -- vocabulary preprocessing tool of the wordembedding application (reference:
// Applications/WordEmbedding/preprocess/word_count.cpp:30-46): counts the words of a corpus
// and writes "word freq" lines sorted by decreasing frequency, the format -read_vocab expects.
//
//   build/bin/word_count -train_file corpus.txt -save_vocab vocab.txt [-min_count 5]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "multiverso/apps/app_api.h"

int main(int argc, char* argv[]) {
  std::string train_file, vocab_file;
  int min_count = 0;
  for (int i = 1; i + 1 < argc; i += 2) {
    if (strcmp(argv[i], "-train_file") == 0) train_file = argv[i + 1];
    else if (strcmp(argv[i], "-save_vocab") == 0 || strcmp(argv[i], "-save_vocab_file") == 0) vocab_file = argv[i + 1];
    else if (strcmp(argv[i], "-min_count") == 0) min_count = atoi(argv[i + 1]);
  }
  if (train_file.empty() || vocab_file.empty()) {
    puts("usage: word_count -train_file <corpus> -save_vocab <vocab> [-min_count <int>]");
    return 2;
  }
  const long long n = MVA_WordCount(train_file.c_str(), vocab_file.c_str(), min_count);
  if (n < 0) {
    fprintf(stderr, "word_count: cannot read %s or write %s\n", train_file.c_str(), vocab_file.c_str());
    return 1;
  }
  printf("%lld words written to %s\n", n, vocab_file.c_str());
  return 0;
}
