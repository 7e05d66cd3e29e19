// WordEmbedding data pipeline: Dictionary, corpus Reader with stop words / sub-sampling,
// Huffman encoder, word_count tool (see include/multiverso/apps/app_api.h).
// Reference behaviour: Applications/WordEmbedding/src/dictionary.cpp:26-190, reader.cpp:37-97,
// util.cpp:116-146 (Sampler), huffman_encoder.cpp:87-196, preprocess/word_count.cpp:30-46.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <queue>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "multiverso/apps/app_api.h"
#include "multiverso/io/io.h"
#include "multiverso/util/log.h"

namespace {

constexpr int kMaxSentenceLength = 1000;   // constant.h:27

struct Dictionary {
  std::vector<std::string> words;
  std::vector<int64_t> freq;
  std::unordered_map<std::string, int> index;
  int64_t total = 0;
  void Insert(const std::string& w, int64_t f) {
    auto it = index.find(w);
    if (it == index.end()) {
      index.emplace(w, static_cast<int>(words.size()));
      words.push_back(w);
      freq.push_back(f);
    } else {
      freq[it->second] += f;
    }
    total += f;
  }
};

// Word tokenizer over a buffered text stream: whitespace separates words, '\n' yields "</s>".
class WordStream {
 public:
  explicit WordStream(const std::string& path) : fp_(fopen(path.c_str(), "rb")) {}
  ~WordStream() { if (fp_) fclose(fp_); }
  bool good() const { return fp_ != nullptr; }
  void rewind_() { if (fp_) { fseek(fp_, 0, SEEK_SET); pos_ = len_ = 0; pending_eol_ = false; } }
  // returns false at EOF; eol=true means a sentence boundary (no word)
  bool Next(std::string* word, bool* eol) {
    word->clear();
    *eol = false;
    if (pending_eol_) { pending_eol_ = false; *eol = true; return true; }
    for (;;) {
      int c = Get();
      if (c < 0) return !word->empty();
      if (c == ' ' || c == '\t' || c == '\r' || c == '\n') {
        if (!word->empty()) { if (c == '\n') pending_eol_ = true; return true; }
        if (c == '\n') { *eol = true; return true; }
        continue;
      }
      if (word->size() < 100) word->push_back(static_cast<char>(c));
    }
  }

 private:
  int Get() {
    if (pos_ >= len_) {
      if (!fp_) return -1;
      len_ = fread(buf_, 1, sizeof buf_, fp_);
      pos_ = 0;
      if (len_ == 0) return -1;
    }
    return static_cast<unsigned char>(buf_[pos_++]);
  }
  FILE* fp_;
  char buf_[1 << 16];
  size_t pos_ = 0, len_ = 0;
  bool pending_eol_ = false;
};

struct Corpus {
  Dictionary* dict;
  WordStream stream;
  std::unordered_set<std::string> stopwords;
  double sample;
  uint64_t rng;
  int sentence_len = 0;
  bool need_break = false;
  Corpus(Dictionary* d, const std::string& path, double s, uint64_t seed)
      : dict(d), stream(path), sample(s), rng(seed ? seed : 1) {}
  double NextUniform() {   // Sampler LCG (util.cpp:144-146)
    rng = rng * 25214903917ull + 11ull;
    return static_cast<double>((rng >> 16) & 0xFFFFFF) / 16777216.0;
  }
};

}  // namespace

extern "C" {

void* MVA_DictLoad(const char* vocab_file, int min_count) {
  std::ifstream in(vocab_file);
  if (!in) {
    multiverso::Log::Error("cannot open vocabulary file %s", vocab_file);
    return nullptr;
  }
  auto* d = new Dictionary();
  std::string w;
  long long f;
  while (in >> w >> f)
    if (f >= min_count) d->Insert(w, f);
  return d;
}

void* MVA_DictFromCorpus(const char* train_file, int min_count) {
  WordStream ws(train_file);
  if (!ws.good()) return nullptr;
  std::unordered_map<std::string, int64_t> counts;
  std::string w;
  bool eol;
  while (ws.Next(&w, &eol))
    if (!eol) ++counts[w];
  std::vector<std::pair<std::string, int64_t>> v(counts.begin(), counts.end());
  std::sort(v.begin(), v.end(), [](auto& a, auto& b) { return a.second != b.second ? a.second > b.second : a.first < b.first; });
  auto* d = new Dictionary();
  for (auto& kv : v)
    if (kv.second >= min_count) d->Insert(kv.first, kv.second);
  return d;
}

int MVA_DictSize(void* dict) { return static_cast<int>(static_cast<Dictionary*>(dict)->words.size()); }
int64_t MVA_DictTotalWords(void* dict) { return static_cast<Dictionary*>(dict)->total; }
void MVA_DictCounts(void* dict, int64_t* out) {
  auto* d = static_cast<Dictionary*>(dict);
  std::copy(d->freq.begin(), d->freq.end(), out);
}
const char* MVA_DictWord(void* dict, int id) {
  auto* d = static_cast<Dictionary*>(dict);
  return (id >= 0 && id < static_cast<int>(d->words.size())) ? d->words[id].c_str() : "";
}
int MVA_DictIndex(void* dict, const char* word) {
  auto* d = static_cast<Dictionary*>(dict);
  auto it = d->index.find(word);
  return it == d->index.end() ? -1 : it->second;
}
void MVA_DictFree(void* dict) { delete static_cast<Dictionary*>(dict); }

int64_t MVA_WordCount(const char* train_file, const char* out_vocab_file, int min_count) {
  void* d = MVA_DictFromCorpus(train_file, min_count);
  if (!d) return -1;
  auto* dict = static_cast<Dictionary*>(d);
  FILE* out = fopen(out_vocab_file, "w");
  if (!out) { delete dict; return -1; }
  for (size_t i = 0; i < dict->words.size(); ++i)
    fprintf(out, "%s %lld\n", dict->words[i].c_str(), static_cast<long long>(dict->freq[i]));
  fclose(out);
  int64_t n = static_cast<int64_t>(dict->words.size());
  delete dict;
  return n;
}

void* MVA_CorpusOpen(void* dict, const char* train_file, const char* stopword_file, double sample,
                     uint64_t seed) {
  auto* c = new Corpus(static_cast<Dictionary*>(dict), train_file, sample, seed);
  if (!c->stream.good()) {
    multiverso::Log::Error("cannot open corpus %s", train_file);
    delete c;
    return nullptr;
  }
  if (stopword_file && *stopword_file) {
    std::ifstream in(stopword_file);
    std::string w;
    while (in >> w) c->stopwords.insert(w);
  }
  return c;
}

int64_t MVA_CorpusNextBlock(void* corpus, int32_t* out, int64_t max_tokens, int64_t* words_read) {
  auto* c = static_cast<Corpus*>(corpus);
  int64_t n = 0, words = 0;
  std::string w;
  bool eol;
  const double train_words = static_cast<double>(c->dict->total);
  while (n < max_tokens - 1) {
    if (c->need_break) {
      if (n > 0 && out[n - 1] != -1) out[n++] = -1;
      c->need_break = false;
      c->sentence_len = 0;
      continue;
    }
    if (!c->stream.Next(&w, &eol)) break;
    if (eol) { c->need_break = c->sentence_len > 0; continue; }
    ++words;
    if (!c->stopwords.empty() && c->stopwords.count(w)) continue;
    auto it = c->dict->index.find(w);
    if (it == c->dict->index.end()) continue;            // OOV words are skipped (reader.cpp:60-66)
    if (c->sample > 0) {
      // word2vec sub-sampling (util.cpp:137-142): keep with p = (sqrt(f/(s*T)) + 1) * (s*T)/f
      const double f = static_cast<double>(c->dict->freq[it->second]);
      const double st = c->sample * train_words;
      const double keep = (std::sqrt(f / st) + 1.0) * st / f;
      if (keep < c->NextUniform()) continue;
    }
    out[n++] = it->second;
    if (++c->sentence_len >= kMaxSentenceLength) c->need_break = true;
  }
  if (words_read) *words_read = words;
  return n;
}

void MVA_CorpusReset(void* corpus) {
  auto* c = static_cast<Corpus*>(corpus);
  c->stream.rewind_();
  c->sentence_len = 0;
  c->need_break = false;
}
void MVA_CorpusClose(void* corpus) { delete static_cast<Corpus*>(corpus); }

int MVA_HuffmanBuild(const int64_t* freq, int n, int max_code, int32_t* points, int8_t* codes, int32_t* lens) {
  if (n <= 0) return -1;
  // classic two-queue construction over nodes [0,n) leaves and [n, 2n-1) inner nodes
  using Item = std::pair<int64_t, int>;
  std::priority_queue<Item, std::vector<Item>, std::greater<Item>> heap;
  for (int i = 0; i < n; ++i) heap.emplace(freq[i], i);
  std::vector<int> parent(2 * static_cast<size_t>(n), -1);
  std::vector<int8_t> branch(2 * static_cast<size_t>(n), 0);
  int next = n;
  while (heap.size() > 1) {
    Item a = heap.top(); heap.pop();
    Item b = heap.top(); heap.pop();
    parent[a.second] = next;
    parent[b.second] = next;
    branch[b.second] = 1;
    heap.emplace(a.first + b.first, next);
    ++next;
  }
  const int root = next - 1;
  int longest = 0;
  std::vector<int> pts;
  std::vector<int8_t> cds;
  for (int w = 0; w < n; ++w) {
    pts.clear();
    cds.clear();
    for (int node = w; node != root && parent[node] >= 0; node = parent[node]) {
      cds.push_back(branch[node]);
      pts.push_back(parent[node] - n);    // inner-node id = row of the output table
    }
    const int len = static_cast<int>(cds.size());
    if (len > max_code) return -1;
    lens[w] = len;
    for (int d = 0; d < len; ++d) {        // root-first order
      points[static_cast<size_t>(w) * max_code + d] = pts[len - 1 - d];
      codes[static_cast<size_t>(w) * max_code + d] = cds[len - 1 - d];
    }
    longest = std::max(longest, len);
  }
  return longest;
}

}  // extern "C"
