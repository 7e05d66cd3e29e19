// LogisticRegression sample readers (see include/multiverso/apps/app_api.h).
// Reference behaviour: Applications/LogisticRegression/src/reader.cpp:17-438 -- SampleReader
// (libsvm / dense text), WeightedSampleReader ("label:weight"), BSparseSampleReader (binary),
// background thread + bounded ring of parsed samples, bias feature appended.
#include <atomic>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "multiverso/apps/app_api.h"
#include "multiverso/io/io.h"
#include "multiverso/util/log.h"

namespace {

struct Sample {
  float label = 0.f, weight = 1.f;
  std::vector<int64_t> keys;
  std::vector<float> vals;
};

enum class Kind { Default, Weight, BSparse };

class Reader {
 public:
  Reader(const std::string& files, Kind kind, bool sparse, int64_t input_size, int cap)
      : kind_(kind), sparse_(sparse), input_size_(input_size), cap_(cap > 0 ? cap : 4096) {
    std::stringstream ss(files);
    std::string f;
    while (std::getline(ss, f, ';'))
      if (!f.empty()) files_.push_back(f);
    Start();
  }
  ~Reader() { Stop(); }
  void Reset() {
    Stop();
    queue_.clear();
    eof_ = false;
    Start();
  }
  // pops up to max samples; returns 0 only when the epoch is exhausted, and -(nnz of the next
  // sample) when not even one sample fits into max_nnz (the caller grows its buffers and retries)
  int64_t Next(int64_t max_samples, int64_t max_nnz, int64_t* row_ptr, int64_t* keys, float* vals,
               float* labels, float* weights) {
    int64_t n = 0, nnz = 0;
    row_ptr[0] = 0;
    while (n < max_samples) {
      std::unique_lock<std::mutex> lk(mu_);
      cv_.wait(lk, [&] { return !queue_.empty() || eof_; });
      if (queue_.empty()) break;
      Sample& s = queue_.front();
      if (nnz + static_cast<int64_t>(s.keys.size()) > max_nnz) {
        if (n == 0) return -static_cast<int64_t>(s.keys.size());
        break;
      }
      std::copy(s.keys.begin(), s.keys.end(), keys + nnz);
      std::copy(s.vals.begin(), s.vals.end(), vals + nnz);
      nnz += static_cast<int64_t>(s.keys.size());
      labels[n] = s.label;
      if (weights) weights[n] = s.weight;
      row_ptr[++n] = nnz;
      queue_.pop_front();
      lk.unlock();
      cv_.notify_all();
    }
    return n;
  }

 private:
  void Start() {
    stop_ = false;
    thread_ = std::thread([this] { Main(); });
  }
  void Stop() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
    }
    cv_.notify_all();
    if (thread_.joinable()) thread_.join();
  }
  void Push(Sample&& s) {
    // bias feature: the reference does input_size += 1 and feeds a constant 1
    s.keys.push_back(input_size_);
    s.vals.push_back(1.f);
    std::unique_lock<std::mutex> lk(mu_);
    cv_.wait(lk, [&] { return static_cast<int>(queue_.size()) < cap_ || stop_; });
    if (stop_) return;
    queue_.push_back(std::move(s));
    lk.unlock();
    cv_.notify_all();
  }
  void ParseTextLine(const std::string& line) {
    const char* p = line.c_str();
    char* end = nullptr;
    Sample s;
    s.label = strtof(p, &end);
    if (end == p) return;
    p = end;
    if (kind_ == Kind::Weight && *p == ':') {
      s.weight = strtof(p + 1, &end);
      p = end;
    }
    if (sparse_) {
      for (;;) {
        while (*p == ' ' || *p == '\t') ++p;
        if (!*p) break;
        long long k = strtoll(p, &end, 10);
        if (end == p) break;
        p = end;
        float v = 1.f;
        if (*p == ':') {
          v = strtof(p + 1, &end);
          p = end;
        }
        if (k >= 0 && k < input_size_) {
          s.keys.push_back(k);
          s.vals.push_back(v);
        }
      }
    } else {
      for (int64_t k = 0; k < input_size_; ++k) {
        float v = strtof(p, &end);
        if (end == p) break;
        p = end;
        if (v != 0.f) {
          s.keys.push_back(k);
          s.vals.push_back(v);
        }
      }
    }
    Push(std::move(s));
  }
  void ParseBinary(const std::string& path) {
    // records: u64 count | i32 label | f64 weight | count x u64 keys (values are 1)
    FILE* fp = fopen(path.c_str(), "rb");
    if (!fp) return;
    for (;;) {
      uint64_t cnt;
      int32_t label;
      double weight;
      if (fread(&cnt, 8, 1, fp) != 1 || fread(&label, 4, 1, fp) != 1 || fread(&weight, 8, 1, fp) != 1) break;
      Sample s;
      s.label = static_cast<float>(label);
      s.weight = static_cast<float>(weight);
      std::vector<uint64_t> ks(cnt);
      if (cnt && fread(ks.data(), 8, cnt, fp) != cnt) break;
      for (uint64_t k : ks)
        if (static_cast<int64_t>(k) < input_size_) {
          s.keys.push_back(static_cast<int64_t>(k));
          s.vals.push_back(1.f);
        }
      Push(std::move(s));
      if (stop_) break;
    }
    fclose(fp);
  }
  void Main() {
    for (const std::string& f : files_) {
      if (stop_) break;
      if (kind_ == Kind::BSparse) {
        ParseBinary(f);
      } else {
        multiverso::TextReader rd(multiverso::URI(f), 1 << 20);
        if (!rd.Good()) {
          multiverso::Log::Error("LogReg reader: cannot open %s", f.c_str());
          continue;
        }
        std::string line;
        while (!stop_ && rd.GetLine(line))
          if (!line.empty()) ParseTextLine(line);
      }
    }
    {
      std::lock_guard<std::mutex> lk(mu_);
      eof_ = true;
    }
    cv_.notify_all();
  }

  Kind kind_;
  bool sparse_;
  int64_t input_size_;
  int cap_;
  std::vector<std::string> files_;
  std::deque<Sample> queue_;
  std::mutex mu_;
  std::condition_variable cv_;
  std::thread thread_;
  bool stop_ = false, eof_ = false;
};

}  // namespace

extern "C" {

void* MVA_LRReaderOpen(const char* files, const char* reader_type, int sparse, int64_t input_size,
                       int buffer_samples) {
  std::string t = reader_type ? reader_type : "default";
  Kind k = t == "weight" ? Kind::Weight : (t == "bsparse" ? Kind::BSparse : Kind::Default);
  return new Reader(files, k, sparse != 0, input_size, buffer_samples);
}
int64_t MVA_LRReaderNext(void* reader, int64_t max_samples, int64_t max_nnz, int64_t* row_ptr,
                         int64_t* keys, float* vals, float* labels, float* weights) {
  return static_cast<Reader*>(reader)->Next(max_samples, max_nnz, row_ptr, keys, vals, labels, weights);
}
void MVA_LRReaderReset(void* reader) { static_cast<Reader*>(reader)->Reset(); }
void MVA_LRReaderClose(void* reader) { delete static_cast<Reader*>(reader); }

}  // extern "C"
