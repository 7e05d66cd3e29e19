// App-defined tables of the LogisticRegression application: SparseTable<T> and FTRLTable<T>
// (counterpart of Applications/LogisticRegression/src/util/sparse_table.h:16-302 and
// util/ftrl_sparse_table.h:11-86). They double as the worked example of the extension point:
// subclass WorkerTable / ServerTable, implement Partition / ProcessReplyGet / ProcessAdd /
// ProcessGet, and declare an Option with DEFINE_TABLE_TYPE.
//
//   * keys are size_t feature ids in [0, size), range-partitioned over the servers;
//   * Add(keys, values): the server does storage[key] -= value (the client sends lr * grad);
//   * Get(keys): returns the current values of exactly those keys; Get() returns every key the
//     server has ever been updated with (properly compacted -- the reference's whole-table
//     branch overflows);
//   * FTRLTable stores {z, n} pairs per key.
#ifndef MULTIVERSO_TABLE_SPARSE_TABLE_H_
#define MULTIVERSO_TABLE_SPARSE_TABLE_H_
#include <algorithm>
#include <unordered_map>
#include <vector>
#include "multiverso/multiverso.h"
#include "multiverso/table/array_table.h"
#include "multiverso/table_interface.h"
#include "multiverso/util/log.h"

namespace multiverso {

template <typename T> class SparseWorkerTable;
template <typename T> class SparseServerTable;

template <typename T>
struct SparseTableOption {
  explicit SparseTableOption(size_t s) : size(s) {}
  size_t size;
  DEFINE_TABLE_TYPE(T, SparseWorkerTable, SparseServerTable);
};

template <typename T>
struct FTRLEntry {
  T z, n;
  FTRLEntry() : z(0), n(0) {}
  FTRLEntry(T z_, T n_) : z(z_), n(n_) {}
  FTRLEntry& operator-=(const FTRLEntry& o) { z -= o.z; n -= o.n; return *this; }
  bool operator!=(const FTRLEntry& o) const { return z != o.z || n != o.n; }
};

template <typename T>
class SparseWorkerTable : public WorkerTable {
 public:
  explicit SparseWorkerTable(size_t size) : size_(size), offsets_(RangeOffsets(size, MV_NumServers())) {}
  template <typename Opt>
  explicit SparseWorkerTable(const Opt& o) : SparseWorkerTable(static_cast<size_t>(o.size)) {}

  // Pull `n` keys into `values` (same order).
  void Get(const size_t* keys, size_t n, T* values) { Wait(GetAsync(keys, n, values)); }
  int GetAsync(const size_t* keys, size_t n, T* values) {
    const int id = NewRequest();
    {
      std::lock_guard<std::mutex> lk(rec_mu_);
      Record& r = rec_[id];
      for (size_t i = 0; i < n; ++i) r.dst[keys[i]] = values + i;
    }
    Submit(id, MsgType::Request_Get, {Blob(keys, n * sizeof(size_t))});
    return id;
  }
  // Pull every key the servers hold; out_keys / out_values are resized.
  void GetAll(std::vector<size_t>* out_keys, std::vector<T>* out_values) {
    const int id = NewRequest();
    {
      std::lock_guard<std::mutex> lk(rec_mu_);
      Record& r = rec_[id];
      r.all_keys = out_keys;
      r.all_vals = out_values;
      out_keys->clear();
      out_values->clear();
    }
    size_t whole = static_cast<size_t>(-1);
    Submit(id, MsgType::Request_Get, {Blob(&whole, sizeof whole)});
    Wait(id);
  }
  void Add(const size_t* keys, const T* values, size_t n, const AddOption* opt = nullptr) {
    Wait(AddAsync(keys, values, n, opt));
  }
  int AddAsync(const size_t* keys, const T* values, size_t n, const AddOption* opt = nullptr) {
    return WorkerTable::AddAsync(Blob(keys, n * sizeof(size_t)), Blob(values, n * sizeof(T)), opt);
  }

  int Partition(const std::vector<Blob>& kv, MsgType type,
                std::unordered_map<int, std::vector<Blob>>* out) override {
    const bool is_add = type == MsgType::Request_Add;
    const Blob& keys = kv[0];
    const size_t n = keys.size<size_t>();
    const int S = MV_NumServers();
    if (!is_add && n == 1 && keys.As<size_t>(0) == static_cast<size_t>(-1)) {
      for (int s = 0; s < S; ++s)
        if (offsets_[s + 1] > offsets_[s]) (*out)[s].push_back(keys);
      return static_cast<int>(out->size());
    }
    std::vector<std::vector<size_t>> bucket(S);
    for (size_t i = 0; i < n; ++i) {
      const size_t k = keys.As<size_t>(i);
      CHECK(k < size_);
      int s = static_cast<int>(std::upper_bound(offsets_.begin(), offsets_.end(), k) - offsets_.begin()) - 1;
      bucket[s].push_back(i);
    }
    for (int s = 0; s < S; ++s) {
      if (bucket[s].empty()) continue;
      Blob ks(bucket[s].size() * sizeof(size_t));
      for (size_t j = 0; j < bucket[s].size(); ++j) ks.As<size_t>(j) = keys.As<size_t>(bucket[s][j]);
      (*out)[s].push_back(ks);
      if (is_add) {
        Blob vs(bucket[s].size() * sizeof(T));
        for (size_t j = 0; j < bucket[s].size(); ++j) vs.As<T>(j) = kv[1].As<T>(bucket[s][j]);
        (*out)[s].push_back(vs);
        if (kv.size() > 2) (*out)[s].push_back(kv[2]);
      }
    }
    return static_cast<int>(out->size());
  }
  void ProcessReplyGet(std::vector<Blob>& reply, int msg_id) override {
    CHECK(reply.size() == 2);
    const size_t n = reply[0].size<size_t>();
    std::lock_guard<std::mutex> lk(rec_mu_);
    Record& r = rec_.at(msg_id);
    for (size_t i = 0; i < n; ++i) {
      const size_t k = reply[0].As<size_t>(i);
      if (r.all_keys) {
        r.all_keys->push_back(k);
        r.all_vals->push_back(reply[1].As<T>(i));
      } else {
        auto it = r.dst.find(k);
        if (it != r.dst.end()) *it->second = reply[1].As<T>(i);
      }
    }
  }

 protected:
  void OnRequestDone(int msg_id) override {
    std::lock_guard<std::mutex> lk(rec_mu_);
    rec_.erase(msg_id);
  }

 private:
  void Submit(int id, MsgType type, std::vector<Blob> blobs) {
    MessagePtr msg(new Message());
    msg->set_src(MV_Rank());
    msg->set_type(type);
    msg->set_msg_id(id);
    msg->set_table_id(table_id_);
    msg->data() = std::move(blobs);
    Zoo::Get()->SendTo("worker", msg);
  }
  struct Record {
    std::unordered_map<size_t, T*> dst;
    std::vector<size_t>* all_keys = nullptr;
    std::vector<T>* all_vals = nullptr;
  };
  size_t size_;
  std::vector<size_t> offsets_;
  std::mutex rec_mu_;
  std::unordered_map<int, Record> rec_;
};

template <typename T>
class SparseServerTable : public ServerTable {
 public:
  explicit SparseServerTable(size_t size) {
    auto off = RangeOffsets(size, MV_NumServers());
    lo_ = off[MV_ServerId()];
    storage_.assign(off[MV_ServerId() + 1] - lo_, T());
    touched_.assign(storage_.size(), 0);
  }
  template <typename Opt>
  explicit SparseServerTable(const Opt& o) : SparseServerTable(static_cast<size_t>(o.size)) {}
  void ProcessAdd(const std::vector<Blob>& data) override {
    CHECK(data.size() >= 2);
    const size_t n = data[0].size<size_t>();
    for (size_t i = 0; i < n; ++i) {
      const size_t local = data[0].As<size_t>(i) - lo_;
      CHECK(local < storage_.size());
      storage_[local] -= data[1].As<T>(i);        // the server subtracts (sparse_table.h:199-214)
      touched_[local] = 1;
    }
  }
  void ProcessGet(const std::vector<Blob>& data, std::vector<Blob>* result) override {
    const Blob& keys = data[0];
    if (keys.size<size_t>() == 1 && keys.As<size_t>(0) == static_cast<size_t>(-1)) {
      size_t cnt = 0;
      for (unsigned char t : touched_) cnt += t;
      Blob ks(cnt * sizeof(size_t)), vs(cnt * sizeof(T));
      size_t j = 0;
      for (size_t i = 0; i < storage_.size(); ++i)
        if (touched_[i]) {
          ks.As<size_t>(j) = i + lo_;
          vs.As<T>(j) = storage_[i];
          ++j;
        }
      result->push_back(ks);
      result->push_back(vs);
      return;
    }
    const size_t n = keys.size<size_t>();
    Blob vs(n * sizeof(T));
    for (size_t i = 0; i < n; ++i) vs.As<T>(i) = storage_[keys.As<size_t>(i) - lo_];
    result->push_back(keys);
    result->push_back(vs);
  }
  void Store(Stream* s) override {
    s->Write(storage_.data(), storage_.size() * sizeof(T));
    s->Write(touched_.data(), touched_.size());
  }
  void Load(Stream* s) override {
    s->Read(storage_.data(), storage_.size() * sizeof(T));
    s->Read(touched_.data(), touched_.size());
  }

 private:
  size_t lo_ = 0;
  std::vector<T> storage_;
  std::vector<unsigned char> touched_;
};

// FTRL: the same table with {z, n} entries (ftrl_sparse_table.h:11-86)
template <typename T> using FTRLWorkerTable = SparseWorkerTable<FTRLEntry<T>>;
template <typename T> using FTRLServerTable = SparseServerTable<FTRLEntry<T>>;
template <typename T>
struct FTRLTableOption {
  explicit FTRLTableOption(size_t s) : size(s) {}
  size_t size;
  using WorkerTableType = FTRLWorkerTable<T>;
  using ServerTableType = FTRLServerTable<T>;
};

}  // namespace multiverso
#endif
