// KVTable<K,V>: distributed hash map, header-only (counterpart of
// include/multiverso/table/kv_table.h:18-124). Keys are hash-partitioned: non-negative
// key % num_servers -> server *id* (the Worker actor maps ids to ranks, Q4). The worker keeps
// a local cache exposed through raw(); the server does table[k] += v. Store/Load are
// implemented (the reference's are Fatal("Not implemented"), Q14).
#ifndef MULTIVERSO_TABLE_KV_TABLE_H_
#define MULTIVERSO_TABLE_KV_TABLE_H_
#include <unordered_map>
#include <vector>
#include "multiverso/multiverso.h"
#include "multiverso/table_interface.h"
#include "multiverso/util/log.h"

namespace multiverso {

template <typename Key, typename Val> class KVWorkerTable;
template <typename Key, typename Val> class KVServerTable;

template <typename Key, typename Val>
struct KVTableOption {
  using WorkerTableType = KVWorkerTable<Key, Val>;
  using ServerTableType = KVServerTable<Key, Val>;
};

template <typename Key, typename Val>
class KVWorkerTable : public WorkerTable {
 public:
  KVWorkerTable() = default;
  explicit KVWorkerTable(const KVTableOption<Key, Val>&) {}

  void Get(Key key) { WorkerTable::Get(Blob(&key, sizeof(Key))); }
  void Get(const std::vector<Key>& keys) {
    WorkerTable::Get(Blob(keys.data(), keys.size() * sizeof(Key)));
  }
  void Add(Key key, Val value) {
    WorkerTable::Add(Blob(&key, sizeof(Key)), Blob(&value, sizeof(Val)));
  }
  void Add(const std::vector<Key>& keys, const std::vector<Val>& values) {
    CHECK(keys.size() == values.size());
    WorkerTable::Add(Blob(keys.data(), keys.size() * sizeof(Key)),
                     Blob(values.data(), values.size() * sizeof(Val)));
  }
  std::unordered_map<Key, Val>& raw() { return table_; }

  int Partition(const std::vector<Blob>& kv, MsgType,
                std::unordered_map<int, std::vector<Blob>>* out) override {
    CHECK(kv.size() == 1 || kv.size() == 2);
    const int S = MV_NumServers();
    const size_t n = kv[0].size<Key>();
    std::vector<std::vector<size_t>> bucket(S);
    for (size_t i = 0; i < n; ++i) bucket[ServerOf(kv[0].As<Key>(i), S)].push_back(i);
    for (int s = 0; s < S; ++s) {
      if (bucket[s].empty()) continue;
      Blob keys(bucket[s].size() * sizeof(Key));
      for (size_t j = 0; j < bucket[s].size(); ++j) keys.As<Key>(j) = kv[0].As<Key>(bucket[s][j]);
      (*out)[s].push_back(keys);
      if (kv.size() == 2) {
        Blob vals(bucket[s].size() * sizeof(Val));
        for (size_t j = 0; j < bucket[s].size(); ++j) vals.As<Val>(j) = kv[1].As<Val>(bucket[s][j]);
        (*out)[s].push_back(vals);
      }
    }
    return static_cast<int>(out->size());
  }
  void ProcessReplyGet(std::vector<Blob>& data, int) override {
    CHECK(data.size() == 2);
    const size_t n = data[0].size<Key>();
    std::lock_guard<std::mutex> lk(cache_mu_);
    for (size_t i = 0; i < n; ++i) table_[data[0].As<Key>(i)] = data[1].As<Val>(i);
  }

 private:
  static int ServerOf(Key k, int S) {
    long long m = static_cast<long long>(k) % S;
    return static_cast<int>(m < 0 ? m + S : m);
  }
  std::mutex cache_mu_;
  std::unordered_map<Key, Val> table_;
};

template <typename Key, typename Val>
class KVServerTable : public ServerTable {
 public:
  KVServerTable() = default;
  explicit KVServerTable(const KVTableOption<Key, Val>&) {}
  void ProcessGet(const std::vector<Blob>& data, std::vector<Blob>* result) override {
    CHECK(data.size() == 1);
    const size_t n = data[0].size<Key>();
    Blob vals(n * sizeof(Val));
    for (size_t i = 0; i < n; ++i) vals.As<Val>(i) = table_[data[0].As<Key>(i)];
    result->push_back(data[0]);
    result->push_back(vals);
  }
  void ProcessAdd(const std::vector<Blob>& data) override {
    CHECK(data.size() == 2);
    const size_t n = data[0].size<Key>();
    for (size_t i = 0; i < n; ++i) table_[data[0].As<Key>(i)] += data[1].As<Val>(i);
  }
  void Store(Stream* s) override {
    uint64_t n = table_.size();
    s->Write(&n, sizeof(n));
    for (auto& kv : table_) {
      s->Write(&kv.first, sizeof(Key));
      s->Write(&kv.second, sizeof(Val));
    }
  }
  void Load(Stream* s) override {
    uint64_t n = 0;
    if (s->Read(&n, sizeof(n)) != sizeof(n)) return;
    table_.clear();
    for (uint64_t i = 0; i < n; ++i) {
      Key k;
      Val v;
      s->Read(&k, sizeof(Key));
      s->Read(&v, sizeof(Val));
      table_[k] = v;
    }
  }

 private:
  std::unordered_map<Key, Val> table_;
};

}  // namespace multiverso
#endif
