// ArrayTable<T>: dense 1-D table, whole-table Get/Add only (counterpart of
// include/multiverso/table/array_table.h:13-73, src/table/array_table.cpp). Contiguous
// element ranges, size/num_servers each, last server takes the remainder; any size >= 1
// (servers beyond `size` own nothing; the reference needs size > num_servers, Q5).
#ifndef MULTIVERSO_TABLE_ARRAY_TABLE_H_
#define MULTIVERSO_TABLE_ARRAY_TABLE_H_
#include <vector>
#include "multiverso/table_interface.h"

namespace multiverso {

// offsets[s] .. offsets[s+1] = the slice of server s
std::vector<size_t> RangeOffsets(size_t total, int num_servers);

template <typename T> class ArrayWorker;
template <typename T> class ArrayServer;

template <typename T>
struct ArrayTableOption {
  explicit ArrayTableOption(size_t s) : size(s) {}
  size_t size;
  DEFINE_TABLE_TYPE(T, ArrayWorker, ArrayServer);
};

template <typename T>
class ArrayWorker : public WorkerTable {
 public:
  explicit ArrayWorker(size_t size);
  explicit ArrayWorker(const ArrayTableOption<T>& option) : ArrayWorker(option.size) {}
  void Get(T* data, size_t size);
  int GetAsync(T* data, size_t size);
  void Add(T* data, size_t size, const AddOption* option = nullptr);
  int AddAsync(T* data, size_t size, const AddOption* option = nullptr);
  size_t size() const { return size_; }

  int Partition(const std::vector<Blob>& kv, MsgType partition_type,
                std::unordered_map<int, std::vector<Blob>>* out) override;
  void ProcessReplyGet(std::vector<Blob>& reply_data, int msg_id) override;

 protected:
  void OnRequestDone(int msg_id) override;

 private:
  size_t size_;
  int num_server_;
  std::vector<size_t> offsets_;
  std::mutex dest_mu_;
  std::unordered_map<int, T*> dest_;
};

template <typename T>
class ArrayServer : public ServerTable {
 public:
  explicit ArrayServer(size_t size);
  explicit ArrayServer(const ArrayTableOption<T>& option) : ArrayServer(option.size) {}
  ~ArrayServer() override;
  void ProcessAdd(const std::vector<Blob>& data) override;
  void ProcessGet(const std::vector<Blob>& data, std::vector<Blob>* result) override;
  void Store(Stream* s) override;
  void Load(Stream* s) override;
  T* raw() { return storage_.data(); }
  size_t shard_size() const { return storage_.size(); }

 private:
  int server_id_;
  std::vector<T> storage_;
  Updater<T>* updater_;
};

}  // namespace multiverso
#endif
