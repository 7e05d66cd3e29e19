// WorkerTable / ServerTable: the table abstraction of the host backend (counterpart of
// include/multiverso/table_interface.h:24-80, src/table.cpp). Differences by design:
//   * request ids are recycled and each in-flight request owns its destination record, so
//     several Gets may be outstanding per table;
//   * Partition receives the request kind and produces one blob vector per *server id*;
//     the Worker actor maps ids to ranks (fixes Q4 for every table type).
#ifndef MULTIVERSO_TABLE_INTERFACE_H_
#define MULTIVERSO_TABLE_INTERFACE_H_
#include <cstdint>
#include <map>
#include <memory>
#include <mutex>
#include <unordered_map>
#include <vector>
#include "multiverso/blob.h"
#include "multiverso/io/io.h"
#include "multiverso/message.h"
#include "multiverso/updater/updater.h"
#include "multiverso/util/waiter.h"

namespace multiverso {

using integer_t = int64_t;   // 64-bit ids end to end (the reference's is int32, Q22)

class WorkerTable {
 public:
  WorkerTable();
  virtual ~WorkerTable();

  void Get(Blob keys, const GetOption* option = nullptr);
  void Add(Blob keys, Blob values, const AddOption* option = nullptr);
  int GetAsync(Blob keys, const GetOption* option = nullptr);
  int AddAsync(Blob keys, Blob values, const AddOption* option = nullptr);
  void Wait(int id);

  // Called on the worker actor thread -------------------------------------------------
  // Split a request into per-server-id blob vectors; returns the number of partitions.
  virtual int Partition(const std::vector<Blob>& kv, MsgType partition_type,
                        std::unordered_map<int, std::vector<Blob>>* out) = 0;
  // Consume one server's Get reply for request `msg_id`.
  virtual void ProcessReplyGet(std::vector<Blob>& reply_data, int msg_id) = 0;
  void Reset(int msg_id, int num_wait);
  void Notify(int msg_id);
  int table_id() const { return table_id_; }

 protected:
  // Hook invoked when a request id is retired (subclasses free per-request records).
  virtual void OnRequestDone(int /*msg_id*/) {}
  int NewRequest();
  int table_id_;
  std::mutex mu_;
  std::unordered_map<int, std::unique_ptr<Waiter>> waiting_;
  std::vector<int> free_ids_;
  int next_id_ = 0;
};

class Serializable {
 public:
  virtual ~Serializable() = default;
  virtual void Store(Stream* s) = 0;
  virtual void Load(Stream* s) = 0;
};

class ServerTable : public Serializable {
 public:
  ServerTable();
  ~ServerTable() override = default;
  virtual void ProcessAdd(const std::vector<Blob>& data) = 0;
  virtual void ProcessGet(const std::vector<Blob>& data, std::vector<Blob>* result) = 0;
  int table_id() const { return table_id_; }

 protected:
  int table_id_;
};

// The AddOption travelling as the trailing blob of an Add request (index `pos`), or the
// defaults with worker 0 when the client sent none.
inline AddOption AddOptionFrom(const std::vector<Blob>& data, size_t pos) {
  AddOption opt;
  if (data.size() > pos && data[pos].size() >= opt.size()) opt.CopyFrom(data[pos].data(), data[pos].size());
  if (opt.worker_id() < 0) opt.set_worker_id(0);
  return opt;
}

// Binds an Option struct to its worker / server table types.
#define DEFINE_TABLE_TYPE(template_type, worker_table_type, server_table_type) \
  using WorkerTableType = worker_table_type<template_type>;                    \
  using ServerTableType = server_table_type<template_type>;

}  // namespace multiverso
#endif
