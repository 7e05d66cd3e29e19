// C API (see include/multiverso/c_api.h; reference src/c_api.cpp:9-93).
#include "multiverso/c_api.h"
#include <string>
#include <vector>
#include "multiverso/dashboard.h"
#include "multiverso/multiverso.h"
#include "multiverso/net.h"
#include "multiverso/table/array_table.h"
#include "multiverso/table/kv_table.h"
#include "multiverso/table/matrix.h"
#include "multiverso/table/matrix_table.h"
#include "multiverso/util/configure.h"
#include "multiverso/util/log.h"

namespace mv = multiverso;
using mv::integer_t;

extern "C" {

void MV_Init(int* argc, char* argv[]) { mv::MV_Init(argc, argv); }
void MV_ShutDown() { mv::MV_ShutDown(); }
void MV_ShutDownEx(int finalize_net) { mv::MV_ShutDown(finalize_net != 0); }
void MV_Barrier() { mv::MV_Barrier(); }
int MV_NumWorkers() { return mv::MV_NumWorkers(); }
int MV_NumServers() { return mv::MV_NumServers(); }
int MV_WorkerId() { return mv::MV_WorkerId(); }
int MV_ServerId() { return mv::MV_ServerId(); }
int MV_Rank() { return mv::MV_Rank(); }
int MV_Size() { return mv::MV_Size(); }
int MV_WorkerIdToRank(int id) { return mv::MV_WorkerIdToRank(id); }
int MV_ServerIdToRank(int id) { return mv::MV_ServerIdToRank(id); }

// ---- reference-compatible float tables ---------------------------------------------------
void MV_NewArrayTable(int size, TableHandler* out) {
  *out = mv::MV_CreateTable(mv::ArrayTableOption<float>(static_cast<size_t>(size)));
}
void MV_GetArrayTable(TableHandler h, float* data, int size) {
  static_cast<mv::ArrayWorker<float>*>(h)->Get(data, static_cast<size_t>(size));
}
void MV_AddArrayTable(TableHandler h, float* data, int size) {
  static_cast<mv::ArrayWorker<float>*>(h)->Add(data, static_cast<size_t>(size));
}
void MV_AddAsyncArrayTable(TableHandler h, float* data, int size) {
  static_cast<mv::ArrayWorker<float>*>(h)->AddAsync(data, static_cast<size_t>(size));
}
void MV_NewMatrixTable(int num_row, int num_col, TableHandler* out) {
  *out = mv::MV_CreateTable(mv::MatrixTableOption<float>(num_row, num_col));
}
void MV_GetMatrixTableAll(TableHandler h, float* data, int size) {
  static_cast<mv::MatrixWorkerTable<float>*>(h)->Get(data, static_cast<size_t>(size));
}
void MV_AddMatrixTableAll(TableHandler h, float* data, int size) {
  static_cast<mv::MatrixWorkerTable<float>*>(h)->Add(data, static_cast<size_t>(size));
}
void MV_AddAsyncMatrixTableAll(TableHandler h, float* data, int size) {
  static_cast<mv::MatrixWorkerTable<float>*>(h)->AddAsync(data, static_cast<size_t>(size));
}
static std::vector<integer_t> Widen(const int* ids, int n) { return std::vector<integer_t>(ids, ids + n); }
void MV_GetMatrixTableByRows(TableHandler h, float* data, int size, int row_ids[], int n) {
  auto ids = Widen(row_ids, n);
  static_cast<mv::MatrixWorkerTable<float>*>(h)->Get(data, static_cast<size_t>(size), ids.data(), n);
}
void MV_AddMatrixTableByRows(TableHandler h, float* data, int size, int row_ids[], int n) {
  auto ids = Widen(row_ids, n);
  static_cast<mv::MatrixWorkerTable<float>*>(h)->Add(data, static_cast<size_t>(size), ids.data(), n);
}
void MV_AddAsyncMatrixTableByRows(TableHandler h, float* data, int size, int row_ids[], int n) {
  auto ids = Widen(row_ids, n);
  static_cast<mv::MatrixWorkerTable<float>*>(h)->AddAsync(data, static_cast<size_t>(size), ids.data(), n);
}

// ---- flags / net / aggregate -----------------------------------------------------------------
int MV_SetFlagInt(const char* name, int v) { return mv::config::Registry::Get().Set<int>(name, v) ? 0 : -1; }
int MV_SetFlagBool(const char* name, int v) { return mv::config::Registry::Get().Set<bool>(name, v != 0) ? 0 : -1; }
int MV_SetFlagDouble(const char* name, double v) { return mv::config::Registry::Get().Set<double>(name, v) ? 0 : -1; }
int MV_SetFlagString(const char* name, const char* v) {
  return mv::config::Registry::Get().Set<std::string>(name, std::string(v)) ? 0 : -1;
}
int MV_NetBindC(int rank, const char* endpoint) {
  std::string e(endpoint);
  return mv::MV_NetBind(rank, const_cast<char*>(e.c_str()));
}
int MV_NetConnectC(int* ranks, const char* endpoints[], int size) {
  std::vector<std::string> keep(endpoints, endpoints + size);
  std::vector<char*> ptrs;
  for (auto& s : keep) ptrs.push_back(const_cast<char*>(s.c_str()));
  return mv::MV_NetConnect(ranks, ptrs.data(), size);
}
void MV_NetFinalizeC() { mv::MV_NetFinalize(); }
void MV_AggregateFloat(float* d, int64_t n) { mv::net::Allreduce<float>(d, static_cast<size_t>(n)); }
void MV_AggregateDouble(double* d, int64_t n) { mv::net::Allreduce<double>(d, static_cast<size_t>(n)); }
void MV_AggregateInt(int* d, int64_t n) { mv::net::Allreduce<int>(d, static_cast<size_t>(n)); }
void MV_AggregateChar(char* d, int64_t n) { mv::net::Allreduce<char>(d, static_cast<size_t>(n)); }

// ---- typed / 64-bit tables -------------------------------------------------------------------
#define MV_DISPATCH_DTYPE(dtype, ...)                                   \
  switch (dtype) {                                                      \
    case 0: { using T = float; __VA_ARGS__; } break;                    \
    case 1: { using T = double; __VA_ARGS__; } break;                   \
    case 2: { using T = int; __VA_ARGS__; } break;                      \
    default: mv::Log::Fatal("c_api: unknown dtype %d", dtype);          \
  }

void MV_NewArrayTable64(int64_t size, int dtype, TableHandler* out) {
  MV_DISPATCH_DTYPE(dtype, *out = mv::MV_CreateTable(mv::ArrayTableOption<T>(static_cast<size_t>(size))));
}
void MV_GetArrayTable64(TableHandler h, int dtype, void* data, int64_t size) {
  MV_DISPATCH_DTYPE(dtype, static_cast<mv::ArrayWorker<T>*>(h)->Get(static_cast<T*>(data), static_cast<size_t>(size)));
}
void MV_AddArrayTable64(TableHandler h, int dtype, void* data, int64_t size, const void* opt20, int async) {
  mv::AddOption opt;
  if (opt20) opt.CopyFrom(static_cast<const char*>(opt20), 20);
  MV_DISPATCH_DTYPE(dtype, {
    auto* t = static_cast<mv::ArrayWorker<T>*>(h);
    if (async) t->AddAsync(static_cast<T*>(data), static_cast<size_t>(size), &opt);
    else t->Add(static_cast<T*>(data), static_cast<size_t>(size), &opt);
  });
}
void MV_NewMatrixTable64(int64_t num_row, int64_t num_col, int dtype, int is_sparse, int is_pipeline,
                         int random_init, double min_value, double max_value, TableHandler* out) {
  MV_DISPATCH_DTYPE(dtype, {
    if (is_sparse) {
      mv::MatrixOption<T> o;
      o.num_row = num_row; o.num_col = num_col; o.is_sparse = true; o.is_pipeline = is_pipeline != 0;
      *out = static_cast<mv::MatrixWorkerTable<T>*>(mv::MV_CreateTable(o));
    } else if (random_init) {
      *out = mv::MV_CreateTable(mv::MatrixTableOption<T>(num_row, num_col, static_cast<T>(min_value), static_cast<T>(max_value)));
    } else {
      *out = mv::MV_CreateTable(mv::MatrixTableOption<T>(num_row, num_col));
    }
  });
}
void MV_GetMatrixTable64(TableHandler h, int dtype, void* data, int64_t size, const int64_t* row_ids,
                         int64_t n, int worker_id_opt) {
  mv::GetOption gopt;
  gopt.set_worker_id(worker_id_opt);
  MV_DISPATCH_DTYPE(dtype, {
    auto* t = static_cast<mv::MatrixWorkerTable<T>*>(h);
    if (row_ids == nullptr) t->Get(static_cast<T*>(data), static_cast<size_t>(size), &gopt);
    else t->Get(static_cast<T*>(data), static_cast<size_t>(size), const_cast<integer_t*>(row_ids), static_cast<int>(n), &gopt);
  });
}
void MV_AddMatrixTable64(TableHandler h, int dtype, void* data, int64_t size, const int64_t* row_ids,
                         int64_t n, const void* opt20, int async) {
  mv::AddOption opt;
  if (opt20) opt.CopyFrom(static_cast<const char*>(opt20), 20);
  MV_DISPATCH_DTYPE(dtype, {
    auto* t = static_cast<mv::MatrixWorkerTable<T>*>(h);
    int id;
    if (row_ids == nullptr) id = t->AddAsync(static_cast<T*>(data), static_cast<size_t>(size), &opt);
    else id = t->AddAsync(static_cast<T*>(data), static_cast<size_t>(size), const_cast<integer_t*>(row_ids), static_cast<int>(n), &opt);
    if (!async) t->Wait(id);
  });
}

#define MV_DISPATCH_VAL(vt, ...)                                         \
  switch (vt) {                                                          \
    case 0: { using V = float; __VA_ARGS__; } break;                     \
    case 1: { using V = double; __VA_ARGS__; } break;                    \
    case 2: { using V = int; __VA_ARGS__; } break;                       \
    case 3: { using V = long long; __VA_ARGS__; } break;                 \
    default: mv::Log::Fatal("c_api: unknown value dtype %d", vt);        \
  }

void MV_NewKVTable(int vt, TableHandler* out) {
  MV_DISPATCH_VAL(vt, *out = mv::MV_CreateTable(mv::KVTableOption<long long, V>()));
}
void MV_KVAdd(TableHandler h, int vt, const int64_t* keys, const void* vals, int64_t n) {
  MV_DISPATCH_VAL(vt, {
    auto* t = static_cast<mv::KVWorkerTable<long long, V>*>(h);
    std::vector<long long> k(keys, keys + n);
    std::vector<V> v(static_cast<const V*>(vals), static_cast<const V*>(vals) + n);
    t->Add(k, v);
  });
}
void MV_KVGet(TableHandler h, int vt, const int64_t* keys, void* vals, int64_t n) {
  MV_DISPATCH_VAL(vt, {
    auto* t = static_cast<mv::KVWorkerTable<long long, V>*>(h);
    std::vector<long long> k(keys, keys + n);
    t->Get(k);
    for (int64_t i = 0; i < n; ++i) static_cast<V*>(vals)[i] = t->raw()[k[i]];
  });
}
int MV_TableId(TableHandler h) { return static_cast<mv::WorkerTable*>(h)->table_id(); }
int MV_SaveTableC(int table_id, const char* uri) { return mv::MV_SaveTable(table_id, uri) ? 0 : -1; }
int MV_LoadTableC(int table_id, const char* uri) { return mv::MV_LoadTable(table_id, uri) ? 0 : -1; }
void MV_DashboardDisplay() { mv::Dashboard::Display(); }
const char* MV_Version() { return "multiverso-b200 0.1.0 (host runtime)"; }

}  // extern "C"
