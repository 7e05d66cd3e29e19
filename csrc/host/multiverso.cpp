// Public C++ API (see include/multiverso/multiverso.h; reference src/multiverso.cpp:11-78).
#include "multiverso/multiverso.h"
#include <memory>
#include "multiverso/dashboard.h"
#include "multiverso/io/io.h"
#include "multiverso/net.h"
#include "multiverso/util/configure.h"
#include "multiverso/util/log.h"
#include "multiverso/zoo.h"

namespace multiverso {

void MV_Init(int* argc, char* argv[]) { Zoo::Get()->Start(argc, argv); }

void MV_ShutDown(bool finalize_net) {
  Zoo::Get()->Stop(finalize_net);
  table_factory::FreeServerTables();
}

void MV_Barrier() { Zoo::Get()->Barrier(); }
int MV_Rank() { return Zoo::Get()->rank(); }
int MV_Size() { return Zoo::Get()->size(); }
int MV_NumWorkers() { return Zoo::Get()->num_workers(); }
int MV_NumServers() { return Zoo::Get()->num_servers(); }
int MV_WorkerId() { return Zoo::Get()->worker_rank(); }
int MV_ServerId() { return Zoo::Get()->server_rank(); }
int MV_WorkerIdToRank(int id) { return Zoo::Get()->worker_id_to_rank(id); }
int MV_ServerIdToRank(int id) { return Zoo::Get()->server_id_to_rank(id); }

template <typename T>
void MV_SetFlag(const std::string& name, const T& value) { SetCMDFlag<T>(name, value); }
template void MV_SetFlag<int>(const std::string&, const int&);
template void MV_SetFlag<bool>(const std::string&, const bool&);
template void MV_SetFlag<std::string>(const std::string&, const std::string&);
template void MV_SetFlag<double>(const std::string&, const double&);

template <typename T>
void MV_Aggregate(T* data, int size) { net::Allreduce<T>(data, static_cast<size_t>(size)); }
template void MV_Aggregate<char>(char*, int);
template void MV_Aggregate<int>(int*, int);
template void MV_Aggregate<float>(float*, int);
template void MV_Aggregate<double>(double*, int);

int MV_NetBind(int rank, char* endpoint) { return NetInterface::Get()->Bind(rank, endpoint); }
int MV_NetConnect(int* ranks, char* endpoints[], int size) {
  return NetInterface::Get()->Connect(ranks, endpoints, size);
}
void MV_NetFinalize() { NetInterface::Get()->Finalize(); }

namespace {
std::string ShardUri(const std::string& uri) {
  return uri + ".shard" + std::to_string(MV_ServerId());
}
}  // namespace

bool MV_SaveTable(int table_id, const std::string& uri) {
  auto& tables = table_factory::ServerTables();
  bool ok = true;
  if (MV_ServerId() >= 0) {
    // decide first, ALWAYS reach the barrier: worker-only ranks are already waiting in it
    ok = table_id >= 0 && table_id < static_cast<int>(tables.size());
    if (ok) {
      std::unique_ptr<Stream> s(StreamFactory::GetStream(URI(ShardUri(uri)), FileOpenMode::BinaryWrite));
      ok = s && s->Good();
      if (ok) tables[table_id]->Store(s.get());
      ok = ok && s->Good();
    }
  }
  MV_Barrier();
  return ok;
}

bool MV_LoadTable(int table_id, const std::string& uri) {
  auto& tables = table_factory::ServerTables();
  bool ok = true;
  if (MV_ServerId() >= 0) {
    ok = table_id >= 0 && table_id < static_cast<int>(tables.size());
    if (ok) {
      std::unique_ptr<Stream> s(StreamFactory::GetStream(URI(ShardUri(uri)), FileOpenMode::BinaryRead));
      ok = s && s->Good();
      if (ok) tables[table_id]->Load(s.get());
      ok = ok && s->Good() && !s->Failed();   // tables mark the stream when the checkpoint is truncated
    }
  }
  MV_Barrier();
  return ok;
}

}  // namespace multiverso
