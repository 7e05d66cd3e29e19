// Public C++ API of the host runtime (counterpart of include/multiverso/multiverso.h:9-65).
#ifndef MULTIVERSO_MULTIVERSO_H_
#define MULTIVERSO_MULTIVERSO_H_
#include <string>
#include "multiverso/table_factory.h"

namespace multiverso {

void MV_Init(int* argc = nullptr, char* argv[] = nullptr);
void MV_Barrier();
void MV_ShutDown(bool finalize_net = true);

int MV_Rank();
int MV_Size();
int MV_NumWorkers();
int MV_NumServers();
int MV_WorkerId();
int MV_ServerId();
int MV_WorkerIdToRank(int worker_id);
int MV_ServerIdToRank(int server_id);

// T in {int, bool, std::string, double}
template <typename T>
void MV_SetFlag(const std::string& name, const T& value);

// Creates the server half and/or worker half according to this rank's role, then barriers
// (collective: every rank must create its tables in the same order -- ids are positional).
template <typename OptionType>
typename OptionType::WorkerTableType* MV_CreateTable(const OptionType& option) {
  auto* table = table_factory::CreateTable(option);
  MV_Barrier();
  return table;
}

// In-place SUM all-reduce (model averaging). T in {char, int, float, double}.
template <typename T>
void MV_Aggregate(T* data, int size);

// Explicit-endpoint bootstrap of the control plane.
int MV_NetBind(int rank, char* endpoint);
int MV_NetConnect(int* ranks, char* endpoints[], int size);
void MV_NetFinalize();

// Checkpoint of a table's server shard + updater state (the reference only exposes the
// Serializable interface and never calls it).
bool MV_SaveTable(int table_id, const std::string& uri);
bool MV_LoadTable(int table_id, const std::string& uri);

}  // namespace multiverso
#endif
