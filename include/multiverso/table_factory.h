// table_factory::CreateTable(option): create the server half if this rank is a server and
// the worker half if it is a worker (include/multiverso/table_factory.h:16-26).
#ifndef MULTIVERSO_TABLE_FACTORY_H_
#define MULTIVERSO_TABLE_FACTORY_H_
#include "multiverso/table_interface.h"
#include "multiverso/zoo.h"

namespace multiverso {
namespace table_factory {
void PushServerTable(ServerTable* table);
void FreeServerTables();
std::vector<ServerTable*>& ServerTables();

template <typename OptionType>
typename OptionType::WorkerTableType* CreateTable(const OptionType& option) {
  if (Zoo::Get()->server_rank() >= 0)
    PushServerTable(new typename OptionType::ServerTableType(option));
  if (Zoo::Get()->worker_rank() >= 0) return new typename OptionType::WorkerTableType(option);
  return nullptr;
}
}  // namespace table_factory
}  // namespace multiverso
#endif
