#include "multiverso/table_factory.h"
namespace multiverso {
namespace table_factory {
std::vector<ServerTable*>& ServerTables() {
  static auto* v = new std::vector<ServerTable*>();
  return *v;
}
void PushServerTable(ServerTable* table) { ServerTables().push_back(table); }
void FreeServerTables() {
  for (ServerTable* t : ServerTables()) delete t;
  ServerTables().clear();
}
}  // namespace table_factory
}  // namespace multiverso
