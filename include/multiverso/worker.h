// Worker actor: the request side of the parameter server inside one process. It keeps the
// process's WorkerTables (indexed by table id), splits every Get / Add request into one message
// per destination server, and completes the caller's waiter as the replies come back
// (reference behaviour: src/worker.cpp:12-88).
#ifndef MULTIVERSO_WORKER_H_
#define MULTIVERSO_WORKER_H_
#include <mutex>
#include <vector>
#include "multiverso/actor.h"

namespace multiverso {
class WorkerTable;

class Worker : public Actor {
 public:
  Worker();
  // Returns the table id (tables are created in the same order on every rank).
  int RegisterTable(WorkerTable* table);

 private:
  WorkerTable* TableOf(const MessagePtr& msg);
  // Request_Get / Request_Add from a user thread: Partition -> arm the waiter -> fan out.
  void Dispatch(MessagePtr& request, MsgType type);
  // Reply_Get (carries data for the table) / Reply_Add (bare acknowledgement).
  void Complete(MessagePtr& reply, bool carries_data);
  std::mutex tables_mu_;
  std::vector<WorkerTable*> tables_;
};
}  // namespace multiverso
#endif
