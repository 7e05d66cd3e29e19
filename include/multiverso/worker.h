// Worker actor: owns the WorkerTable cache; partitions requests per server and completes
// them on replies (src/worker.cpp:12-88).
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
  int RegisterTable(WorkerTable* table);

 private:
  void ProcessGet(MessagePtr& msg);
  void ProcessAdd(MessagePtr& msg);
  void ProcessReplyGet(MessagePtr& msg);
  void ProcessReplyAdd(MessagePtr& msg);
  void FanOut(MessagePtr& msg, MsgType type);
  std::mutex mu_;
  std::vector<WorkerTable*> cache_;
};
}  // namespace multiverso
#endif
