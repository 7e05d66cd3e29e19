// Worker actor (see include/multiverso/worker.h).
#include "multiverso/worker.h"
#include "multiverso/dashboard.h"
#include "multiverso/table_interface.h"
#include "multiverso/util/log.h"
#include "multiverso/zoo.h"

namespace multiverso {

Worker::Worker() : Actor(actor::kWorker) {
  RegisterHandler(MsgType::Request_Get, [this](MessagePtr& m) { ProcessGet(m); });
  RegisterHandler(MsgType::Request_Add, [this](MessagePtr& m) { ProcessAdd(m); });
  RegisterHandler(MsgType::Reply_Get, [this](MessagePtr& m) { ProcessReplyGet(m); });
  RegisterHandler(MsgType::Reply_Add, [this](MessagePtr& m) { ProcessReplyAdd(m); });
}

int Worker::RegisterTable(WorkerTable* table) {
  std::lock_guard<std::mutex> lk(mu_);
  cache_.push_back(table);
  return static_cast<int>(cache_.size()) - 1;
}

// Partition the request by server id, arm the waiter with the partition count, then send
// one message per destination through the communicator.
void Worker::FanOut(MessagePtr& msg, MsgType type) {
  WorkerTable* table;
  {
    std::lock_guard<std::mutex> lk(mu_);
    table = cache_.at(msg->table_id());
  }
  std::unordered_map<int, std::vector<Blob>> parts;
  const int n = table->Partition(msg->data(), type, &parts);
  table->Reset(msg->msg_id(), n);
  if (n == 0) return;
  for (auto& kv : parts) {
    MessagePtr out(new Message());
    out->set_src(Zoo::Get()->rank());
    out->set_dst(Zoo::Get()->server_id_to_rank(kv.first));
    out->set_type(type);
    out->set_table_id(msg->table_id());
    out->set_msg_id(msg->msg_id());
    out->data() = std::move(kv.second);
    SendTo(actor::kCommunicator, out);
  }
}

void Worker::ProcessGet(MessagePtr& msg) {
  MONITOR_BEGIN(WORKER_PROCESS_GET)
  FanOut(msg, MsgType::Request_Get);
  MONITOR_END(WORKER_PROCESS_GET)
}

void Worker::ProcessAdd(MessagePtr& msg) {
  MONITOR_BEGIN(WORKER_PROCESS_ADD)
  FanOut(msg, MsgType::Request_Add);
  MONITOR_END(WORKER_PROCESS_ADD)
}

void Worker::ProcessReplyGet(MessagePtr& msg) {
  MONITOR_BEGIN(WORKER_PROCESS_REPLY_GET)
  WorkerTable* table;
  {
    std::lock_guard<std::mutex> lk(mu_);
    table = cache_.at(msg->table_id());
  }
  table->ProcessReplyGet(msg->data(), msg->msg_id());
  table->Notify(msg->msg_id());
  MONITOR_END(WORKER_PROCESS_REPLY_GET)
}

void Worker::ProcessReplyAdd(MessagePtr& msg) {
  WorkerTable* table;
  {
    std::lock_guard<std::mutex> lk(mu_);
    table = cache_.at(msg->table_id());
  }
  table->Notify(msg->msg_id());
}

}  // namespace multiverso
