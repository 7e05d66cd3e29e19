// Worker actor (see include/multiverso/worker.h).
#include "multiverso/worker.h"
#include "multiverso/dashboard.h"
#include "multiverso/table_interface.h"
#include "multiverso/util/log.h"
#include "multiverso/zoo.h"

namespace multiverso {

Worker::Worker() : Actor(actor::kWorker) {
  RegisterHandler(MsgType::Request_Get, [this](MessagePtr& m) {
    MONITOR_BEGIN(WORKER_PROCESS_GET)
    Dispatch(m, MsgType::Request_Get);
    MONITOR_END(WORKER_PROCESS_GET)
  });
  RegisterHandler(MsgType::Request_Add, [this](MessagePtr& m) {
    MONITOR_BEGIN(WORKER_PROCESS_ADD)
    Dispatch(m, MsgType::Request_Add);
    MONITOR_END(WORKER_PROCESS_ADD)
  });
  RegisterHandler(MsgType::Reply_Get, [this](MessagePtr& m) {
    MONITOR_BEGIN(WORKER_PROCESS_REPLY_GET)
    Complete(m, true);
    MONITOR_END(WORKER_PROCESS_REPLY_GET)
  });
  RegisterHandler(MsgType::Reply_Add, [this](MessagePtr& m) { Complete(m, false); });
}

int Worker::RegisterTable(WorkerTable* table) {
  std::lock_guard<std::mutex> lk(tables_mu_);
  tables_.push_back(table);
  return static_cast<int>(tables_.size()) - 1;
}

WorkerTable* Worker::TableOf(const MessagePtr& msg) {
  std::lock_guard<std::mutex> lk(tables_mu_);
  return tables_.at(msg->table_id());
}

// Partition the request by server id, arm the waiter with the partition count, then send
// one message per destination through the communicator.
void Worker::Dispatch(MessagePtr& request, MsgType type) {
  WorkerTable* table = TableOf(request);
  std::unordered_map<int, std::vector<Blob>> parts;
  const int n = table->Partition(request->data(), type, &parts);
  table->Reset(request->msg_id(), n);
  if (n == 0) return;
  for (auto& kv : parts) {
    MessagePtr out(new Message());
    out->set_src(Zoo::Get()->rank());
    out->set_dst(Zoo::Get()->server_id_to_rank(kv.first));
    out->set_type(type);
    out->set_table_id(request->table_id());
    out->set_msg_id(request->msg_id());
    out->data() = std::move(kv.second);
    SendTo(actor::kCommunicator, out);
  }
}

void Worker::Complete(MessagePtr& reply, bool carries_data) {
  WorkerTable* table = TableOf(reply);
  if (carries_data) table->ProcessReplyGet(reply->data(), reply->msg_id());
  table->Notify(reply->msg_id());
}

}  // namespace multiverso
