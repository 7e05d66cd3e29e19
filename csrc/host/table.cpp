// WorkerTable / ServerTable bases (see include/multiverso/table_interface.h).
#include "multiverso/table_interface.h"
#include "multiverso/actor.h"
#include "multiverso/dashboard.h"
#include "multiverso/util/configure.h"
#include "multiverso/util/log.h"
#include "multiverso/zoo.h"

namespace multiverso {

MV_DEFINE_double(request_stall_warn_s, 60.0, "log a line whenever a table request has waited this long for its servers");

WorkerTable::WorkerTable() { table_id_ = Zoo::Get()->RegisterTable(this); }
WorkerTable::~WorkerTable() = default;
ServerTable::ServerTable() { table_id_ = Zoo::Get()->RegisterTable(this); }

int WorkerTable::NewRequest() {
  std::lock_guard<std::mutex> lk(mu_);
  int id;
  if (!free_ids_.empty()) {
    id = free_ids_.back();
    free_ids_.pop_back();
  } else {
    id = next_id_++;
  }
  waiting_[id].reset(new Waiter(1));   // armed; Reset() sets the real partition count
  return id;
}

namespace {
void Submit(int table_id, int id, MsgType type, std::vector<Blob>&& blobs) {
  MessagePtr msg(new Message());
  msg->set_src(Zoo::Get()->rank());
  msg->set_type(type);
  msg->set_msg_id(id);
  msg->set_table_id(table_id);
  msg->data() = std::move(blobs);
  Zoo::Get()->SendTo(actor::kWorker, msg);
}
}  // namespace

int WorkerTable::GetAsync(Blob keys, const GetOption* option) {
  const int id = NewRequest();
  std::vector<Blob> blobs;
  blobs.push_back(std::move(keys));
  if (option) blobs.emplace_back(option->data(), option->size());
  Submit(table_id_, id, MsgType::Request_Get, std::move(blobs));
  return id;
}

int WorkerTable::AddAsync(Blob keys, Blob values, const AddOption* option) {
  const int id = NewRequest();
  std::vector<Blob> blobs;
  blobs.push_back(std::move(keys));
  blobs.push_back(std::move(values));
  if (option) blobs.emplace_back(option->data(), option->size());
  Submit(table_id_, id, MsgType::Request_Add, std::move(blobs));
  return id;
}

void WorkerTable::Get(Blob keys, const GetOption* option) {
  MONITOR_BEGIN(WORKER_TABLE_SYNC_GET)
  Wait(GetAsync(std::move(keys), option));
  MONITOR_END(WORKER_TABLE_SYNC_GET)
}

void WorkerTable::Add(Blob keys, Blob values, const AddOption* option) {
  MONITOR_BEGIN(WORKER_TABLE_SYNC_ADD)
  Wait(AddAsync(std::move(keys), std::move(values), option));
  MONITOR_END(WORKER_TABLE_SYNC_ADD)
}

void WorkerTable::Wait(int id) {
  Waiter* w = nullptr;
  {
    std::lock_guard<std::mutex> lk(mu_);
    auto it = waiting_.find(id);
    if (it == waiting_.end()) return;   // already waited
    w = it->second.get();
  }
  // A request that waits this long is worth a line in the log (the reference hangs silently when a
  // server died or a BSP schedule is uneven); the wait itself has no deadline.
  double waited = 0;
  while (!w->WaitFor(MV_CONFIG(request_stall_warn_s))) {
    waited += MV_CONFIG(request_stall_warn_s);
    Log::Error("table %d: request %d still waits for %d server repl%s after %.0f s\n", table_id_, id, w->outstanding(),
               w->outstanding() == 1 ? "y" : "ies", waited);
  }
  OnRequestDone(id);
  std::lock_guard<std::mutex> lk(mu_);
  waiting_.erase(id);
  free_ids_.push_back(id);
}

void WorkerTable::Reset(int msg_id, int num_wait) {
  std::lock_guard<std::mutex> lk(mu_);
  auto it = waiting_.find(msg_id);
  if (it != waiting_.end()) it->second->Reset(num_wait);
}

void WorkerTable::Notify(int msg_id) {
  Waiter* w = nullptr;
  {
    std::lock_guard<std::mutex> lk(mu_);
    auto it = waiting_.find(msg_id);
    if (it == waiting_.end()) return;
    w = it->second.get();
  }
  w->Notify();
}

}  // namespace multiverso
