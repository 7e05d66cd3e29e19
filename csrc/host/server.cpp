// Server / SyncServer actors (see include/multiverso/server.h).
#include "multiverso/server.h"
#include <algorithm>
#include <cmath>
#include "multiverso/dashboard.h"
#include "multiverso/table_interface.h"
#include "multiverso/util/configure.h"
#include "multiverso/util/log.h"
#include "multiverso/zoo.h"

namespace multiverso {

MV_DEFINE_bool(sync, false, "sync or async");
MV_DEFINE_int(backup_worker_ratio, 0, "ratio% of backup workers, set 20 means 20%");

Server::Server() : Actor(actor::kServer) {
  RegisterHandler(MsgType::Request_Get, [this](MessagePtr& m) { ProcessGet(m); });
  RegisterHandler(MsgType::Request_Add, [this](MessagePtr& m) { ProcessAdd(m); });
  RegisterHandler(MsgType::Server_Finish_Train, [this](MessagePtr& m) { ProcessFinishTrain(m); });
}

int Server::RegisterTable(ServerTable* table) {
  store_.push_back(table);
  return static_cast<int>(store_.size()) - 1;
}

void Server::ServeGet(MessagePtr& msg) {
  MONITOR_BEGIN(SERVER_PROCESS_GET)
  MessagePtr reply(msg->CreateReplyMessage());
  store_.at(msg->table_id())->ProcessGet(msg->data(), &reply->data());
  SendTo(actor::kCommunicator, reply);
  MONITOR_END(SERVER_PROCESS_GET)
}

void Server::ServeAdd(MessagePtr& msg) {
  MONITOR_BEGIN(SERVER_PROCESS_ADD)
  MessagePtr reply(msg->CreateReplyMessage());
  store_.at(msg->table_id())->ProcessAdd(msg->data());
  SendTo(actor::kCommunicator, reply);
  MONITOR_END(SERVER_PROCESS_ADD)
}

void Server::ProcessGet(MessagePtr& msg) { ServeGet(msg); }
void Server::ProcessAdd(MessagePtr& msg) { ServeAdd(msg); }
void Server::ProcessFinishTrain(MessagePtr&) {}

Server* Server::GetServer() {
  if (MV_CONFIG(sync)) {
    Log::Info("Create a sync server");
    return new SyncServer();
  }
  Log::Debug("Create an async server");
  return new Server();
}

// ---------------------------------------------------------------------------------------
bool VectorClock::Update(int i) {
  if (local_[i] != INT_MAX) ++local_[i];
  int m = Min();
  if (m > global_) {
    global_ = m;
    return true;
  }
  return false;
}
bool VectorClock::FinishTrain(int i) {
  local_[i] = INT_MAX;
  int m = Min();
  if (m > global_) {
    global_ = m;
    return true;
  }
  return false;
}
int VectorClock::Min() const { return *std::min_element(local_.begin(), local_.end()); }
int VectorClock::KthLargest(int k) const {
  std::vector<int> v(local_);
  std::sort(v.begin(), v.end(), std::greater<int>());
  return v[std::min<int>(k, static_cast<int>(v.size())) - 1];
}

SyncServer::SyncServer()
    : get_clock_(Zoo::Get()->num_workers()), add_clock_(Zoo::Get()->num_workers()),
      pending_adds_(Zoo::Get()->num_workers(), 0) {
  const int W = Zoo::Get()->num_workers();
  const int ratio = std::max(0, std::min(99, MV_CONFIG(backup_worker_ratio)));
  quorum_ = std::max(1, static_cast<int>(std::ceil((100 - ratio) / 100.0 * W)));
  if (ratio > 0) Log::Info("SyncServer: %d of %d workers form the quorum (backup_worker_ratio=%d)", quorum_, W, ratio);
}

// The add-clock value up to which parameters are "complete": all workers (or the quorum of
// fastest workers when backup workers are configured) have contributed that many Adds.
int SyncServer::AddFrontier() const { return add_clock_.KthLargest(quorum_); }

void SyncServer::ProcessAdd(MessagePtr& msg) {
  const int w = Zoo::Get()->rank_to_worker_id(msg->src());
  // A worker that already finished its i-th Get while others have not may not change the
  // parameters those others still have to read: park its next-iteration Add.
  if (get_clock_.local(w) > get_clock_.KthLargest(quorum_) || pending_adds_[w] > 0) {
    ++pending_adds_[w];
    parked_add_.push_back(std::move(msg));
    return;
  }
  ServeAdd(msg);
  add_clock_.Update(w);
  DrainGets();
}

void SyncServer::ProcessGet(MessagePtr& msg) {
  const int w = Zoo::Get()->rank_to_worker_id(msg->src());
  // A Get must observe the Adds of every (quorum) worker up to this worker's own add count.
  if (pending_adds_[w] > 0 || add_clock_.local(w) > AddFrontier()) {
    parked_get_.push_back(std::move(msg));
    return;
  }
  ServeGet(msg);
  get_clock_.Update(w);
  DrainAdds();
}

void SyncServer::ProcessFinishTrain(MessagePtr& msg) {
  const int w = Zoo::Get()->rank_to_worker_id(msg->src());
  add_clock_.FinishTrain(w);
  get_clock_.FinishTrain(w);
  DrainAdds();
  DrainGets();
}

void SyncServer::DrainGets() {
  bool progress = true;
  while (progress) {
    progress = false;
    std::vector<char> blocked(add_clock_.size(), 0);   // keep per-worker FIFO order
    for (auto it = parked_get_.begin(); it != parked_get_.end();) {
      const int w = Zoo::Get()->rank_to_worker_id((*it)->src());
      if (!blocked[w] && pending_adds_[w] == 0 && add_clock_.local(w) <= AddFrontier()) {
        MessagePtr m = std::move(*it);
        it = parked_get_.erase(it);
        ServeGet(m);
        get_clock_.Update(w);
        progress = true;
      } else {
        blocked[w] = 1;
        ++it;
      }
    }
    if (progress) DrainAdds();
  }
}

void SyncServer::DrainAdds() {
  bool progress = true;
  while (progress) {
    progress = false;
    std::vector<char> blocked(add_clock_.size(), 0);
    for (auto it = parked_add_.begin(); it != parked_add_.end();) {
      const int w = Zoo::Get()->rank_to_worker_id((*it)->src());
      if (!blocked[w] && get_clock_.local(w) <= get_clock_.KthLargest(quorum_)) {
        MessagePtr m = std::move(*it);
        it = parked_add_.erase(it);
        --pending_adds_[w];
        ServeAdd(m);
        add_clock_.Update(w);
        progress = true;
      } else {
        blocked[w] = 1;
        ++it;
      }
    }
  }
  // newly applied Adds may release parked Gets
  if (!parked_get_.empty()) {
    for (auto& g : parked_get_) {
      const int w = Zoo::Get()->rank_to_worker_id(g->src());
      if (pending_adds_[w] == 0 && add_clock_.local(w) <= AddFrontier()) {
        DrainGets();
        break;
      }
    }
  }
}

}  // namespace multiverso
