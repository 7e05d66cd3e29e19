// Controller actor, rank 0 (see include/multiverso/controller.h).
#include "multiverso/controller.h"
#include "multiverso/util/log.h"
#include "multiverso/zoo.h"

namespace multiverso {

Controller::Controller() : Actor(actor::kController) {
  RegisterHandler(MsgType::Control_Barrier, [this](MessagePtr& m) { ProcessBarrier(m); });
  RegisterHandler(MsgType::Control_Register, [this](MessagePtr& m) { ProcessRegister(m); });
}

// Collect one barrier message per rank, then release everybody (own rank last, so the local
// caller cannot race ahead of the remote replies being queued).
void Controller::ProcessBarrier(MessagePtr& msg) {
  barrier_waiting_.push_back(std::move(msg));
  if (static_cast<int>(barrier_waiting_.size()) < Zoo::Get()->size()) return;
  MessagePtr mine;
  for (auto& m : barrier_waiting_) {
    MessagePtr reply(m->CreateReplyMessage());
    if (reply->dst() == Zoo::Get()->rank()) mine = std::move(reply);
    else SendTo(actor::kCommunicator, reply);
  }
  if (mine) SendTo(actor::kCommunicator, mine);
  barrier_waiting_.clear();
}

// Collect one Node per rank; worker / server ids are dense in RANK order (the reference
// assigns them in arrival order, src/controller.cpp:51-54, which makes ids nondeterministic
// across runs); broadcast Node[size] + {num_workers, num_servers}.
void Controller::ProcessRegister(MessagePtr& msg) {
  const int size = Zoo::Get()->size();
  if (nodes_.empty()) nodes_.assign(size, Node());
  Node n = msg->data()[0].As<Node>(0);
  nodes_[n.rank] = n;
  register_waiting_.push_back(std::move(msg));
  if (static_cast<int>(register_waiting_.size()) < size) return;
  int nw = 0, ns = 0;
  for (int r = 0; r < size; ++r) {
    nodes_[r].worker_id = node::is_worker(nodes_[r].role) ? nw++ : -1;
    nodes_[r].server_id = node::is_server(nodes_[r].role) ? ns++ : -1;
  }
  Blob table(nodes_.data(), sizeof(Node) * size);
  int counts[2] = {nw, ns};
  Blob cnt(counts, sizeof counts);
  MessagePtr mine;
  for (auto& m : register_waiting_) {
    MessagePtr reply(m->CreateReplyMessage());
    reply->Push(table);
    reply->Push(cnt);
    if (reply->dst() == Zoo::Get()->rank()) mine = std::move(reply);
    else SendTo(actor::kCommunicator, reply);
  }
  if (mine) SendTo(actor::kCommunicator, mine);
  register_waiting_.clear();
  nodes_.clear();
}

}  // namespace multiverso
