// Controller actor, rank 0 (see include/multiverso/controller.h).
#include "multiverso/controller.h"
#include "multiverso/util/log.h"
#include "multiverso/zoo.h"

namespace multiverso {

Controller::Controller() : Actor(actor::kController) {
  RegisterHandler(MsgType::Control_Barrier, [this](MessagePtr& m) { OnBarrierArrival(m); });
  RegisterHandler(MsgType::Control_Register, [this](MessagePtr& m) { OnRegistration(m); });
}

// Own rank last, so the local caller cannot race ahead of the remote replies being queued.
void Controller::ReleaseAll(std::vector<MessagePtr>* parked, const std::vector<Blob>& payload) {
  MessagePtr mine;
  for (auto& m : *parked) {
    MessagePtr reply(m->CreateReplyMessage());
    for (const Blob& b : payload) reply->Push(b);
    if (reply->dst() == Zoo::Get()->rank()) mine = std::move(reply);
    else SendTo(actor::kCommunicator, reply);
  }
  if (mine) SendTo(actor::kCommunicator, mine);
  parked->clear();
}

void Controller::OnBarrierArrival(MessagePtr& msg) {
  at_barrier_.push_back(std::move(msg));
  if (static_cast<int>(at_barrier_.size()) == Zoo::Get()->size()) ReleaseAll(&at_barrier_, {});
}

// Worker / server ids are dense in RANK order (the reference assigns them in arrival order,
// src/controller.cpp:51-54, which makes ids nondeterministic across runs); the broadcast is
// Node[size] + {num_workers, num_servers}.
void Controller::OnRegistration(MessagePtr& msg) {
  const int size = Zoo::Get()->size();
  if (roster_.empty()) roster_.assign(size, Node());
  const Node n = msg->data()[0].As<Node>(0);
  roster_[n.rank] = n;
  registering_.push_back(std::move(msg));
  if (static_cast<int>(registering_.size()) < size) return;
  int counts[2] = {0, 0};   // workers, servers
  for (Node& node : roster_) {
    node.worker_id = node::is_worker(node.role) ? counts[0]++ : -1;
    node.server_id = node::is_server(node.role) ? counts[1]++ : -1;
  }
  ReleaseAll(&registering_, {Blob(roster_.data(), sizeof(Node) * size), Blob(counts, sizeof counts)});
  roster_.clear();
}

}  // namespace multiverso
