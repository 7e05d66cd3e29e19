// Controller actor (rank 0 only): the two rendezvous of the control plane. Registration gathers
// one Node per rank, hands out dense worker / server ids in rank order and broadcasts the table;
// the barrier gathers one message per rank and releases everybody (reference behaviour:
// src/controller.cpp:12-102).
#ifndef MULTIVERSO_CONTROLLER_H_
#define MULTIVERSO_CONTROLLER_H_
#include <vector>
#include "multiverso/actor.h"
#include "multiverso/node.h"

namespace multiverso {
class Controller : public Actor {
 public:
  Controller();

 private:
  void OnBarrierArrival(MessagePtr& msg);
  void OnRegistration(MessagePtr& msg);
  // Reply to every parked request; this rank's own reply goes out last.
  void ReleaseAll(std::vector<MessagePtr>* parked, const std::vector<Blob>& payload);
  std::vector<MessagePtr> at_barrier_;
  std::vector<MessagePtr> registering_;
  std::vector<Node> roster_;
};
}  // namespace multiverso
#endif
