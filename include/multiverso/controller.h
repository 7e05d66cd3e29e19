// Controller actor (rank 0 only): node registration -> dense worker/server ids, and the
// global barrier (src/controller.cpp:12-102).
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
  void ProcessBarrier(MessagePtr& msg);
  void ProcessRegister(MessagePtr& msg);
  std::vector<MessagePtr> barrier_waiting_;
  std::vector<MessagePtr> register_waiting_;
  std::vector<Node> nodes_;
};
}  // namespace multiverso
#endif
