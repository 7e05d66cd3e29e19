// Communicator actor: bridge between local actors and NetInterface (src/communicator.cpp).
// TcpNet is THREAD_MULTIPLE, so: the actor thread sends, a second thread receives and
// forwards by message-type range: (0,32) -> server, (-32,0) -> worker, >32 -> controller,
// anything else -> the Zoo mailbox.
#ifndef MULTIVERSO_COMMUNICATOR_H_
#define MULTIVERSO_COMMUNICATOR_H_
#include <memory>
#include <thread>
#include "multiverso/actor.h"

namespace multiverso {
class NetInterface;

class Communicator : public Actor {
 public:
  Communicator();
  ~Communicator() override;
  void StopReceiver();

 private:
  void ProcessMessage(MessagePtr& msg);
  void LocalForward(MessagePtr& msg);
  void ReceiveLoop();
  NetInterface* net_;
  std::unique_ptr<std::thread> recv_thread_;
};
}  // namespace multiverso
#endif
