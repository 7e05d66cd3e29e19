// Actor: a named mailbox + handler table + thread (counterpart of include/multiverso/actor.h:
// 18-67). Start() returns once the thread is running; Stop() drains the mailbox, then joins.
#ifndef MULTIVERSO_ACTOR_H_
#define MULTIVERSO_ACTOR_H_
#include <atomic>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <thread>
#include "multiverso/message.h"
#include "multiverso/util/mt_queue.h"

namespace multiverso {

namespace actor {
const std::string kCommunicator = "communicator";
const std::string kController = "controller";
const std::string kServer = "server";
const std::string kWorker = "worker";
}  // namespace actor

class Actor {
 public:
  explicit Actor(const std::string& name);
  virtual ~Actor();
  void Start();
  void Stop();
  // Enqueue a message for this actor (any thread).
  void Receive(MessagePtr& msg) { mailbox_.Push(std::move(msg)); }
  const std::string& name() const { return name_; }

 protected:
  using Handler = std::function<void(MessagePtr&)>;
  void RegisterHandler(MsgType type, Handler h) { handlers_[static_cast<int>(type)] = std::move(h); }
  // Hand a message to another actor of this process (through the Zoo).
  void SendTo(const std::string& dst_name, MessagePtr& msg);
  virtual void Main();
  void Dispatch(MessagePtr& msg);

  std::string name_;
  MtQueue<MessagePtr> mailbox_;
  std::map<int, Handler> handlers_;
  std::unique_ptr<std::thread> thread_;
  std::atomic<bool> is_working_{false};
  std::atomic<int> in_flight_{0};
};

}  // namespace multiverso
#endif
