// Actor runtime (see include/multiverso/actor.h).
#include "multiverso/actor.h"
#include <chrono>
#include "multiverso/util/log.h"
#include "multiverso/zoo.h"

namespace multiverso {

Actor::Actor(const std::string& name) : name_(name) { Zoo::Get()->RegisterActor(name, this); }

Actor::~Actor() {
  if (thread_ && thread_->joinable()) {
    mailbox_.Exit();
    thread_->join();
  }
}

void Actor::Start() {
  thread_.reset(new std::thread([this] {
    is_working_ = true;
    Main();
    is_working_ = false;
  }));
  while (!is_working_) std::this_thread::yield();
}

void Actor::Stop() {
  // let queued and in-flight messages finish, then end the loop
  while (!mailbox_.Empty() || in_flight_.load() > 0)
    std::this_thread::sleep_for(std::chrono::microseconds(200));
  mailbox_.Exit();
  if (thread_ && thread_->joinable()) thread_->join();
  thread_.reset();
}

void Actor::SendTo(const std::string& dst_name, MessagePtr& msg) { Zoo::Get()->SendTo(dst_name, msg); }

void Actor::Dispatch(MessagePtr& msg) {
  auto it = handlers_.find(static_cast<int>(msg->type()));
  if (it == handlers_.end()) it = handlers_.find(static_cast<int>(MsgType::Default));
  if (it == handlers_.end()) {
    Log::Fatal("actor %s: unexpected message type %d", name_.c_str(), static_cast<int>(msg->type()));
    return;
  }
  it->second(msg);
}

void Actor::Main() {
  MessagePtr msg;
  while (mailbox_.Pop(msg)) {
    in_flight_.fetch_add(1);
    Dispatch(msg);
    msg.reset();
    in_flight_.fetch_sub(1);
  }
}

}  // namespace multiverso
