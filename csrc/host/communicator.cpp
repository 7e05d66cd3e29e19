// Communicator actor (see include/multiverso/communicator.h).
#include "multiverso/communicator.h"
#include "multiverso/net.h"
#include "multiverso/util/log.h"
#include "multiverso/zoo.h"

namespace multiverso {

namespace {
// Routing by message-type range (src/communicator.cpp:13-29).
inline bool to_server(int t) { return t > 0 && t < 32; }
inline bool to_worker(int t) { return t < 0 && t > -32; }
inline bool to_controller(int t) { return t > 32; }
}  // namespace

Communicator::Communicator() : Actor(actor::kCommunicator), net_(NetInterface::Get()) {
  RegisterHandler(MsgType::Default, [this](MessagePtr& m) { ProcessMessage(m); });
  recv_thread_.reset(new std::thread([this] { ReceiveLoop(); }));
}

Communicator::~Communicator() { StopReceiver(); }

namespace { constexpr int kPoisonId = -0x7fffffff; }

void Communicator::StopReceiver() {
  if (!recv_thread_) return;
  // Wake the blocking Recv with a message to ourselves; the net stays usable afterwards,
  // so MV_ShutDown(false) + MV_Init works (Test/unittests/multiverso_env.h:15-17).
  if (net_->active()) {
    MessagePtr poison(new Message());
    poison->set_src(net_->rank());
    poison->set_dst(net_->rank());
    poison->set_type(MsgType::Default);
    poison->set_msg_id(kPoisonId);
    net_->Send(poison);
  }
  if (recv_thread_->joinable()) recv_thread_->join();
  recv_thread_.reset();
}

void Communicator::ProcessMessage(MessagePtr& msg) {
  if (msg->dst() != net_->rank()) {
    net_->Send(msg);
    return;
  }
  LocalForward(msg);
}

void Communicator::LocalForward(MessagePtr& msg) {
  const int t = static_cast<int>(msg->type());
  if (to_server(t)) SendTo(actor::kServer, msg);
  else if (to_worker(t)) SendTo(actor::kWorker, msg);
  else if (to_controller(t)) SendTo(actor::kController, msg);
  else Zoo::Get()->Receive(msg);
}

void Communicator::ReceiveLoop() {
  // TcpNet::Recv blocks; it returns -1 once the net is finalized
  for (;;) {
    MessagePtr msg;
    size_t n = net_->Recv(&msg);
    if (n == static_cast<size_t>(-1) || !msg) break;
    if (msg->type() == MsgType::Default && msg->msg_id() == kPoisonId) break;
    LocalForward(msg);
  }
}

}  // namespace multiverso
