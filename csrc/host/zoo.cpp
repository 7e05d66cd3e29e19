// Zoo (see include/multiverso/zoo.h).
#include "multiverso/zoo.h"
#include "multiverso/actor.h"
#include "multiverso/communicator.h"
#include "multiverso/controller.h"
#include "multiverso/dashboard.h"
#include "multiverso/net.h"
#include "multiverso/server.h"
#include "multiverso/table_interface.h"
#include "multiverso/util/configure.h"
#include "multiverso/util/log.h"
#include "multiverso/worker.h"

namespace multiverso {

MV_DEFINE_string(ps_role, "default", "none / worker / server / default");
MV_DEFINE_bool(ma, false, "model average, will not start server if true");
MV_DECLARE_bool(sync);

namespace {
int ParsePSRole(const std::string& r) {
  if (r == "none") return Role::NONE;
  if (r == "worker") return Role::WORKER;
  if (r == "server") return Role::SERVER;
  return Role::ALL;
}
}  // namespace

Zoo* Zoo::Get() {
  static Zoo* zoo = new Zoo();   // leaked: actor threads may outlive static destruction
  return zoo;
}

int Zoo::rank() const { return net_ ? net_->rank() : 0; }
int Zoo::size() const { return net_ ? net_->size() : 1; }

void Zoo::Start(int* argc, char** argv) {
  Log::Debug("Zoo started");
  ParseCMDFlags(argc, argv);
  net_ = NetInterface::Get();
  net_->Init(argc, argv);
  ma_mode_ = MV_CONFIG(ma);
  nodes_.assign(size(), Node());
  if (!ma_mode_) StartPS();
  started_ = true;
}

void Zoo::Stop(bool finalize_net) {
  if (!started_) return;
  if (ps_running_) StopPS();
  started_ = false;
  if (finalize_net && net_) net_->Finalize();
}

void Zoo::StartPS() {
  const int role = ParsePSRole(MV_CONFIG(ps_role));
  Log::Debug("Rank %d: start PS with role %d", rank(), role);
  // Order matters: the controller must exist before anybody registers, the communicator
  // before any message leaves, server before worker so tables find both halves.
  if (rank() == 0) {
    owned_.emplace_back(new Controller());
    owned_.back()->Start();
  }
  owned_.emplace_back(new Communicator());
  owned_.back()->Start();
  nodes_[rank()].rank = rank();
  nodes_[rank()].role = role;
  RegisterNode();
  if (node::is_server(role)) {
    owned_.emplace_back(Server::GetServer());
    owned_.back()->Start();
  }
  if (node::is_worker(role)) {
    owned_.emplace_back(new Worker());
    owned_.back()->Start();
  }
  ps_running_ = true;
  started_ = true;
  Barrier();
  Log::Debug("Rank %d: multiverso started, worker id %d server id %d", rank(), worker_rank(), server_rank());
}

void Zoo::StopPS() {
  if (MV_CONFIG(sync)) FinishTrain();
  Barrier();
  // Stop in reverse order of creation; the communicator's receive thread ends when the net
  // is finalized (or immediately for a single process).
  for (auto it = owned_.rbegin(); it != owned_.rend(); ++it) (*it)->Stop();
  ps_running_ = false;
  for (auto& a : owned_)
    if (auto* c = dynamic_cast<Communicator*>(a.get())) c->StopReceiver();
  owned_.clear();
  actors_.clear();
  // drain anything left in the Zoo mailbox (e.g. the poison message)
  MessagePtr junk;
  while (mailbox_.TryPop(junk)) {}
}

void Zoo::RegisterNode() {
  MessagePtr msg(new Message());
  msg->set_src(rank());
  msg->set_dst(0);
  msg->set_type(MsgType::Control_Register);
  msg->Push(Blob(&nodes_[rank()], sizeof(Node)));
  SendTo(actor::kCommunicator, msg);
  MessagePtr reply;
  for (;;) {
    CHECK(mailbox_.Pop(reply));
    if (reply->type() == MsgType::Control_Reply_Register) break;
  }
  CHECK(reply->data().size() == 2);
  const Blob& table = reply->data()[0];
  CHECK(table.size<Node>() == static_cast<size_t>(size()));
  num_workers_ = reply->data()[1].As<int>(0);
  num_servers_ = reply->data()[1].As<int>(1);
  worker_id_to_rank_.assign(num_workers_, -1);
  server_id_to_rank_.assign(num_servers_, -1);
  for (int r = 0; r < size(); ++r) {
    nodes_[r] = table.As<Node>(r);
    if (nodes_[r].worker_id >= 0) worker_id_to_rank_[nodes_[r].worker_id] = r;
    if (nodes_[r].server_id >= 0) server_id_to_rank_[nodes_[r].server_id] = r;
  }
}

void Zoo::FinishTrain() {
  if (worker_rank() < 0) return;
  for (int s = 0; s < num_servers_; ++s) {
    MessagePtr msg(new Message());
    msg->set_src(rank());
    msg->set_dst(server_id_to_rank(s));
    msg->set_type(MsgType::Server_Finish_Train);
    SendTo(actor::kCommunicator, msg);
  }
}

void Zoo::Barrier() {
  if (!ps_running_) {
    // model-averaging mode has no actors: barrier = a 1-element all-reduce
    if (net_ && net_->active() && size() > 1) {
      int one = 1;
      net::Allreduce<int>(&one, 1);
    }
    return;
  }
  MessagePtr msg(new Message());
  msg->set_src(rank());
  msg->set_dst(0);
  msg->set_type(MsgType::Control_Barrier);
  SendTo(actor::kCommunicator, msg);
  MessagePtr reply;
  for (;;) {
    CHECK(mailbox_.Pop(reply));
    if (reply->type() == MsgType::Control_Reply_Barrier) break;
  }
}

void Zoo::SendTo(const std::string& name, MessagePtr& msg) {
  auto it = actors_.find(name);
  if (it == actors_.end()) {
    Log::Fatal("Zoo::SendTo: no actor named '%s' on rank %d (message type %d)", name.c_str(), rank(),
               static_cast<int>(msg->type()));
    return;
  }
  it->second->Receive(msg);
}

int Zoo::RegisterTable(WorkerTable* t) {
  return dynamic_cast<Worker*>(actors_.at(actor::kWorker))->RegisterTable(t);
}
int Zoo::RegisterTable(ServerTable* t) {
  return dynamic_cast<Server*>(actors_.at(actor::kServer))->RegisterTable(t);
}

}  // namespace multiverso
