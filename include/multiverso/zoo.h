// Zoo: the per-process system singleton (counterpart of include/multiverso/zoo.h,
// src/zoo.cpp:41-186): flag parsing, net init, role selection (-ps_role), ordered actor
// start-up (controller on rank 0 -> communicator -> register -> server -> worker -> barrier),
// intra-process routing by actor name, Barrier, table registration, id <-> rank maps.
#ifndef MULTIVERSO_ZOO_H_
#define MULTIVERSO_ZOO_H_
#include <map>
#include <memory>
#include <string>
#include <vector>
#include "multiverso/message.h"
#include "multiverso/node.h"
#include "multiverso/util/mt_queue.h"

namespace multiverso {

class Actor;
class NetInterface;
class WorkerTable;
class ServerTable;

class Zoo {
 public:
  static Zoo* Get();
  void Start(int* argc, char** argv);
  void Stop(bool finalize_net);
  void Barrier();

  void SendTo(const std::string& actor_name, MessagePtr& msg);
  void Receive(MessagePtr& msg) { mailbox_.Push(std::move(msg)); }

  int rank() const;
  int size() const;
  int worker_rank() const { return started_ ? nodes_[rank()].worker_id : -1; }
  int server_rank() const { return started_ ? nodes_[rank()].server_id : -1; }
  int num_workers() const { return num_workers_; }
  int num_servers() const { return num_servers_; }
  int worker_id_to_rank(int id) const { return worker_id_to_rank_.at(id); }
  int server_id_to_rank(int id) const { return server_id_to_rank_.at(id); }
  int rank_to_worker_id(int r) const { return nodes_[r].worker_id; }
  int rank_to_server_id(int r) const { return nodes_[r].server_id; }
  bool started() const { return started_; }
  bool model_average() const { return ma_mode_; }

  int RegisterTable(WorkerTable* worker_table);
  int RegisterTable(ServerTable* server_table);
  void RegisterActor(const std::string& name, Actor* actor) { actors_[name] = actor; }

 private:
  Zoo() = default;
  void StartPS();
  void StopPS();
  void RegisterNode();
  void FinishTrain();

  bool started_ = false;
  bool ma_mode_ = false;
  bool ps_running_ = false;
  std::map<std::string, Actor*> actors_;
  std::vector<std::unique_ptr<Actor>> owned_;
  MtQueue<MessagePtr> mailbox_;
  std::vector<Node> nodes_;
  std::vector<int> worker_id_to_rank_, server_id_to_rank_;
  int num_workers_ = 0, num_servers_ = 0;
  NetInterface* net_ = nullptr;
};

}  // namespace multiverso
#endif
