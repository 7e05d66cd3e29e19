// Server actors: async Server and BSP SyncServer (src/server.cpp:20-231).
//  * Server      -- applies every Add / serves every Get on arrival.
//  * SyncServer  -- per-worker get/add vector clocks; guarantees that all workers' i-th Get
//                   returns identical parameters computed after every worker's i-th Add(s).
//                   Ahead-of-clock requests are parked and drained when the clock advances;
//                   Server_Finish_Train retires a worker's clocks. -backup_worker_ratio=R
//                   (defined but never read in the reference, Q9) is honoured here: the add
//                   clock advances once ceil((1-R/100) * W) workers have contributed.
#ifndef MULTIVERSO_SERVER_H_
#define MULTIVERSO_SERVER_H_
#include <climits>
#include <list>
#include <vector>
#include "multiverso/actor.h"

namespace multiverso {
class ServerTable;

class Server : public Actor {
 public:
  Server();
  int RegisterTable(ServerTable* table);
  static Server* GetServer();   // Server or SyncServer according to -sync

 protected:
  virtual void ProcessGet(MessagePtr& msg);
  virtual void ProcessAdd(MessagePtr& msg);
  virtual void ProcessFinishTrain(MessagePtr& msg);
  void ServeGet(MessagePtr& msg);
  void ServeAdd(MessagePtr& msg);
  std::vector<ServerTable*> store_;
};

class VectorClock {
 public:
  explicit VectorClock(int n) : local_(n, 0), global_(0) {}
  // Advances worker i; returns true if the global (min) clock moved.
  bool Update(int i);
  bool FinishTrain(int i);
  int local(int i) const { return local_[i]; }
  int global() const { return global_; }
  int size() const { return static_cast<int>(local_.size()); }
  // global clock when only the `quorum` fastest... slowest-excluded workers count
  int KthLargest(int k) const;

 private:
  int Min() const;
  std::vector<int> local_;
  int global_;
};

class SyncServer : public Server {
 public:
  SyncServer();

 protected:
  void ProcessGet(MessagePtr& msg) override;
  void ProcessAdd(MessagePtr& msg) override;
  void ProcessFinishTrain(MessagePtr& msg) override;

 private:
  void DrainGets();
  void DrainAdds();
  int AddFrontier() const;   // add-clock value every Get must not outrun
  VectorClock get_clock_, add_clock_;
  int quorum_;
  std::vector<int> pending_adds_;             // parked Adds per worker
  std::list<MessagePtr> parked_add_, parked_get_;
};

}  // namespace multiverso
#endif
