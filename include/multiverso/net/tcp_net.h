// TcpNet: full-mesh TCP implementation of NetInterface. Replaces both MPINetWrapper
// (include/multiverso/net/mpi_net.h) and ZMQNetWrapper (include/multiverso/net/zmq_net.h).
// Bootstrap, in priority order:
//   1. explicit Bind(rank, "ip:port") + Connect(ranks, endpoints, n)     (ZMQ-style, C# path)
//   2. -machine_file=<one ip per line> and -port (own rank = line matching a local NIC)
//   3. environment: MV_RANK/MV_SIZE or RANK/WORLD_SIZE, MASTER_ADDR, MV_PORT or MASTER_PORT+64
//   4. nothing set: single process, size 1 (loop-back only)
// Frames: [u32 kind][u32 src][u64 payload_len][payload]; kind 0 = Message, 1 = raw bytes.
#ifndef MULTIVERSO_NET_TCP_NET_H_
#define MULTIVERSO_NET_TCP_NET_H_
#include <atomic>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include "multiverso/net.h"
#include "multiverso/util/mt_queue.h"

namespace multiverso {

class TcpNet : public NetInterface {
 public:
  TcpNet();
  ~TcpNet() override;
  void Init(int* argc, char** argv) override;
  void Finalize() override;
  int Bind(int rank, char* endpoint) override;
  int Connect(int* ranks, char* endpoints[], int size) override;
  bool active() const override { return active_; }
  std::string name() const override { return "TCP"; }
  int rank() const override { return rank_; }
  int size() const override { return size_; }
  size_t Send(MessagePtr& msg) override;
  size_t Recv(MessagePtr* msg) override;
  void SendTo(int rank, const char* buf, int len) override;
  void RecvFrom(int rank, char* buf, int len) override;
  void SendRecv(int send_rank, const char* send_buf, int send_len, int recv_rank, char* recv_buf,
                int recv_len) override;
  int thread_level_support() override { return THREAD_MULTIPLE; }

  static bool ParseEndpoint(const std::string& ep, std::string* host, int* port);

 private:
  struct RawChunk {
    std::vector<char> bytes;
    size_t consumed = 0;
  };
  void EstablishMesh(const std::vector<std::string>& endpoints);
  void ReceiverLoop();
  void WriteFrame(int dst, uint32_t kind, const std::vector<std::pair<const void*, size_t>>& parts);
  bool ReadExact(int fd, void* buf, size_t n);

  bool active_ = false;
  int rank_ = 0, size_ = 1;
  int listen_fd_ = -1;
  std::string bound_endpoint_;
  std::vector<int> fds_;                        // socket per peer (-1 for self)
  std::vector<std::unique_ptr<std::mutex>> send_mu_;
  MtQueue<MessagePtr> inbox_;
  std::vector<std::unique_ptr<MtQueue<std::shared_ptr<RawChunk>>>> raw_in_;
  std::vector<std::shared_ptr<RawChunk>> raw_partial_;
  std::thread receiver_;
  std::atomic<bool> stopping_{false};
};

}  // namespace multiverso
#endif
