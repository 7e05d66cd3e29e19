// NetInterface: the control-plane / host-backend transport (counterpart of
// include/multiverso/net.h:9-49). The reference selects MPI or ZeroMQ at compile time;
// neither exists here, so the single implementation is TcpNet (full-mesh TCP sockets,
// THREAD_MULTIPLE) with the same Init / Bind / Connect / Send / Recv / SendTo / RecvFrom /
// SendRecv surface. On the GPU data path there is no NetInterface at all (peer-mapped HBM).
#ifndef MULTIVERSO_NET_H_
#define MULTIVERSO_NET_H_
#include <cstddef>
#include <string>
#include "multiverso/message.h"

namespace multiverso {

enum NetThreadLevel { THREAD_SERIALIZED = 0, THREAD_MULTIPLE = 1 };

class NetInterface {
 public:
  static NetInterface* Get();
  virtual ~NetInterface() = default;
  virtual void Init(int* argc = nullptr, char** argv = nullptr) = 0;
  virtual void Finalize() = 0;
  // Explicit-endpoint bootstrap (MV_NetBind / MV_NetConnect, used by the C# binding).
  virtual int Bind(int rank, char* endpoint) = 0;
  virtual int Connect(int* ranks, char* endpoints[], int size) = 0;
  virtual bool active() const = 0;
  virtual std::string name() const = 0;
  virtual int rank() const = 0;
  virtual int size() const = 0;
  // Message transport. Send returns the number of bytes written; Recv blocks until a
  // message arrives (returns its size) or the net is finalized (returns -1).
  virtual size_t Send(MessagePtr& msg) = 0;
  virtual size_t Recv(MessagePtr* msg) = 0;
  // Raw byte transport for the allreduce engine.
  virtual void SendTo(int rank, const char* buf, int len) = 0;
  virtual void RecvFrom(int rank, char* buf, int len) = 0;
  virtual void SendRecv(int send_rank, const char* send_buf, int send_len, int recv_rank,
                        char* recv_buf, int recv_len) = 0;
  virtual int thread_level_support() = 0;
};

namespace net {
// In-place SUM all-reduce over all ranks (MV_Aggregate). T in {char,int,float,double}.
template <typename T>
void Allreduce(T* data, size_t count);
}  // namespace net

}  // namespace multiverso
#endif
