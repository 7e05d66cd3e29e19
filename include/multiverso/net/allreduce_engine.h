// AllreduceEngine: hand-rolled host collectives on NetInterface::SendRecv (counterpart of
// include/multiverso/net/allreduce_engine.h + src/net/allreduce_engine.cpp/allreduce_topo.cpp,
// which the reference never compiles on Linux nor calls). Always available here:
//   Allgather      -- Bruck: ceil(log2 N) rounds, block i goes to rank (r - 2^k), final rotate
//   ReduceScatter  -- recursive halving; non-power-of-two sizes fold the extra ranks into a
//                     partner first and unfold at the end
//   Allreduce      -- small: allgather + local reduce; large: reduce-scatter + allgather
#ifndef MULTIVERSO_NET_ALLREDUCE_ENGINE_H_
#define MULTIVERSO_NET_ALLREDUCE_ENGINE_H_
#include <cstddef>
#include <functional>
#include <vector>

namespace multiverso {

class NetInterface;

using ReduceFunction = std::function<void(const char* src, char* dst, int len_bytes)>;

struct BruckStep {
  int send_to, recv_from, blocks;   // number of blocks exchanged in this round
};
std::vector<BruckStep> BruckSchedule(int rank, int size);

struct HalvingStep {
  int peer;
  int send_lo, send_hi;   // block range sent to the peer
  int keep_lo, keep_hi;   // block range kept (and received into)
};
std::vector<HalvingStep> RecursiveHalvingSchedule(int vrank, int pow2);

class AllreduceEngine {
 public:
  explicit AllreduceEngine(NetInterface* net);
  // in-place allreduce of `count` elements of `type_size` bytes
  void Allreduce(char* data, int count, int type_size, const ReduceFunction& reducer);
  // every rank contributes block_len bytes; out holds size*block_len bytes in rank order
  void Allgather(const char* in, int block_len, char* out);
  // data: size blocks (block_start/len in bytes); on return this rank's block is reduced
  void ReduceScatter(char* data, const std::vector<int>& block_start,
                     const std::vector<int>& block_len, const ReduceFunction& reducer);
  // variable-size allgather used after ReduceScatter
  void AllgatherV(char* data, const std::vector<int>& block_start, const std::vector<int>& block_len);

 private:
  NetInterface* net_;
  int rank_, size_;
};

}  // namespace multiverso
#endif
