// Host collectives on NetInterface::SendRecv (see include/multiverso/net/allreduce_engine.h).
#include "multiverso/net/allreduce_engine.h"

#include <algorithm>
#include <cstring>

#include "multiverso/net.h"
#include "multiverso/util/log.h"

namespace multiverso {

// Bruck all-gather: after round k (distance d = 2^k) every rank holds the blocks of ranks
// r, r+1, ..., r+2d-1 (mod N) in its local order; a final rotation restores rank order.
std::vector<BruckStep> BruckSchedule(int rank, int size) {
  std::vector<BruckStep> steps;
  for (int d = 1; d < size; d <<= 1) {
    BruckStep s;
    s.send_to = (rank - d + size) % size;
    s.recv_from = (rank + d) % size;
    s.blocks = std::min(d, size - d);
    steps.push_back(s);
  }
  return steps;
}

// Recursive halving among `pow2` virtual ranks over `pow2` block groups: at each level the
// group [lo, hi) splits in two; a rank keeps the half containing its own index and sends the
// other half to the partner at distance half.
std::vector<HalvingStep> RecursiveHalvingSchedule(int vrank, int pow2) {
  std::vector<HalvingStep> steps;
  int lo = 0, hi = pow2;
  while (hi - lo > 1) {
    int mid = (lo + hi) / 2, half = (hi - lo) / 2;
    HalvingStep s;
    if (vrank < mid) {
      s.peer = vrank + half;
      s.keep_lo = lo; s.keep_hi = mid; s.send_lo = mid; s.send_hi = hi;
      hi = mid;
    } else {
      s.peer = vrank - half;
      s.keep_lo = mid; s.keep_hi = hi; s.send_lo = lo; s.send_hi = mid;
      lo = mid;
    }
    steps.push_back(s);
  }
  return steps;
}

AllreduceEngine::AllreduceEngine(NetInterface* net)
    : net_(net), rank_(net->rank()), size_(net->size()) {}

void AllreduceEngine::Allgather(const char* in, int block_len, char* out) {
  if (size_ == 1) {
    std::memcpy(out, in, block_len);
    return;
  }
  // local order: slot j holds the block of rank (rank_ + j) % size_
  std::vector<char> tmp(static_cast<size_t>(block_len) * size_);
  std::memcpy(tmp.data(), in, block_len);
  int have = 1;
  for (const BruckStep& s : BruckSchedule(rank_, size_)) {
    net_->SendRecv(s.send_to, tmp.data(), s.blocks * block_len, s.recv_from,
                   tmp.data() + static_cast<size_t>(have) * block_len, s.blocks * block_len);
    have += s.blocks;
  }
  for (int j = 0; j < size_; ++j)
    std::memcpy(out + static_cast<size_t>((rank_ + j) % size_) * block_len,
                tmp.data() + static_cast<size_t>(j) * block_len, block_len);
}

void AllreduceEngine::AllgatherV(char* data, const std::vector<int>& start,
                                 const std::vector<int>& len) {
  // ring all-gather of variable-size blocks (N-1 steps, each rank forwards what it got last)
  for (int step = 0; step < size_ - 1; ++step) {
    int send_block = (rank_ - step + size_) % size_;
    int recv_block = (rank_ - step - 1 + size_) % size_;
    net_->SendRecv((rank_ + 1) % size_, data + start[send_block], len[send_block],
                   (rank_ - 1 + size_) % size_, data + start[recv_block], len[recv_block]);
  }
}

void AllreduceEngine::ReduceScatter(char* data, const std::vector<int>& start,
                                    const std::vector<int>& len, const ReduceFunction& reducer) {
  if (size_ == 1) return;
  int pow2 = 1;
  while (pow2 * 2 <= size_) pow2 <<= 1;
  const int extra = size_ - pow2;           // ranks [pow2, size_) fold into ranks [0, extra)
  const int total = start[size_ - 1] + len[size_ - 1];
  std::vector<char> recv(static_cast<size_t>(total));
  if (rank_ >= pow2) {
    // extra rank: hand everything to the partner, get the own reduced block back at the end
    int partner = rank_ - pow2;
    net_->SendTo(partner, data, total);
    net_->RecvFrom(partner, data + start[rank_], len[rank_]);
    return;
  }
  if (rank_ < extra) {
    net_->RecvFrom(rank_ + pow2, recv.data(), total);
    reducer(recv.data(), data, total);
  }
  // block groups: group g (< pow2) = real blocks {g} plus {g + pow2} when g < extra
  auto group_lo = [&](int g) { return start[g]; };
  auto span = [&](int glo, int ghi, std::vector<std::pair<int, int>>* out) {
    out->clear();
    out->emplace_back(group_lo(glo), start[ghi - 1] + len[ghi - 1] - start[glo]);   // [glo, ghi)
    int elo = glo + pow2, ehi = std::min(ghi, extra) + pow2;
    if (glo < extra && elo < ehi) out->emplace_back(start[elo], start[ehi - 1] + len[ehi - 1] - start[elo]);
  };
  std::vector<std::pair<int, int>> send_spans, keep_spans;
  for (const HalvingStep& s : RecursiveHalvingSchedule(rank_, pow2)) {
    span(s.send_lo, s.send_hi, &send_spans);
    span(s.keep_lo, s.keep_hi, &keep_spans);
    for (size_t i = 0; i < 2; ++i) {
      int slen = i < send_spans.size() ? send_spans[i].second : 0;
      int klen = i < keep_spans.size() ? keep_spans[i].second : 0;
      const char* sp = slen ? data + send_spans[i].first : nullptr;
      char* kp = klen ? recv.data() + keep_spans[i].first : nullptr;
      if (slen || klen) net_->SendRecv(s.peer, sp, slen, s.peer, kp, klen);
      if (klen) reducer(kp, data + keep_spans[i].first, klen);
    }
  }
  if (rank_ < extra) net_->SendTo(rank_ + pow2, data + start[rank_ + pow2], len[rank_ + pow2]);
}

void AllreduceEngine::Allreduce(char* data, int count, int type_size, const ReduceFunction& reducer) {
  if (size_ == 1 || count == 0) return;
  const int bytes = count * type_size;
  if (count < size_ || bytes < 4096) {
    // small: all-gather everything, reduce locally in rank order (deterministic)
    std::vector<char> all(static_cast<size_t>(bytes) * size_);
    Allgather(data, bytes, all.data());
    std::memcpy(data, all.data(), bytes);
    for (int r = 1; r < size_; ++r) reducer(all.data() + static_cast<size_t>(r) * bytes, data, bytes);
    return;
  }
  std::vector<int> start(size_), len(size_);
  int per = count / size_;
  for (int r = 0; r < size_; ++r) {
    start[r] = r * per * type_size;
    len[r] = (r == size_ - 1 ? count - per * (size_ - 1) : per) * type_size;
  }
  ReduceScatter(data, start, len, reducer);
  AllgatherV(data, start, len);
}

}  // namespace multiverso
