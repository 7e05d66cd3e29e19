// multiverso-b200 :: device runtime :: symmetric allocations through the CUDA driver's virtual memory
// management API, bound to an NVLS multicast object (internal header).
//
// The cudaIpc slabs of SymmBuffer cannot be bound to a multicast object: NVLS (multimem.ld_reduce / multimem.st,
// the in-switch reduction the K6 / K1 kernels use when they are given a multicast address) needs physical
// allocations created with cuMemCreate, shared between the processes as POSIX file descriptors and mapped by
// every rank next to ONE multicast object that all devices have been added to.  The Python backend gets this from
// torch.distributed._symmetric_memory (runtime.py: MulticastBuffer); this file is the same recipe for the
// all-native runtime, with no dependency besides libcuda.so.1 (resolved with dlopen at the first use, so the
// library still loads on a machine without a driver).
//
// Reference counterpart: none (the reference moves every byte through MPI_Isend / ZMQ, mpi_net.h:147-151 for the
// all-reduce); SURVEY 5.8 maps MV_Aggregate and the dense Add onto NVLS.
#ifndef MULTIVERSO_DEVICE_RT_VMM_H_
#define MULTIVERSO_DEVICE_RT_VMM_H_

#include <cstddef>
#include <functional>
#include <string>

namespace multiverso {
namespace device {
namespace vmm {

constexpr int kMaxPeers = 8;

// all-gather of `bytes` bytes per rank over the control plane (rank-major result) -- also the only rendezvous
// this module uses
using AllGatherFn = std::function<void(const void* mine, size_t bytes, void* all)>;

struct Mapping {
  void* ptrs[kMaxPeers] = {nullptr};   // this rank's view of every rank's slab (ptrs[rank] = own)
  void* multicast = nullptr;           // the multicast view of all slabs
  size_t size = 0;                     // bytes mapped per rank (>= the request, multiple of the granularity)
  // driver handles for the teardown
  unsigned long long handles[kMaxPeers] = {0};
  unsigned long long mc_handle = 0;
  int rank = 0, world = 0, dev = 0;
};

// Can this process use the driver API at all, does the device support multicast objects and POSIX-fd handles?
bool Available(int dev, std::string* why);

// Collective.  On success every rank holds the same layout; when ANY rank fails every rank returns false and
// nothing stays allocated (the caller falls back to the cudaIpc path).
bool Allocate(size_t bytes, int rank, int world, int dev, const AllGatherFn& allgather, Mapping* out, std::string* why);

// Collective: unmap the peers and the multicast view, rendezvous, release the own slab.
void Release(Mapping* m, const AllGatherFn& allgather);

// Duplicate file descriptor `fd` of process `pid` into this process (pidfd_open + pidfd_getfd).  Returns the new
// descriptor or -1.  The owner must have called AllowFdDuplication() (Yama ptrace scope 1 otherwise refuses a
// sibling process).
int DupFdFromPid(int pid, int fd);
void AllowFdDuplication();

}  // namespace vmm
}  // namespace device
}  // namespace multiverso

// C entry points for the tests (tests/test_vmm_fd_exchange.py drives them through ctypes on the CPU box).
extern "C" {
int mvd_dup_fd_from_pid(int pid, int fd);
void mvd_allow_fd_duplication(void);
int mvd_vmm_available(int dev, char* why, int why_len);
int mvd_vmm_selftest_hostmapped(long long bytes, char* msg, int msg_len);   /* test double only, see vmm.cpp */
}

#endif  // MULTIVERSO_DEVICE_RT_VMM_H_
