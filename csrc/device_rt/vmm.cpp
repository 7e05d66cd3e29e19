// multiverso-b200 :: device runtime :: VMM symmetric allocations bound to an NVLS multicast object (see vmm.h)
#include "vmm.h"

#include <cuda.h>   // types and enums only: every entry point is resolved from libcuda.so.1 at run time
#include <dlfcn.h>
#include <sys/prctl.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <cerrno>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <algorithm>
#include <vector>

#include "multiverso/multiverso.h"

#ifndef SYS_pidfd_open
#define SYS_pidfd_open 434
#endif
#ifndef SYS_pidfd_getfd
#define SYS_pidfd_getfd 438
#endif
#ifndef PR_SET_PTRACER
#define PR_SET_PTRACER 0x59616d61
#endif
#ifndef PR_SET_PTRACER_ANY
#define PR_SET_PTRACER_ANY ((unsigned long)-1)
#endif

namespace multiverso {
namespace device {
namespace vmm {
namespace {

// The subset of the driver API this module needs, resolved once.
struct Driver {
  void* lib = nullptr;
  bool ok = false;
  std::string error;
#define MVD_DRV(name) decltype(&::name) name = nullptr
  MVD_DRV(cuInit);
  MVD_DRV(cuGetErrorString);
  MVD_DRV(cuDeviceGet);
  MVD_DRV(cuDeviceGetAttribute);
  MVD_DRV(cuCtxGetCurrent);
  MVD_DRV(cuMemCreate);
  MVD_DRV(cuMemRelease);
  MVD_DRV(cuMemMap);
  MVD_DRV(cuMemUnmap);
  MVD_DRV(cuMemSetAccess);
  MVD_DRV(cuMemAddressReserve);
  MVD_DRV(cuMemAddressFree);
  MVD_DRV(cuMemExportToShareableHandle);
  MVD_DRV(cuMemImportFromShareableHandle);
  MVD_DRV(cuMulticastCreate);
  MVD_DRV(cuMulticastAddDevice);
  MVD_DRV(cuMulticastBindMem);
  MVD_DRV(cuMulticastUnbind);
  MVD_DRV(cuMulticastGetGranularity);
#undef MVD_DRV
};

Driver& Drv() {
  static Driver d;
  static std::once_flag once;
  std::call_once(once, [] {
    d.lib = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
    if (!d.lib) {
      d.error = "libcuda.so.1 not found";
      return;
    }
    bool all = true;
#define MVD_SYM(name)                                                        \
  do {                                                                       \
    d.name = reinterpret_cast<decltype(d.name)>(dlsym(d.lib, #name));        \
    if (!d.name) {                                                           \
      all = false;                                                           \
      d.error = std::string("libcuda.so.1 lacks ") + #name;                  \
    }                                                                        \
  } while (0)
    MVD_SYM(cuInit);
    MVD_SYM(cuGetErrorString);
    MVD_SYM(cuDeviceGet);
    MVD_SYM(cuDeviceGetAttribute);
    MVD_SYM(cuCtxGetCurrent);
    MVD_SYM(cuMemCreate);
    MVD_SYM(cuMemRelease);
    MVD_SYM(cuMemMap);
    MVD_SYM(cuMemUnmap);
    MVD_SYM(cuMemSetAccess);
    MVD_SYM(cuMemAddressReserve);
    MVD_SYM(cuMemAddressFree);
    MVD_SYM(cuMemExportToShareableHandle);
    MVD_SYM(cuMemImportFromShareableHandle);
    MVD_SYM(cuMulticastCreate);
    MVD_SYM(cuMulticastAddDevice);
    MVD_SYM(cuMulticastBindMem);
    MVD_SYM(cuMulticastUnbind);
    MVD_SYM(cuMulticastGetGranularity);
#undef MVD_SYM
    if (all && d.cuInit(0) != CUDA_SUCCESS) {
      all = false;
      d.error = "cuInit failed (no usable GPU driver)";
    }
    d.ok = all;
  });
  return d;
}

std::string ErrName(CUresult r) {
  const char* s = nullptr;
  if (Drv().cuGetErrorString && Drv().cuGetErrorString(r, &s) == CUDA_SUCCESS && s) return s;
  return "CUDA driver error " + std::to_string(static_cast<int>(r));
}

// One step of the collective: `what` names it in the error text.
#define MVD_TRY(call, what)                                 \
  do {                                                      \
    const CUresult r__ = (call);                            \
    if (r__ != CUDA_SUCCESS && ok) {                        \
      ok = false;                                           \
      err = std::string(what) + ": " + ErrName(r__);        \
    }                                                       \
  } while (0)

// Every rank reports its status; the step succeeded only if it succeeded everywhere.
bool AllOk(bool mine, int world, const AllGatherFn& allgather) {
  char flag = mine ? 1 : 0;
  std::vector<char> all(static_cast<size_t>(world), 0);
  allgather(&flag, 1, all.data());
  for (char f : all)
    if (!f) return false;
  return true;
}

struct Wire {          // what the ranks tell each other about their handles
  int pid;
  int mem_fd;
  int mc_fd;           // rank 0 only
  int ok;
};

size_t RoundUp(size_t v, size_t g) { return (v + g - 1) / g * g; }

// Map `handle` (a slab or the multicast object) read-write for device `dev`.
bool MapHandle(Driver& d, CUmemGenericAllocationHandle handle, size_t size, size_t align, int dev, void** out, std::string* err) {
  CUdeviceptr va = 0;
  CUresult r = d.cuMemAddressReserve(&va, size, align, 0, 0);
  if (r != CUDA_SUCCESS) {
    *err = "cuMemAddressReserve: " + ErrName(r);
    return false;
  }
  r = d.cuMemMap(va, size, 0, handle, 0);
  if (r != CUDA_SUCCESS) {
    *err = "cuMemMap: " + ErrName(r);
    d.cuMemAddressFree(va, size);
    return false;
  }
  CUmemAccessDesc acc;
  std::memset(&acc, 0, sizeof acc);
  acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  acc.location.id = dev;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  r = d.cuMemSetAccess(va, size, &acc, 1);
  if (r != CUDA_SUCCESS) {
    *err = "cuMemSetAccess: " + ErrName(r);
    d.cuMemUnmap(va, size);
    d.cuMemAddressFree(va, size);
    return false;
  }
  *out = reinterpret_cast<void*>(va);
  return true;
}

void UnmapPtr(Driver& d, void* p, size_t size) {
  if (!p) return;
  const CUdeviceptr va = reinterpret_cast<CUdeviceptr>(p);
  d.cuMemUnmap(va, size);
  d.cuMemAddressFree(va, size);
}

}  // namespace

void AllowFdDuplication() {
  // Yama ptrace scope 1 lets only ancestors attach; the ranks of a job are siblings
  prctl(PR_SET_PTRACER, PR_SET_PTRACER_ANY, 0, 0, 0);
}

int DupFdFromPid(int pid, int fd) {
  if (pid == static_cast<int>(getpid())) return dup(fd);
  const int pidfd = static_cast<int>(syscall(SYS_pidfd_open, pid, 0));
  if (pidfd < 0) return -1;
  const int got = static_cast<int>(syscall(SYS_pidfd_getfd, pidfd, fd, 0));
  const int saved = errno;
  close(pidfd);
  errno = saved;
  return got;
}

bool Available(int dev, std::string* why) {
  Driver& d = Drv();
  if (!d.ok) {
    if (why) *why = d.error;
    return false;
  }
  CUdevice cudev;
  if (d.cuDeviceGet(&cudev, dev) != CUDA_SUCCESS) {
    if (why) *why = "cuDeviceGet failed";
    return false;
  }
  int vm = 0, fdh = 0, mc = 0;
  d.cuDeviceGetAttribute(&vm, CU_DEVICE_ATTRIBUTE_VIRTUAL_MEMORY_MANAGEMENT_SUPPORTED, cudev);
  d.cuDeviceGetAttribute(&fdh, CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED, cudev);
  d.cuDeviceGetAttribute(&mc, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, cudev);
  if (!vm || !fdh || !mc) {
    if (why)
      *why = std::string("device lacks ") + (!vm ? "virtual memory management" : !fdh ? "POSIX fd handles" : "multicast objects");
    return false;
  }
  return true;
}

bool Allocate(size_t bytes, int rank, int world, int dev, const AllGatherFn& allgather, Mapping* out, std::string* why) {
  std::string err;
  bool ok = world >= 2 && world <= kMaxPeers && Available(dev, &err);
  if (world < 2 || world > kMaxPeers) err = "multicast allocations need 2.." + std::to_string(kMaxPeers) + " ranks";
  Driver& d = Drv();
  Mapping m;
  m.rank = rank;
  m.world = world;
  m.dev = dev;
  auto fail = [&](const std::string& e) {
    if (why) *why = e;
    return false;
  };
  // step 0: everybody can use the driver, a context is current
  if (ok) {
    CUcontext ctx = nullptr;
    if (d.cuCtxGetCurrent(&ctx) != CUDA_SUCCESS || ctx == nullptr) {
      ok = false;
      err = "no current CUDA context (the runtime must have selected the device first)";
    }
  }
  if (!AllOk(ok, world, allgather)) return fail(ok ? "a peer cannot use NVLS multicast" : err);

  // step 1: own slab + (rank 0) the multicast object, exported as file descriptors
  CUdevice cudev = 0;
  d.cuDeviceGet(&cudev, dev);
  CUmulticastObjectProp mcprop;
  std::memset(&mcprop, 0, sizeof mcprop);
  mcprop.numDevices = static_cast<unsigned int>(world);
  mcprop.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  mcprop.size = std::max<size_t>(bytes, 1);
  size_t gran = 0;
  MVD_TRY(d.cuMulticastGetGranularity(&gran, &mcprop, CU_MULTICAST_GRANULARITY_RECOMMENDED), "cuMulticastGetGranularity");
  if (ok && gran == 0) {
    ok = false;
    err = "multicast granularity 0";
  }
  m.size = ok ? RoundUp(std::max<size_t>(bytes, 1), gran) : 0;
  mcprop.size = m.size;
  CUmemGenericAllocationHandle mine = 0, mc = 0;
  int mem_fd = -1, mc_fd = -1;
  if (ok) {
    CUmemAllocationProp prop;
    std::memset(&prop, 0, sizeof prop);
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = dev;
    prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    MVD_TRY(d.cuMemCreate(&mine, m.size, &prop, 0), "cuMemCreate");
    if (ok) MVD_TRY(d.cuMemExportToShareableHandle(&mem_fd, mine, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0), "cuMemExportToShareableHandle");
    if (ok && rank == 0) {
      MVD_TRY(d.cuMulticastCreate(&mc, &mcprop), "cuMulticastCreate");
      if (ok) MVD_TRY(d.cuMemExportToShareableHandle(&mc_fd, mc, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0), "export of the multicast object");
    }
  }
  AllowFdDuplication();
  Wire me{static_cast<int>(getpid()), mem_fd, mc_fd, ok ? 1 : 0};
  std::vector<Wire> wire(static_cast<size_t>(world));
  allgather(&me, sizeof me, wire.data());
  bool everybody = true;
  for (const Wire& w : wire) everybody = everybody && w.ok;

  // step 2: import the peers' slabs and the multicast object, add this device to it
  std::vector<CUmemGenericAllocationHandle> handles(static_cast<size_t>(world), 0);
  handles[rank] = mine;
  if (ok && everybody) {
    for (int r = 0; r < world && ok; ++r) {
      if (r == rank) continue;
      const int fd = DupFdFromPid(wire[r].pid, wire[r].mem_fd);
      if (fd < 0) {
        ok = false;
        err = std::string("pidfd_getfd of a peer's slab: ") + std::strerror(errno);
        break;
      }
      MVD_TRY(d.cuMemImportFromShareableHandle(&handles[r], reinterpret_cast<void*>(static_cast<intptr_t>(fd)),
                                               CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR), "cuMemImportFromShareableHandle");
      close(fd);
    }
    if (ok && rank != 0) {
      const int fd = DupFdFromPid(wire[0].pid, wire[0].mc_fd);
      if (fd < 0) {
        ok = false;
        err = std::string("pidfd_getfd of the multicast object: ") + std::strerror(errno);
      } else {
        MVD_TRY(d.cuMemImportFromShareableHandle(&mc, reinterpret_cast<void*>(static_cast<intptr_t>(fd)),
                                                 CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR), "import of the multicast object");
        close(fd);
      }
    }
    if (ok) MVD_TRY(d.cuMulticastAddDevice(mc, cudev), "cuMulticastAddDevice");
  }
  // every import is done (or has failed) everywhere: the exported descriptors can go, and -- on success -- all
  // devices have been added, which cuMulticastBindMem requires
  const bool step2 = AllOk(ok && everybody, world, allgather);
  if (mem_fd >= 0) close(mem_fd);
  if (mc_fd >= 0) close(mc_fd);

  // step 3: bind the own slab, map everything
  if (step2) {
    MVD_TRY(d.cuMulticastBindMem(mc, 0, mine, 0, m.size, 0), "cuMulticastBindMem");
    for (int r = 0; r < world && ok; ++r) {
      if (!MapHandle(d, handles[r], m.size, gran, dev, &m.ptrs[r], &err)) ok = false;
    }
    if (ok && !MapHandle(d, mc, m.size, gran, dev, &m.multicast, &err)) ok = false;
  }
  const bool step3 = step2 && AllOk(ok, world, allgather);
  if (!step3) {
    // undo whatever this rank got; nobody keeps a mapping of anybody's slab
    for (int r = 0; r < world; ++r) UnmapPtr(d, m.ptrs[r], m.size);
    UnmapPtr(d, m.multicast, m.size);
    AllOk(true, world, allgather);                       // every mapping of every slab is gone
    if (mc) {
      if (step2) d.cuMulticastUnbind(mc, cudev, 0, m.size);
      d.cuMemRelease(mc);
    }
    for (int r = 0; r < world; ++r)
      if (handles[r]) d.cuMemRelease(handles[r]);
    return fail(err.empty() ? "a peer failed to set up the multicast mapping" : err);
  }
  for (int r = 0; r < world; ++r) m.handles[r] = handles[r];
  m.mc_handle = mc;
  *out = m;
  return true;
}

void Release(Mapping* m, const AllGatherFn& allgather) {
  if (!m || m->size == 0) return;
  Driver& d = Drv();
  CUdevice cudev = 0;
  d.cuDeviceGet(&cudev, m->dev);
  UnmapPtr(d, m->multicast, m->size);
  for (int r = 0; r < m->world; ++r) {
    UnmapPtr(d, m->ptrs[r], m->size);
    m->ptrs[r] = nullptr;
  }
  m->multicast = nullptr;
  AllOk(true, m->world, allgather);                      // nobody maps anybody's slab any more
  if (m->mc_handle) {
    d.cuMulticastUnbind(m->mc_handle, cudev, 0, m->size);
    d.cuMemRelease(m->mc_handle);
  }
  for (int r = 0; r < m->world; ++r)
    if (m->handles[r]) d.cuMemRelease(m->handles[r]);
  *m = Mapping();
}

}  // namespace vmm
}  // namespace device
}  // namespace multiverso

extern "C" int mvd_dup_fd_from_pid(int pid, int fd) { return multiverso::device::vmm::DupFdFromPid(pid, fd); }
extern "C" void mvd_allow_fd_duplication(void) { multiverso::device::vmm::AllowFdDuplication(); }
extern "C" int mvd_vmm_available(int dev, char* why, int why_len) {
  std::string w;
  const bool ok = multiverso::device::vmm::Available(dev, &w);
  if (why && why_len > 0) std::snprintf(why, static_cast<size_t>(why_len), "%s", w.c_str());
  return ok ? 1 : 0;
}

// Protocol self-test for the CPU box: allocate over the host control plane (MV_Init must have run), stamp the own
// slab THROUGH A CPU POINTER, read every peer's stamp through the local mapping of its slab, write through the
// multicast view, release.  Only meaningful against the driver test double (tests/fake_libcuda.c), where a "slab"
// is host memory; with a real driver the pointers are device addresses and this function must not be called.
// Returns 1 = allocated and verified, 0 = collectively fell back (msg = why), -1 = verification failed.
extern "C" int mvd_vmm_selftest_hostmapped(long long bytes, char* msg, int msg_len) {
  namespace vmm = multiverso::device::vmm;
  const int rank = multiverso::MV_Rank(), world = multiverso::MV_Size();
  const vmm::AllGatherFn gather = [rank, world](const void* mine, size_t n, void* all) {
    std::memset(all, 0, n * static_cast<size_t>(world));
    std::memcpy(static_cast<char*>(all) + n * static_cast<size_t>(rank), mine, n);
    if (world > 1) multiverso::MV_Aggregate(static_cast<char*>(all), static_cast<int>(n * static_cast<size_t>(world)));
  };
  auto say = [&](const std::string& m) {
    if (msg && msg_len > 0) std::snprintf(msg, static_cast<size_t>(msg_len), "%s", m.c_str());
  };
  vmm::Mapping m;
  std::string why;
  if (!vmm::Allocate(static_cast<size_t>(bytes), rank, world, 0, gather, &m, &why)) {
    say(why);
    return 0;
  }
  auto rendezvous = [&] {
    char one = 1;
    std::vector<char> all(static_cast<size_t>(world));
    gather(&one, 1, all.data());
  };
  bool ok = m.size >= static_cast<size_t>(bytes) && m.multicast != nullptr;
  const size_t words = m.size / 4;
  auto* mine = static_cast<volatile uint32_t*>(m.ptrs[rank]);
  mine[0] = 0xC0DE0000u + static_cast<uint32_t>(rank);
  mine[words - 1] = 0xFEED0000u + static_cast<uint32_t>(rank);
  if (rank == 0) static_cast<volatile uint32_t*>(m.multicast)[1] = 0xABCD1234u;
  rendezvous();
  for (int r = 0; r < world; ++r) {
    auto* p = static_cast<volatile uint32_t*>(m.ptrs[r]);
    ok = ok && p != nullptr && p[0] == 0xC0DE0000u + static_cast<uint32_t>(r) && p[words - 1] == 0xFEED0000u + static_cast<uint32_t>(r);
  }
  ok = ok && static_cast<volatile uint32_t*>(m.multicast)[1] == 0xABCD1234u;
  rendezvous();
  vmm::Release(&m, gather);
  say(ok ? "ok" : "a peer's slab did not show its stamp through the local mapping");
  return ok ? 1 : -1;
}
