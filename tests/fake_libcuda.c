/* Test double of the 19 CUDA driver entry points csrc/device_rt/vmm.cpp resolves from libcuda.so.1, so that the
 * collective allocation protocol (slab per rank, descriptors handed between processes, one multicast object, peer
 * and multicast mappings, agreed fallback, teardown) can be EXECUTED on a machine without a GPU:
 *   - a physical allocation / multicast object is a memfd; export = dup, import = dup of the received descriptor
 *   - cuMemAddressReserve = PROT_NONE anonymous mmap, cuMemMap = MAP_FIXED|MAP_SHARED mmap of the memfd over it
 * so a slab mapped by two processes really is the same memory, and a test can write through one rank's pointer
 * and read through another's.  FAKE_CUDA_FAIL=<function name> makes that call fail in this process.
 * (tests/test_vmm_fd_exchange.py builds it as libcuda.so.1 into a scratch directory; same idea as
 * tests/fake_libhdfs.c.) */
#define _GNU_SOURCE
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <unistd.h>

typedef int CUresult;
typedef int CUdevice;
typedef unsigned long long CUdeviceptr;
typedef unsigned long long CUhandle;

enum { OK = 0, ERR_INVALID = 1, ERR_OOM = 2, ERR_NOT_SUPPORTED = 801 };

#define MAX_H 64
static struct { int fd; size_t size; int used; } H[MAX_H];

static int fails(const char* fn) {
  const char* f = getenv("FAKE_CUDA_FAIL");
  return f && strcmp(f, fn) == 0;
}
#define MAYBE_FAIL(name) do { if (fails(name)) return ERR_NOT_SUPPORTED; } while (0)

static CUhandle put(int fd, size_t size) {
  for (int i = 0; i < MAX_H; ++i)
    if (!H[i].used) {
      H[i].used = 1;
      H[i].fd = fd;
      H[i].size = size;
      return (CUhandle)(i + 1);
    }
  return 0;
}
static int slot(CUhandle h) { return (h >= 1 && h <= MAX_H && H[h - 1].used) ? (int)h - 1 : -1; }

CUresult cuInit(unsigned int flags) { (void)flags; MAYBE_FAIL("cuInit"); return OK; }
CUresult cuGetErrorString(CUresult e, const char** s) {
  *s = e == OK ? "no error" : e == ERR_NOT_SUPPORTED ? "fake: operation not supported" : "fake: invalid value";
  return OK;
}
CUresult cuDeviceGet(CUdevice* d, int ordinal) { *d = ordinal; return OK; }
CUresult cuDeviceGetAttribute(int* v, int attr, CUdevice d) {
  (void)d;
  *v = (attr == 102 || attr == 103) ? 1 : (attr == 132 ? !fails("multicast_attr") : 0);
  return OK;
}
CUresult cuCtxGetCurrent(void** ctx) { *ctx = (void*)0x1; return OK; }

CUresult cuMemCreate(CUhandle* h, size_t size, const void* prop, unsigned long long flags) {
  (void)prop; (void)flags;
  MAYBE_FAIL("cuMemCreate");
  int fd = memfd_create("fake-cuda-slab", 0);
  if (fd < 0 || ftruncate(fd, (off_t)size) != 0) return ERR_OOM;
  *h = put(fd, size);
  return *h ? OK : ERR_OOM;
}
CUresult cuMulticastCreate(CUhandle* h, const void* prop) {
  MAYBE_FAIL("cuMulticastCreate");
  /* CUmulticastObjectProp: unsigned numDevices; size_t size; ... */
  size_t size = *(const size_t*)((const char*)prop + 8);
  int fd = memfd_create("fake-cuda-mc", 0);
  if (fd < 0 || ftruncate(fd, (off_t)size) != 0) return ERR_OOM;
  *h = put(fd, size);
  return *h ? OK : ERR_OOM;
}
CUresult cuMemRelease(CUhandle h) {
  int s = slot(h);
  if (s < 0) return ERR_INVALID;
  close(H[s].fd);
  H[s].used = 0;
  return OK;
}
CUresult cuMemExportToShareableHandle(void* out, CUhandle h, int type, unsigned long long flags) {
  (void)type; (void)flags;
  MAYBE_FAIL("cuMemExportToShareableHandle");
  int s = slot(h);
  if (s < 0) return ERR_INVALID;
  *(int*)out = dup(H[s].fd);
  return *(int*)out >= 0 ? OK : ERR_INVALID;
}
CUresult cuMemImportFromShareableHandle(CUhandle* h, void* os, int type) {
  (void)type;
  MAYBE_FAIL("cuMemImportFromShareableHandle");
  int fd = dup((int)(intptr_t)os);
  if (fd < 0) return ERR_INVALID;
  off_t size = lseek(fd, 0, SEEK_END);
  *h = put(fd, (size_t)size);
  return *h ? OK : ERR_OOM;
}
CUresult cuMulticastGetGranularity(size_t* g, const void* prop, int option) {
  (void)prop; (void)option;
  MAYBE_FAIL("cuMulticastGetGranularity");
  *g = 1 << 16;
  return OK;
}
CUresult cuMulticastAddDevice(CUhandle h, CUdevice d) { (void)d; MAYBE_FAIL("cuMulticastAddDevice"); return slot(h) < 0 ? ERR_INVALID : OK; }
CUresult cuMulticastBindMem(CUhandle mc, size_t mo, CUhandle mem, size_t off, size_t size, unsigned long long flags) {
  (void)mo; (void)off; (void)size; (void)flags;
  MAYBE_FAIL("cuMulticastBindMem");
  return (slot(mc) < 0 || slot(mem) < 0) ? ERR_INVALID : OK;
}
CUresult cuMulticastUnbind(CUhandle mc, CUdevice d, size_t off, size_t size) { (void)d; (void)off; (void)size; return slot(mc) < 0 ? ERR_INVALID : OK; }

CUresult cuMemAddressReserve(CUdeviceptr* p, size_t size, size_t align, CUdeviceptr addr, unsigned long long flags) {
  (void)align; (void)addr; (void)flags;
  MAYBE_FAIL("cuMemAddressReserve");
  void* a = mmap(NULL, size, PROT_NONE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
  if (a == MAP_FAILED) return ERR_OOM;
  *p = (CUdeviceptr)(uintptr_t)a;
  return OK;
}
CUresult cuMemAddressFree(CUdeviceptr p, size_t size) { return munmap((void*)(uintptr_t)p, size) == 0 ? OK : ERR_INVALID; }
CUresult cuMemMap(CUdeviceptr p, size_t size, size_t off, CUhandle h, unsigned long long flags) {
  (void)flags;
  MAYBE_FAIL("cuMemMap");
  int s = slot(h);
  if (s < 0 || size > H[s].size) return ERR_INVALID;
  void* a = mmap((void*)(uintptr_t)p, size, PROT_NONE, MAP_SHARED | MAP_FIXED, H[s].fd, (off_t)off);
  return a == MAP_FAILED ? ERR_INVALID : OK;
}
CUresult cuMemUnmap(CUdeviceptr p, size_t size) {
  void* a = mmap((void*)(uintptr_t)p, size, PROT_NONE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_FIXED, -1, 0);
  return a == MAP_FAILED ? ERR_INVALID : OK;
}
CUresult cuMemSetAccess(CUdeviceptr p, size_t size, const void* desc, size_t count) {
  (void)desc; (void)count;
  MAYBE_FAIL("cuMemSetAccess");
  return mprotect((void*)(uintptr_t)p, size, PROT_READ | PROT_WRITE) == 0 ? OK : ERR_INVALID;
}
