// multiverso-b200 :: device-side common definitions (sm_100a only).
//
// Everything the data-plane kernels share: the peer-pointer table, the
// system-scope memory-model helpers used on signal pads, 128-bit vector
// access helpers and the server-side updater functors (the reference's
// Updater<T> family, include/multiverso/updater/*.h, re-expressed as
// register-level epilogues so they fuse into the Add kernels).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "mvb200.h"

#define MVB_DEVINL __device__ __forceinline__

// ---------------------------------------------------------------------------
// Peer table. Passed BY VALUE to kernels (lands in the constant bank, no extra
// dependent load on the critical path). p[r] is rank r's mapping of the same
// symmetric allocation (cudaIpc-mapped for r != me, local pointer for me).
// ---------------------------------------------------------------------------
struct MvbPeers {
  void* p[MVB_MAX_RANKS];
};

// ---------------------------------------------------------------------------
// system-scope memory model helpers (signal pads live in peer-visible HBM).
// Writers: data stores ... fence.acq_rel.sys ... st.release.sys(flag)
// Readers: ld.acquire.sys(flag) ... data loads
// ---------------------------------------------------------------------------
MVB_DEVINL void st_release_sys_u64(uint64_t* p, uint64_t v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
MVB_DEVINL uint64_t ld_acquire_sys_u64(const uint64_t* p) {
  uint64_t v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
MVB_DEVINL uint64_t ld_relaxed_sys_u64(const uint64_t* p) {
  uint64_t v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
MVB_DEVINL void fence_sys() { asm volatile("fence.acq_rel.sys;" ::: "memory"); }

// Spin until *flag >= target, with a wall-clock watchdog (a dead
// rank must produce a diagnostic, not a silent hang). Returns false on timeout.
MVB_DEVINL bool spin_wait_ge(const uint64_t* flag, uint64_t target, long long budget_cycles) {
  long long t0 = clock64();
  int backoff = 8;
  while (ld_acquire_sys_u64(flag) < target) {
    __nanosleep(backoff);
    if (backoff < 256) backoff <<= 1;
    if (clock64() - t0 > budget_cycles) return false;
  }
  return true;
}

// ---------------------------------------------------------------------------
// 16-byte vector access. Peer reads use the non-coherent path with no L1
// allocation: peer lines bypass the local L2 anyway (B300_MICROARCH: "L1-cache,
// L2-BYPASS") and L1 is invalidated at every launch, so .nc is safe for data
// produced by earlier kernels and fenced by a signal.
// ---------------------------------------------------------------------------
MVB_DEVINL uint4 ld_nc_v4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
MVB_DEVINL uint4 ld_v4(const void* p) {
  uint4 r;
  asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
MVB_DEVINL void st_v4(void* p, uint4 v) {
  asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w)
               : "memory");
}
MVB_DEVINL void st_na_v4(void* p, uint4 v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x),
               "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}
// One-sided vector reduction into (possibly peer) memory: the async-PS push.
MVB_DEVINL void red_add_v4_f32(float* p, float4 v) {
  asm volatile("red.relaxed.sys.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x),
               "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
MVB_DEVINL void red_add_f32(float* p, float v) {
  asm volatile("red.relaxed.sys.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
MVB_DEVINL void red_add_f64(double* p, double v) {
  asm volatile("red.relaxed.sys.global.add.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
MVB_DEVINL void red_add_s32(int* p, int v) {
  asm volatile("red.relaxed.sys.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// NVLS (NVSwitch in-network reduction): one load on the MULTICAST address returns the sum of
// the same location in every rank's copy; one store replicates to every rank's copy.
MVB_DEVINL float4 multimem_ld_reduce_add_v4_f32(const void* mc_ptr) {
  float4 r;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(mc_ptr)
               : "memory");
  return r;
}
MVB_DEVINL void multimem_st_v4_f32(void* mc_ptr, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc_ptr), "f"(v.x),
               "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

template <typename T>
struct VecOf;  // 16-byte vector of T
template <>
struct VecOf<float> {
  static constexpr int N = 4;
  float v[4];
};
template <>
struct VecOf<int> {
  static constexpr int N = 4;
  int v[4];
};
template <>
struct VecOf<double> {
  static constexpr int N = 2;
  double v[2];
};

template <typename T>
MVB_DEVINL VecOf<T> vec_load_nc(const T* p) {
  uint4 r = ld_nc_v4(p);
  return *reinterpret_cast<VecOf<T>*>(&r);
}
template <typename T>
MVB_DEVINL VecOf<T> vec_load(const T* p) {
  uint4 r = ld_v4(p);
  return *reinterpret_cast<VecOf<T>*>(&r);
}
template <typename T>
MVB_DEVINL void vec_store(T* p, const VecOf<T>& v) {
  st_v4(p, *reinterpret_cast<const uint4*>(&v));
}

MVB_DEVINL void red_add(float* p, float v) { red_add_f32(p, v); }
MVB_DEVINL void red_add(double* p, double v) { red_add_f64(p, v); }
MVB_DEVINL void red_add(int* p, int v) { red_add_s32(p, v); }

// ---------------------------------------------------------------------------
// Updater functors. `Apply` consumes ONE worker's delta for
// ONE element, entirely in registers. State slabs:
//   momentum : s0 = smoothed gradient          (shared across workers)
//   adagrad  : s0 = G^2 history                (per worker)
//   dcasgd   : s0 = shadow copy                (per worker)
//   dcasgda  : s0 = shadow copy, s1 = mean sq  (per worker)
// Semantics follow the reference accessors (Q1/Q3); AdaGrad implements the
// *intended* rule (Q2: the reference copies the history by value and subtracts
// squares, so its history never persists).
// ---------------------------------------------------------------------------
template <int UPD, typename T>
struct Updater;

template <typename T>
struct Updater<MVB_UPD_DEFAULT, T> {
  static constexpr int kStates = 0;
  static constexpr bool kPerWorker = false;
  MVB_DEVINL static void Apply(T& d, T g, T&, T&, const MvbAddOpt&) { d += g; }
};
template <typename T>
struct Updater<MVB_UPD_SGD, T> {
  static constexpr int kStates = 0;
  static constexpr bool kPerWorker = false;
  MVB_DEVINL static void Apply(T& d, T g, T&, T&, const MvbAddOpt&) { d -= g; }
};
template <typename T>
struct Updater<MVB_UPD_MOMENTUM, T> {
  static constexpr int kStates = 1;
  static constexpr bool kPerWorker = false;
  MVB_DEVINL static void Apply(T& d, T g, T& s0, T&, const MvbAddOpt& o) {
    s0 = (T)o.momentum * s0 + (T)(1.0f - o.momentum) * g;
    d -= s0;
  }
};
template <typename T>
struct Updater<MVB_UPD_ADAGRAD, T> {
  static constexpr int kStates = 1;
  static constexpr bool kPerWorker = true;
  MVB_DEVINL static void Apply(T& d, T g, T& s0, T&, const MvbAddOpt& o) {
    T gn = g / (T)o.lr;
    s0 += gn * gn;
    d -= (T)o.rho * rsqrt_t(s0 + (T)1e-6) * gn;
  }
  MVB_DEVINL static float rsqrt_t(float x) { return rsqrtf(x); }
  MVB_DEVINL static double rsqrt_t(double x) { return rsqrt(x); }
};
template <typename T>
struct Updater<MVB_UPD_DCASGD, T> {
  static constexpr int kStates = 1;
  static constexpr bool kPerWorker = true;
  MVB_DEVINL static void Apply(T& d, T g, T& s0, T&, const MvbAddOpt& o) {
    T gn = g / (T)o.lr;
    d -= (T)o.lr * (gn + (T)o.lambda * gn * gn * (d - s0));
    s0 = d;
  }
};
template <typename T>
struct Updater<MVB_UPD_DCASGDA, T> {
  static constexpr int kStates = 2;
  static constexpr bool kPerWorker = true;
  MVB_DEVINL static void Apply(T& d, T g, T& s0, T& s1, const MvbAddOpt& o) {
    T gn = g / (T)o.lr;
    s1 = (T)o.momentum * s1 + (T)(1.0f - o.momentum) * gn * gn;
    d -= (T)o.lr * (gn + (T)o.lambda * Updater<MVB_UPD_ADAGRAD, T>::rsqrt_t(s1 + (T)1e-7) * gn *
                             gn * (d - s0));
    s0 = d;
  }
};
// int tables always use the plain add (reference: Updater<int>::GetUpdater).

MVB_DEVINL float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

#define MVB_CUDA_CHECK(expr)                                  \
  do {                                                        \
    cudaError_t _e = (expr);                                  \
    if (_e != cudaSuccess) {                                  \
      mvb_set_error(#expr, _e, __FILE__, __LINE__);           \
      return (int)_e;                                         \
    }                                                         \
  } while (0)

extern "C" void mvb_set_error(const char* what, cudaError_t e, const char* file, int line);

static inline int mvb_num_sms() {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}
