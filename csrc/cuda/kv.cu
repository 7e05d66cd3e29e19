// multiverso-b200 :: K5, the KVTable as a GPU hash table.
//
// Reference: KVWorkerTable/KVServerTable (include/multiverso/table/kv_table.h:18-118):
// keys are hash-partitioned over servers (key % num_servers), the server keeps an
// unordered_map and ProcessAdd does table_[k] += v, ProcessGet default-constructs
// missing keys.  Here every server shard is an open-addressing table in HBM
// (int64 keys, linear probing); remote shards are reached with system-scope CAS /
// atomic adds through the peer mapping, so Add/Get are single one-sided kernels.
#include "mvb_common.cuh"

namespace {

constexpr long long kEmpty = (long long)0x8000000000000000ull;

struct KVDev {
  int S;
  int64_t cap;
  long long* keys[MVB_MAX_RANKS];
  void* vals[MVB_MAX_RANKS];
};

MVB_DEVINL uint64_t mix64(uint64_t x) {
  x ^= x >> 33;
  x *= 0xff51afd7ed558ccdull;
  x ^= x >> 33;
  x *= 0xc4ceb9fe1a85ec53ull;
  x ^= x >> 33;
  return x;
}
MVB_DEVINL int owner_of(long long key, int S) {
  long long m = key % S;
  return (int)(m < 0 ? m + S : m);
}

template <typename V>
MVB_DEVINL void atomic_add_sys(V* p, V v);
template <>
MVB_DEVINL void atomic_add_sys<float>(float* p, float v) { atomicAdd_system(p, v); }
template <>
MVB_DEVINL void atomic_add_sys<double>(double* p, double v) { atomicAdd_system(p, v); }
template <>
MVB_DEVINL void atomic_add_sys<int>(int* p, int v) { atomicAdd_system(p, v); }
template <>
MVB_DEVINL void atomic_add_sys<long long>(long long* p, long long v) {
  atomicAdd_system(reinterpret_cast<unsigned long long*>(p), (unsigned long long)v);
}

__global__ void kv_init_kernel(long long* keys, int64_t cap) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += stride)
    keys[i] = kEmpty;
}

template <typename V>
__global__ void kv_add_kernel(const __grid_constant__ KVDev kv, const long long* __restrict__ keys,
                              const V* __restrict__ vals, int64_t n, int* err) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const long long key = keys[i];
    const int o = owner_of(key, kv.S);
    long long* tk = kv.keys[o];
    V* tv = reinterpret_cast<V*>(kv.vals[o]);
    uint64_t slot = mix64((uint64_t)key) & (uint64_t)(kv.cap - 1);
    bool done = false;
    for (int64_t probe = 0; probe < kv.cap; ++probe) {
      long long cur = *reinterpret_cast<volatile long long*>(tk + slot);
      if (cur == kEmpty) {
        cur = (long long)atomicCAS_system(reinterpret_cast<unsigned long long*>(tk + slot),
                                          (unsigned long long)kEmpty, (unsigned long long)key);
        if (cur == kEmpty) cur = key;
      }
      if (cur == key) {
        atomic_add_sys<V>(tv + slot, vals[i]);
        done = true;
        break;
      }
      slot = (slot + 1) & (uint64_t)(kv.cap - 1);
    }
    if (!done && err) atomicExch(err, 5001);  // table full
  }
}

template <typename V>
__global__ void kv_get_kernel(const __grid_constant__ KVDev kv, const long long* __restrict__ keys,
                              V* __restrict__ out, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const long long key = keys[i];
    const int o = owner_of(key, kv.S);
    const long long* tk = kv.keys[o];
    const V* tv = reinterpret_cast<const V*>(kv.vals[o]);
    uint64_t slot = mix64((uint64_t)key) & (uint64_t)(kv.cap - 1);
    V result = (V)0;
    for (int64_t probe = 0; probe < kv.cap; ++probe) {
      long long cur = *reinterpret_cast<const volatile long long*>(tk + slot);
      if (cur == key) {
        result = *reinterpret_cast<const volatile V*>(tv + slot);
        break;
      }
      if (cur == kEmpty) break;
      slot = (slot + 1) & (uint64_t)(kv.cap - 1);
    }
    out[i] = result;
  }
}

template <typename V>
__global__ void kv_dump_kernel(const long long* keys, const V* vals, int64_t cap,
                               long long* out_keys, V* out_vals, unsigned long long* count) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += stride) {
    long long k = keys[i];
    if (k != kEmpty) {
      unsigned long long pos = atomicAdd(count, 1ull);
      out_keys[pos] = k;
      out_vals[pos] = vals[i];
    }
  }
}

// growth: number of live keys of a shard, and re-insertion of a shard into a larger local table
__global__ void kv_count_kernel(const long long* keys, int64_t cap, unsigned long long* count) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  unsigned long long c = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cap; i += stride) c += keys[i] != kEmpty;
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(count, c);
}
template <typename V>
__global__ void kv_rehash_kernel(const long long* okeys, const V* ovals, int64_t ocap, long long* nkeys, V* nvals,
                                 int64_t ncap, int* err) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < ocap; i += stride) {
    const long long key = okeys[i];
    if (key == kEmpty) continue;
    uint64_t slot = mix64((uint64_t)key) & (uint64_t)(ncap - 1);
    bool done = false;
    for (int64_t probe = 0; probe < ncap; ++probe) {
      const long long cur = (long long)atomicCAS(reinterpret_cast<unsigned long long*>(nkeys + slot),
                                                 (unsigned long long)kEmpty, (unsigned long long)key);
      if (cur == kEmpty) { nvals[slot] = ovals[i]; done = true; break; }   // keys are unique in the old table
      slot = (slot + 1) & (uint64_t)(ncap - 1);
    }
    if (!done && err) atomicExch(err, 5002);
  }
}

KVDev to_dev(const MvbKV* h) {
  KVDev d{};
  d.S = h->nservers;
  d.cap = h->capacity;
  for (int s = 0; s < MVB_MAX_RANKS; ++s) {
    d.keys[s] = s < h->nservers ? (long long*)h->keys[s] : nullptr;
    d.vals[s] = s < h->nservers ? h->vals[s] : nullptr;
  }
  return d;
}
int blocks_for(int64_t n) {
  int64_t b = (n + 255) / 256;
  int64_t cap = (int64_t)mvb_num_sms() * 8;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace

extern "C" int mvb_kv_init(void* keys, int64_t capacity, void* stream) {
  kv_init_kernel<<<blocks_for(capacity), 256, 0, (cudaStream_t)stream>>>((long long*)keys, capacity);
  MVB_CUDA_CHECK(cudaGetLastError());
  return 0;
}
extern "C" int mvb_kv_add(const MvbKV* kv, const int64_t* keys, const void* vals, int64_t n,
                          int* err_flag, void* stream) {
  if (n <= 0) return 0;
  if (kv->capacity & (kv->capacity - 1)) return -4;
  KVDev d = to_dev(kv);
  cudaStream_t st = (cudaStream_t)stream;
  int b = blocks_for(n);
  switch (kv->vtype) {
    case MVB_F32: kv_add_kernel<float><<<b, 256, 0, st>>>(d, (const long long*)keys, (const float*)vals, n, err_flag); break;
    case MVB_F64: kv_add_kernel<double><<<b, 256, 0, st>>>(d, (const long long*)keys, (const double*)vals, n, err_flag); break;
    case MVB_I32: kv_add_kernel<int><<<b, 256, 0, st>>>(d, (const long long*)keys, (const int*)vals, n, err_flag); break;
    case MVB_I64: kv_add_kernel<long long><<<b, 256, 0, st>>>(d, (const long long*)keys, (const long long*)vals, n, err_flag); break;
    default: return -1;
  }
  MVB_CUDA_CHECK(cudaGetLastError());
  return 0;
}
extern "C" int mvb_kv_get(const MvbKV* kv, const int64_t* keys, void* out_vals, int64_t n,
                          void* stream) {
  if (n <= 0) return 0;
  KVDev d = to_dev(kv);
  cudaStream_t st = (cudaStream_t)stream;
  int b = blocks_for(n);
  switch (kv->vtype) {
    case MVB_F32: kv_get_kernel<float><<<b, 256, 0, st>>>(d, (const long long*)keys, (float*)out_vals, n); break;
    case MVB_F64: kv_get_kernel<double><<<b, 256, 0, st>>>(d, (const long long*)keys, (double*)out_vals, n); break;
    case MVB_I32: kv_get_kernel<int><<<b, 256, 0, st>>>(d, (const long long*)keys, (int*)out_vals, n); break;
    case MVB_I64: kv_get_kernel<long long><<<b, 256, 0, st>>>(d, (const long long*)keys, (long long*)out_vals, n); break;
    default: return -1;
  }
  MVB_CUDA_CHECK(cudaGetLastError());
  return 0;
}
extern "C" int mvb_kv_dump(int vtype, const void* keys, const void* vals, int64_t capacity,
                           int64_t* out_keys, void* out_vals, int64_t* out_count, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  MVB_CUDA_CHECK(cudaMemsetAsync(out_count, 0, 8, st));
  int b = blocks_for(capacity);
  switch (vtype) {
    case MVB_F32: kv_dump_kernel<float><<<b, 256, 0, st>>>((const long long*)keys, (const float*)vals, capacity, (long long*)out_keys, (float*)out_vals, (unsigned long long*)out_count); break;
    case MVB_F64: kv_dump_kernel<double><<<b, 256, 0, st>>>((const long long*)keys, (const double*)vals, capacity, (long long*)out_keys, (double*)out_vals, (unsigned long long*)out_count); break;
    case MVB_I32: kv_dump_kernel<int><<<b, 256, 0, st>>>((const long long*)keys, (const int*)vals, capacity, (long long*)out_keys, (int*)out_vals, (unsigned long long*)out_count); break;
    case MVB_I64: kv_dump_kernel<long long><<<b, 256, 0, st>>>((const long long*)keys, (const long long*)vals, capacity, (long long*)out_keys, (long long*)out_vals, (unsigned long long*)out_count); break;
    default: return -1;
  }
  MVB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

// Live keys of one shard (device count, zeroed here).
extern "C" int mvb_kv_count(const void* keys, int64_t capacity, int64_t* out_count, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  MVB_CUDA_CHECK(cudaMemsetAsync(out_count, 0, 8, st));
  kv_count_kernel<<<blocks_for(capacity), 256, 0, st>>>((const long long*)keys, capacity, (unsigned long long*)out_count);
  MVB_CUDA_CHECK(cudaGetLastError());
  return 0;
}
// Re-insert every live (key, value) of a shard into a new, initialised, larger LOCAL table (growth).
extern "C" int mvb_kv_rehash(int vtype, const void* old_keys, const void* old_vals, int64_t old_cap, void* new_keys,
                             void* new_vals, int64_t new_cap, int* err_flag, void* stream) {
  if (new_cap & (new_cap - 1)) return -4;
  cudaStream_t st = (cudaStream_t)stream;
  const int b = blocks_for(old_cap);
  switch (vtype) {
    case MVB_F32: kv_rehash_kernel<float><<<b, 256, 0, st>>>((const long long*)old_keys, (const float*)old_vals, old_cap, (long long*)new_keys, (float*)new_vals, new_cap, err_flag); break;
    case MVB_F64: kv_rehash_kernel<double><<<b, 256, 0, st>>>((const long long*)old_keys, (const double*)old_vals, old_cap, (long long*)new_keys, (double*)new_vals, new_cap, err_flag); break;
    case MVB_I32: kv_rehash_kernel<int><<<b, 256, 0, st>>>((const long long*)old_keys, (const int*)old_vals, old_cap, (long long*)new_keys, (int*)new_vals, new_cap, err_flag); break;
    case MVB_I64: kv_rehash_kernel<long long><<<b, 256, 0, st>>>((const long long*)old_keys, (const long long*)old_vals, old_cap, (long long*)new_keys, (long long*)new_vals, new_cap, err_flag); break;
    default: return -1;
  }
  MVB_CUDA_CHECK(cudaGetLastError());
  return 0;
}
