// multiverso-b200 :: row-sparse table kernels.
//
//  K4 get_rows        -- MatrixWorkerTable::Get(row ids): Partition by row/num_row_each,
//                        server packs rows, client scatters (src/table/matrix_table.cpp:
//                        266-313, 442-450, 332-338) -> one gather kernel, warp per row,
//                        128-bit peer loads straight from the owner's shard.
//  K3 add_rows_red    -- row Add for stateless updaters: one-sided vector red.add into
//                        the owner shard (commutative => race-free async-PS semantics).
//     add_rows_owner  -- row Add for stateful updaters, applied by the OWNER exactly once
//                        per (worker,row), reading ids/values from (peer) staging.
//  K10 row_nonzero_mask, stale bitmap -- zero-row skip and delta-pull bookkeeping of the
//                        sparse Matrix tables (src/table/matrix.cpp:151-164, 516-572).
#include <type_traits>
#include "mvb_common.cuh"

namespace {

struct RowMapDev {
  int64_t num_row, num_col, rps;
  int S;
  void* shard[MVB_MAX_RANKS];
};

MVB_DEVINL void locate_row(const RowMapDev& m, int64_t r, int& owner, int64_t& local) {
  int64_t o = r / m.rps;
  if (o > m.S - 1) o = m.S - 1;  // last server takes the remainder (matrix_table.cpp:270-273)
  owner = (int)o;
  local = r - o * m.rps;
}

template <typename T, bool VECOK>
__global__ void __launch_bounds__(256)
get_rows_kernel(const __grid_constant__ RowMapDev m, const int64_t* __restrict__ ids, int64_t k,
                T* __restrict__ out, int64_t out_ld) {
  constexpr int VEC = VecOf<T>::N;
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = warp; i < k; i += nwarps) {
    const int64_t r = ids[i];
    if (r < 0 || r >= m.num_row) continue;
    int owner;
    int64_t local;
    locate_row(m, r, owner, local);
    const T* src = reinterpret_cast<const T*>(m.shard[owner]) + local * m.num_col;
    T* dst = out + i * out_ld;
    if constexpr (VECOK) {
      const int nvec = (int)(m.num_col / VEC);
      int v = lane;
      for (; v + 96 < nvec; v += 128) {
        uint4 a = ld_nc_v4(src + (int64_t)v * VEC);
        uint4 b = ld_nc_v4(src + (int64_t)(v + 32) * VEC);
        uint4 c = ld_nc_v4(src + (int64_t)(v + 64) * VEC);
        uint4 d = ld_nc_v4(src + (int64_t)(v + 96) * VEC);
        st_v4(dst + (int64_t)v * VEC, a);
        st_v4(dst + (int64_t)(v + 32) * VEC, b);
        st_v4(dst + (int64_t)(v + 64) * VEC, c);
        st_v4(dst + (int64_t)(v + 96) * VEC, d);
      }
      uint4 t[3];
      int cnt = 0;
#pragma unroll
      for (int j = 0; j < 3; ++j)
        if (v + 32 * j < nvec) { t[j] = ld_nc_v4(src + (int64_t)(v + 32 * j) * VEC); cnt = j + 1; }
#pragma unroll
      for (int j = 0; j < 3; ++j)
        if (j < cnt) st_v4(dst + (int64_t)(v + 32 * j) * VEC, t[j]);
    } else {
      for (int64_t c = lane; c < m.num_col; c += 32) dst[c] = __ldg(src + c);
    }
  }
}

template <typename T, bool VECOK>
__global__ void __launch_bounds__(256)
add_rows_red_kernel(const __grid_constant__ RowMapDev m, const int64_t* __restrict__ ids,
                    int64_t k, const T* __restrict__ vals, int64_t vals_ld, float sign) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = warp; i < k; i += nwarps) {
    const int64_t r = ids[i];
    if (r < 0 || r >= m.num_row) continue;
    int owner;
    int64_t local;
    locate_row(m, r, owner, local);
    T* dst = reinterpret_cast<T*>(m.shard[owner]) + local * m.num_col;
    const T* src = vals + i * vals_ld;
    if constexpr (VECOK && std::is_same<T, float>::value) {
      const int nvec = (int)(m.num_col / 4);
      for (int v = lane; v < nvec; v += 32) {
        float4 x = *reinterpret_cast<const float4*>(src + (int64_t)v * 4);
        x.x *= sign; x.y *= sign; x.z *= sign; x.w *= sign;
        red_add_v4_f32(reinterpret_cast<float*>(dst) + (int64_t)v * 4, x);
      }
    } else {
      for (int64_t c = lane; c < m.num_col; c += 32) red_add(dst + c, (T)(src[c] * (T)sign));
    }
  }
}

// AddDeltaParameter fused (Applications/WordEmbedding/src/communicator.cpp:157-203): push
// (trained - pulled) * scale for every cached row straight into the owners, no delta tensor.
__global__ void __launch_bounds__(256)
add_rows_delta_kernel(const __grid_constant__ RowMapDev m, const int64_t* __restrict__ ids, int64_t k,
                      const float* __restrict__ cur, const float* __restrict__ old, int64_t ld, float scale) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int nvec = (int)(m.num_col / 4);
  for (int64_t i = warp; i < k; i += nwarps) {
    const int64_t r = ids[i];
    if (r < 0 || r >= m.num_row) continue;
    int owner;
    int64_t local;
    locate_row(m, r, owner, local);
    float* dst = reinterpret_cast<float*>(m.shard[owner]) + local * m.num_col;
    const float4* a = reinterpret_cast<const float4*>(cur + i * ld);
    const float4* b = reinterpret_cast<const float4*>(old + i * ld);
    for (int v = lane; v < nvec; v += 32) {
      float4 x = a[v], y = b[v];
      x.x = (x.x - y.x) * scale; x.y = (x.y - y.y) * scale;
      x.z = (x.z - y.z) * scale; x.w = (x.w - y.w) * scale;
      if (x.x != 0.f || x.y != 0.f || x.z != 0.f || x.w != 0.f) red_add_v4_f32(dst + (int64_t)v * 4, x);
    }
  }
}

template <int UPD, typename T>
__global__ void __launch_bounds__(256)
add_rows_owner_kernel(T* __restrict__ shard, T* __restrict__ st0, T* __restrict__ st1,
                      int64_t row_lo, int64_t row_hi, int64_t num_col, int64_t state_stride,
                      const int64_t* __restrict__ ids, int64_t k, const T* __restrict__ vals,
                      int64_t vals_ld, MvbAddOpt opt) {
  using U = Updater<UPD, T>;
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t woff = U::kPerWorker ? (int64_t)opt.worker_id * state_stride : 0;
  for (int64_t i = warp; i < k; i += nwarps) {
    const int64_t r = ids[i];
    if (r < row_lo || r >= row_hi) continue;
    const int64_t base = (r - row_lo) * num_col;
    const T* src = vals + i * vals_ld;
    for (int64_t c = lane; c < num_col; c += 32) {
      T d = shard[base + c], s0 = (T)0, s1 = (T)0;
      if constexpr (U::kStates >= 1) s0 = st0[woff + base + c];
      if constexpr (U::kStates >= 2) s1 = st1[woff + base + c];
      U::Apply(d, src[c], s0, s1, opt);
      if constexpr (U::kStates >= 1) st0[woff + base + c] = s0;
      if constexpr (U::kStates >= 2) st1[woff + base + c] = s1;
      shard[base + c] = d;
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
row_nonzero_kernel(const T* __restrict__ data, int64_t rows, int64_t cols, int64_t ld,
                   uint8_t* __restrict__ mask) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < rows; r += nwarps) {
    const T* row = data + r * ld;
    int nz = 0;
    for (int64_t c = lane; c < cols; c += 32) nz |= (row[c] != (T)0);
    nz = __any_sync(0xffffffffu, nz);
    if (lane == 0) mask[r] = (uint8_t)(nz ? 1 : 0);
  }
}

__global__ void stale_mark_kernel(uint8_t* stale, int64_t rows, int nworkers, const int64_t* ids,
                                  int64_t k) {
  const int64_t n = (k < 0 ? rows : k) * nworkers;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride) {
    const int64_t per = (k < 0 ? rows : k);
    const int w = (int)(t / per);
    const int64_t j = t - (int64_t)w * per;
    const int64_t r = k < 0 ? j : ids[j];
    if (r >= 0 && r < rows) stale[(int64_t)w * rows + r] = 1;
  }
}
__global__ void stale_take_kernel(uint8_t* stale_w, int64_t rows, const int64_t* ids, int64_t k,
                                  uint8_t* out) {
  const int64_t n = k < 0 ? rows : k;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride) {
    const int64_t r = k < 0 ? t : ids[t];
    if (r >= 0 && r < rows) {
      out[t] = stale_w[r];
      stale_w[r] = 0;
    } else {
      out[t] = 0;
    }
  }
}

// ordered compaction of a byte mask into the list of set positions (one CTA: chunk per thread, block scan)
__global__ void __launch_bounds__(1024)
mask_compact_kernel(const uint8_t* __restrict__ mask, int64_t n, int64_t* __restrict__ out, int64_t* __restrict__ count) {
  __shared__ int64_t part[1024];
  const int64_t per = (n + 1023) / 1024;
  const int64_t lo = (int64_t)threadIdx.x * per, hi = lo + per < n ? lo + per : n;
  int64_t c = 0;
  for (int64_t i = lo; i < hi; ++i) c += mask[i] != 0;
  part[threadIdx.x] = c;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {
    const int64_t t = threadIdx.x >= o ? part[threadIdx.x - o] : 0;
    __syncthreads();
    part[threadIdx.x] += t;
    __syncthreads();
  }
  int64_t pos = part[threadIdx.x] - c;
  for (int64_t i = lo; i < hi; ++i)
    if (mask[i]) out[pos++] = i;
  if (threadIdx.x == 1023) *count = part[1023];
}

RowMapDev to_dev(const MvbRowMap* m) {
  RowMapDev d{};
  d.num_row = m->num_row;
  d.num_col = m->num_col;
  d.S = m->nservers;
  d.rps = m->rows_per_server > 0 ? m->rows_per_server : 1;
  for (int s = 0; s < MVB_MAX_RANKS; ++s) d.shard[s] = s < m->nservers ? m->shard_ptrs[s] : nullptr;
  return d;
}

int grid_for_warps(int64_t warps_needed, int threads) {
  int64_t blocks = (warps_needed * 32 + threads - 1) / threads;
  int64_t cap = (int64_t)mvb_num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

template <typename T>
int get_rows_t(const MvbRowMap* m, const int64_t* ids, int64_t k, void* out, int64_t out_ld,
               cudaStream_t st) {
  RowMapDev d = to_dev(m);
  bool vec = (m->num_col % VecOf<T>::N == 0) && (out_ld % VecOf<T>::N == 0) &&
             ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  int grid = grid_for_warps(k, 256);
  if (vec)
    get_rows_kernel<T, true><<<grid, 256, 0, st>>>(d, ids, k, (T*)out, out_ld);
  else
    get_rows_kernel<T, false><<<grid, 256, 0, st>>>(d, ids, k, (T*)out, out_ld);
  MVB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

template <typename T>
int add_rows_red_t(const MvbRowMap* m, const int64_t* ids, int64_t k, const void* vals,
                   int64_t vals_ld, float sign, cudaStream_t st) {
  RowMapDev d = to_dev(m);
  bool vec = (m->num_col % 4 == 0) && (vals_ld % 4 == 0) &&
             ((reinterpret_cast<uintptr_t>(vals) & 15) == 0) && sizeof(T) == 4;
  int grid = grid_for_warps(k, 256);
  if (vec)
    add_rows_red_kernel<T, true><<<grid, 256, 0, st>>>(d, ids, k, (const T*)vals, vals_ld, sign);
  else
    add_rows_red_kernel<T, false><<<grid, 256, 0, st>>>(d, ids, k, (const T*)vals, vals_ld, sign);
  MVB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

template <typename T>
int add_rows_owner_t(int upd, void* shard, void* s0, void* s1, int64_t lo, int64_t hi,
                     int64_t cols, int64_t sstride, const int64_t* ids, int64_t k,
                     const void* vals, int64_t vld, const MvbAddOpt* opt, cudaStream_t st) {
  int grid = grid_for_warps(k, 256);
#define MVB_ROWS_OWNER(U)                                                                  \
  add_rows_owner_kernel<U, T><<<grid, 256, 0, st>>>((T*)shard, (T*)s0, (T*)s1, lo, hi, cols, \
                                                    sstride, ids, k, (const T*)vals, vld, *opt)
  if constexpr (std::is_same<T, int>::value) {
    MVB_ROWS_OWNER(MVB_UPD_DEFAULT);
  } else {
    switch (upd) {
      case MVB_UPD_DEFAULT: MVB_ROWS_OWNER(MVB_UPD_DEFAULT); break;
      case MVB_UPD_SGD: MVB_ROWS_OWNER(MVB_UPD_SGD); break;
      case MVB_UPD_MOMENTUM: MVB_ROWS_OWNER(MVB_UPD_MOMENTUM); break;
      case MVB_UPD_ADAGRAD: MVB_ROWS_OWNER(MVB_UPD_ADAGRAD); break;
      case MVB_UPD_DCASGD: MVB_ROWS_OWNER(MVB_UPD_DCASGD); break;
      case MVB_UPD_DCASGDA: MVB_ROWS_OWNER(MVB_UPD_DCASGDA); break;
      default: return -2;
    }
  }
#undef MVB_ROWS_OWNER
  MVB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace

extern "C" int mvb_get_rows(int dtype, const MvbRowMap* m, const int64_t* row_ids, int64_t k,
                            void* out, int64_t out_ld, void* stream) {
  if (k <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case MVB_F32: return get_rows_t<float>(m, row_ids, k, out, out_ld, st);
    case MVB_F64: return get_rows_t<double>(m, row_ids, k, out, out_ld, st);
    case MVB_I32: return get_rows_t<int>(m, row_ids, k, out, out_ld, st);
  }
  return -1;
}
extern "C" int mvb_add_rows_red(int dtype, const MvbRowMap* m, const int64_t* row_ids, int64_t k,
                                const void* vals, int64_t vals_ld, float sign, void* stream) {
  if (k <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case MVB_F32: return add_rows_red_t<float>(m, row_ids, k, vals, vals_ld, sign, st);
    case MVB_F64: return add_rows_red_t<double>(m, row_ids, k, vals, vals_ld, sign, st);
    case MVB_I32: return add_rows_red_t<int>(m, row_ids, k, vals, vals_ld, sign, st);
  }
  return -1;
}
extern "C" int mvb_add_rows_delta(const MvbRowMap* m, const int64_t* row_ids, int64_t k, const float* cur,
                                  const float* old, int64_t ld, float scale, void* stream) {
  if (k <= 0) return 0;
  if (m->num_col % 4 || ld % 4) return -9;
  RowMapDev d = to_dev(m);
  add_rows_delta_kernel<<<grid_for_warps(k, 256), 256, 0, (cudaStream_t)stream>>>(d, row_ids, k, cur, old, ld, scale);
  MVB_CUDA_CHECK(cudaGetLastError());
  return 0;
}
extern "C" int mvb_add_rows_owner(int dtype, int updater, void* shard, void* state0, void* state1,
                                  int64_t row_lo, int64_t row_hi, int64_t num_col,
                                  int64_t state_stride, const int64_t* row_ids, int64_t k,
                                  const void* vals, int64_t vals_ld, const MvbAddOpt* opt,
                                  void* stream) {
  if (k <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case MVB_F32:
      return add_rows_owner_t<float>(updater, shard, state0, state1, row_lo, row_hi, num_col,
                                     state_stride, row_ids, k, vals, vals_ld, opt, st);
    case MVB_F64:
      return add_rows_owner_t<double>(updater, shard, state0, state1, row_lo, row_hi, num_col,
                                      state_stride, row_ids, k, vals, vals_ld, opt, st);
    case MVB_I32:
      return add_rows_owner_t<int>(updater, shard, state0, state1, row_lo, row_hi, num_col,
                                   state_stride, row_ids, k, vals, vals_ld, opt, st);
  }
  return -1;
}
extern "C" int mvb_row_nonzero_mask(int dtype, const void* data, int64_t rows, int64_t cols,
                                    int64_t ld, uint8_t* mask, void* stream) {
  if (rows <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  int grid = grid_for_warps(rows, 256);
  switch (dtype) {
    case MVB_F32: row_nonzero_kernel<float><<<grid, 256, 0, st>>>((const float*)data, rows, cols, ld, mask); break;
    case MVB_F64: row_nonzero_kernel<double><<<grid, 256, 0, st>>>((const double*)data, rows, cols, ld, mask); break;
    case MVB_I32: row_nonzero_kernel<int><<<grid, 256, 0, st>>>((const int*)data, rows, cols, ld, mask); break;
    default: return -1;
  }
  MVB_CUDA_CHECK(cudaGetLastError());
  return 0;
}
extern "C" int mvb_stale_mark(uint8_t* stale, int64_t rows, int nworkers, const int64_t* row_ids,
                              int64_t k, void* stream) {
  int64_t n = (k < 0 ? rows : k) * nworkers;
  if (n <= 0) return 0;
  int blocks = (int)((n + 255) / 256 < 4096 ? (n + 255) / 256 : 4096);
  stale_mark_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(stale, rows, nworkers, row_ids, k);
  MVB_CUDA_CHECK(cudaGetLastError());
  return 0;
}
extern "C" int mvb_stale_take(uint8_t* stale_w, int64_t rows, const int64_t* row_ids, int64_t k,
                              uint8_t* out_mask, void* stream) {
  int64_t n = k < 0 ? rows : k;
  if (n <= 0) return 0;
  int blocks = (int)((n + 255) / 256 < 4096 ? (n + 255) / 256 : 4096);
  stale_take_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(stale_w, rows, row_ids, k, out_mask);
  MVB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

// Positions of the non-zero bytes of mask[0..n), ascending; *count = how many (device).
extern "C" int mvb_mask_compact(const uint8_t* mask, int64_t n, int64_t* out_ids, int64_t* count, void* stream) {
  if (n <= 0) return 0;
  mask_compact_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(mask, n, out_ids, count);
  MVB_CUDA_CHECK(cudaGetLastError());
  return 0;
}
