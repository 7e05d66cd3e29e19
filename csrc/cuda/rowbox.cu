// multiverso-b200 :: row mailboxes -- the device-side row Add with OWNER-SIDE apply.
//
// Reference: a row Add travels as a message [ids, packed rows, AddOption] to the owning server, which
// applies the updater once per row, in arrival order, with no lockstep between workers
// (src/table/matrix_table.cpp:266-313 Partition, :403-412 ProcessAdd, src/server.cpp:48-58).
//
// Here the "message" is a mailbox slot in the owner's HBM, one per (owner, source worker), mapped
// into every peer (symmetric allocation):
//
//     slot = [ header 128 B : seq (doorbell), count, AddOption ][ ids : cap x int32 ][ rows : cap x ld ]
//     (two slots per pair, alternating with the epoch parity)
//
//   push   (source, any stream)   the worker's sorted id list is split by owner with one binary search
//          per owner (seg kernel); warps write (trained - pulled) * scale -- or the caller's values --
//          straight into the owners' slots with 128-bit PLAIN stores over NVLink, the last CTA
//          publishes count + the worker's AddOption and rings the doorbell (st.release.sys seq = epoch).
//          Before touching a slot the kernel waits for the owner's ack of its previous push (a flag
//          in the SOURCE's memory, written by the owner), so a slot is never overwritten unread.
//   apply  (owner, its own stream) per source: acquire the doorbell (spin with watchdog, or -- poll
//          mode -- return immediately when nothing new arrived), then warp-per-row plain
//          read-modify-write of the shard through the Updater functor with the SOURCE's AddOption:
//          every (worker,row) is applied exactly once, stateful updaters included; the last CTA bumps
//          the device-side `applied` counter and acks.
//
// No atomics on the data path (co-scheduled under K7 the red.add push ran at ~90 GB/s, the plain-store
// gather at ~800 GB/s), no host barrier, no all_gather_object; workers may push different numbers
// of times -- the owner drains whatever has arrived.  The kernels use no shared memory and <= 40
// registers so they co-reside with the persistent K7 CTAs on every SM.
#include "mvb_common.cuh"

namespace {

constexpr int kHdrBytes = 128;

struct BoxHdr {
  uint64_t seq;        // doorbell: epoch of the push whose data is complete
  int32_t count;       // rows in this push
  int32_t pad0;
  MvbAddOpt opt;       // the pushing worker's AddOption (20 bytes)
};
static_assert(sizeof(BoxHdr) <= kHdrBytes, "header");

struct BoxDev {
  int me, S;
  int64_t num_row, rps, cap;
  int64_t ld_bytes;            // row pitch inside a slot == shard row pitch
  int64_t slot_bytes, ids_bytes;
  unsigned char* box[MVB_MAX_RANKS];     // rank r's mailbox slab (S slots: one per source)
  uint64_t* ack[MVB_MAX_RANKS];          // rank r's ack array [S]
  int* seg;                              // [S + 1] segment starts of the sorted id list (device scratch)
  unsigned int* done;                    // [2] grid completion counters (push, apply)
  uint64_t* applied;                     // [S] epochs applied per source (owner side, device)
  int* go;                               // [S] poll result: 1 = a new push of that source is complete
  int* err;
  long long budget;
};

// Two slots per (owner, source), used alternately (epoch parity): a push only needs the ack of the push
// before last, so an owner that looks into its mailboxes once per step never stalls a worker that pushes
// once per step, whatever their relative timing -- and two ranks can never wait for each other's ack.
constexpr int kSlots = 2;
MVB_DEVINL unsigned char* slot_of(const BoxDev& b, int owner, int source, uint64_t epoch) {
  return b.box[owner] + ((size_t)source * kSlots + (size_t)(epoch & 1)) * b.slot_bytes;
}
MVB_DEVINL int owner_of(const BoxDev& b, int64_t r) {
  int64_t o = r / b.rps;
  if (o > b.S - 1) o = b.S - 1;
  return (int)o;
}

// seg[o] = first position of the sorted id list that belongs to owner o
__global__ void rowbox_seg_kernel(const __grid_constant__ BoxDev b, const int* __restrict__ ids,
                                  const int* __restrict__ n_ptr, int64_t n_max) {
  int64_t n = n_ptr ? (int64_t)*n_ptr : n_max;
  if (n > n_max) n = n_max;
  const int o = threadIdx.x;
  if (o > b.S) return;
  if (o == b.S) { b.seg[o] = (int)n; return; }
  const int64_t key = (int64_t)o * b.rps;          // first row of owner o
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if ((int64_t)ids[mid] < key) lo = mid + 1; else hi = mid;
  }
  b.seg[o] = (int)lo;
}

struct PushArgs {
  const int* ids;
  const int* n_ptr;
  int64_t n_max;
  const unsigned char* cur;    // DELTA: trained rows; else: values
  const unsigned char* old;    // DELTA: pulled rows
  int64_t src_ld_bytes;
  float scale;
  uint64_t epoch;
  MvbAddOpt opt;
  int row_bytes;
};

// 128-thread CTAs, <= 56 registers: one warp per SM sub-partition fits beside the persistent K7 CTA (we_block.cu)
constexpr int kBoxThreads = 128;

template <bool DELTA>
__global__ void __launch_bounds__(kBoxThreads, 9)
rowbox_push_kernel(const __grid_constant__ BoxDev b, const __grid_constant__ PushArgs a) {
  __shared__ int s_flag;
  // the owners must have consumed my previous push before its slot is overwritten
  if (threadIdx.x == 0) s_flag = 0;
  __syncthreads();
  if (threadIdx.x < b.S && a.epoch > kSlots) {
    if (!spin_wait_ge(b.ack[b.me] + threadIdx.x, a.epoch - kSlots, b.budget)) {
      if (b.err) atomicExch(b.err, 3000 + threadIdx.x);
      s_flag = 1;
    }
  }
  __syncthreads();
  if (s_flag) return;
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  int64_t n = a.n_ptr ? (int64_t)*a.n_ptr : a.n_max;
  if (n > a.n_max) n = a.n_max;
  const int nvec = a.row_bytes >> 4;
  for (int64_t s = warp; s < n; s += nwarps) {
    const int64_t r = (int64_t)__ldg(a.ids + s);
    if (r < 0 || r >= b.num_row) continue;
    const int o = owner_of(b, r);
    const int64_t idx = s - b.seg[o];
    if (idx < 0 || idx >= b.cap) {
      if (lane == 0 && b.err) atomicExch(b.err, 3100 + o);
      continue;
    }
    unsigned char* slot = slot_of(b, o, b.me, a.epoch);
    if (lane == 0) reinterpret_cast<int*>(slot + kHdrBytes)[idx] = (int)r;
    unsigned char* dst = slot + kHdrBytes + b.ids_bytes + (size_t)idx * b.ld_bytes;
    const unsigned char* c = a.cur + s * a.src_ld_bytes;
    const unsigned char* od = DELTA ? a.old + s * a.src_ld_bytes : nullptr;
    for (int v0 = 0; v0 < nvec; v0 += 96) {
      float4 x[3], y[3];
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const int v = v0 + lane + 32 * j;
        if (v < nvec) {
          x[j] = *reinterpret_cast<const float4*>(c + (size_t)v * 16);
          if (DELTA) y[j] = *reinterpret_cast<const float4*>(od + (size_t)v * 16);
        }
      }
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const int v = v0 + lane + 32 * j;
        if (v < nvec) {
          float4 d;
          if (DELTA) {
            d.x = (x[j].x - y[j].x) * a.scale; d.y = (x[j].y - y[j].y) * a.scale;
            d.z = (x[j].z - y[j].z) * a.scale; d.w = (x[j].w - y[j].w) * a.scale;
          } else {
            d.x = x[j].x * a.scale; d.y = x[j].y * a.scale; d.z = x[j].z * a.scale; d.w = x[j].w * a.scale;
          }
          st_na_v4(dst + (size_t)v * 16, *reinterpret_cast<uint4*>(&d));
        }
      }
    }
  }
  // ---- completion: the last CTA publishes counts + options and rings the doorbells ----
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    s_flag = (atomicAdd(b.done + 0, 1u) == gridDim.x - 1) ? 2 : 0;
  }
  __syncthreads();
  if (s_flag != 2) return;
  __threadfence_system();
  if (threadIdx.x < b.S) {
    const int o = threadIdx.x;
    BoxHdr* h = reinterpret_cast<BoxHdr*>(slot_of(b, o, b.me, a.epoch));
    h->count = b.seg[o + 1] - b.seg[o];
    h->opt = a.opt;
    fence_sys();
    st_release_sys_u64(&h->seq, a.epoch);
  }
  if (threadIdx.x == 0) b.done[0] = 0u;
}

struct ApplyArgs {
  float* shard;
  float* st0;
  float* st1;
  int64_t state_stride;      // elements per worker slab of per-worker state
  int64_t row_lo;
  int src;
  int row_floats;
};

// One decision per (launch, source), taken by ONE thread: go[w] = "the next push of worker w is complete".
// (Deciding per CTA would let early CTAs skip and late CTAs apply the same push.)
__global__ void rowbox_poll_kernel(const __grid_constant__ BoxDev b, int wait, int only_src) {
  const int w = threadIdx.x;
  if (w >= b.S) return;
  if (only_src >= 0 && w != only_src) return;
  const uint64_t want = b.applied[w] + 1;
  BoxHdr* h = reinterpret_cast<BoxHdr*>(slot_of(b, b.me, w, want));
  int go;
  if (wait) {
    go = 1;
    if (!spin_wait_ge(&h->seq, want, b.budget)) {
      if (b.err) atomicExch(b.err, 3200 + w);
      go = 0;
    }
  } else {
    go = ld_acquire_sys_u64(&h->seq) >= want ? 1 : 0;
  }
  b.go[w] = go;
}

template <int UPD>
__global__ void __launch_bounds__(kBoxThreads, 9)
rowbox_apply_kernel(const __grid_constant__ BoxDev b, const __grid_constant__ ApplyArgs a) {
  using U = Updater<UPD, float>;
  __shared__ int s_go;
  if (!b.go[a.src]) return;                          // same answer in every CTA (written by the poll kernel)
  const uint64_t want = b.applied[a.src] + 1;        // only the last CTA of a launch bumps `applied`
  unsigned char* slot = slot_of(b, b.me, a.src, want);
  BoxHdr* h = reinterpret_cast<BoxHdr*>(slot);
  const int count = __ldcg(&h->count);
  MvbAddOpt opt;
  {
    const int* po = reinterpret_cast<const int*>(&h->opt);
    int* pd = reinterpret_cast<int*>(&opt);
#pragma unroll
    for (int k = 0; k < (int)(sizeof(MvbAddOpt) / 4); ++k) pd[k] = __ldcg(po + k);
  }
  opt.worker_id = a.src;
  const int* ids = reinterpret_cast<const int*>(slot + kHdrBytes);
  const unsigned char* vals = slot + kHdrBytes + b.ids_bytes;
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t woff = U::kPerWorker ? (int64_t)a.src * a.state_stride : 0;
  const int nvec = a.row_floats >> 2;
  for (int64_t i = warp; i < count; i += nwarps) {
    const int64_t r = (int64_t)__ldcg(ids + i);
    const int64_t base = (r - a.row_lo) * (int64_t)a.row_floats;
    if (r < a.row_lo || base < 0) continue;
    const float4* src = reinterpret_cast<const float4*>(vals + (size_t)i * b.ld_bytes);
    for (int v = lane; v < nvec; v += 32) {
      const float4 g = __ldcg(src + v);
      float4* dp = reinterpret_cast<float4*>(a.shard + base) + v;
      float4 d = *dp;
      float4 s0 = make_float4(0, 0, 0, 0), s1 = make_float4(0, 0, 0, 0);
      float4* s0p = nullptr;
      float4* s1p = nullptr;
      if constexpr (U::kStates >= 1) { s0p = reinterpret_cast<float4*>(a.st0 + woff + base) + v; s0 = *s0p; }
      if constexpr (U::kStates >= 2) { s1p = reinterpret_cast<float4*>(a.st1 + woff + base) + v; s1 = *s1p; }
      U::Apply(d.x, g.x, s0.x, s1.x, opt);
      U::Apply(d.y, g.y, s0.y, s1.y, opt);
      U::Apply(d.z, g.z, s0.z, s1.z, opt);
      U::Apply(d.w, g.w, s0.w, s1.w, opt);
      *dp = d;
      if constexpr (U::kStates >= 1) *s0p = s0;
      if constexpr (U::kStates >= 2) *s1p = s1;
    }
  }
  // ---- completion: the last CTA records the epoch and acks the source ----
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    s_go = (atomicAdd(b.done + 1, 1u) == gridDim.x - 1) ? 2 : 0;
  }
  __syncthreads();
  if (s_go != 2 || threadIdx.x != 0) return;
  b.applied[a.src] = want;
  b.done[1] = 0u;
  fence_sys();
  st_release_sys_u64(b.ack[a.src] + b.me, want);
}

BoxDev to_dev(const MvbRowBox* h) {
  BoxDev b{};
  b.me = h->me;
  b.S = h->map.nservers;
  b.num_row = h->map.num_row;
  b.rps = h->map.rows_per_server > 0 ? h->map.rows_per_server : 1;
  b.cap = h->cap;
  b.ld_bytes = h->map.num_col * 4;
  b.ids_bytes = (h->cap * 4 + 127) / 128 * 128;
  b.slot_bytes = h->slot_bytes;
  for (int r = 0; r < MVB_MAX_RANKS; ++r) {
    b.box[r] = r < b.S ? reinterpret_cast<unsigned char*>(h->box[r]) : nullptr;
    b.ack[r] = r < b.S ? reinterpret_cast<uint64_t*>(h->ack[r]) : nullptr;
  }
  b.seg = h->seg;
  b.done = h->done;
  b.applied = h->applied;
  b.go = h->go;
  b.err = h->err_flag;
  int dev = 0, khz = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  const double ts = h->timeout_s > 0 ? h->timeout_s : 120.0;
  b.budget = (long long)(ts * 1e3 * (khz > 0 ? khz : 1965000));
  return b;
}

int grid_of(int ctas_per_sm) { return mvb_num_sms() * (ctas_per_sm > 0 ? ctas_per_sm : 1); }

}  // namespace

// bytes of one (owner, source) slot for `cap` rows of `num_col` floats
extern "C" int mvb_rowbox_slots(void) { return kSlots; }
extern "C" int64_t mvb_rowbox_slot_bytes(int64_t cap, int64_t num_col) {
  return kHdrBytes + (cap * 4 + 127) / 128 * 128 + cap * num_col * 4;
}

static int push_common(const MvbRowBox* h, bool delta, const int* ids, const int* n_ptr, int64_t n_max,
                       const float* cur, const float* old, int64_t ld, float scale, uint64_t epoch,
                       const MvbAddOpt* opt, int ctas_per_sm, void* stream) {
  if (h->map.num_col % 4 || ld % 4) return -9;
  cudaStream_t st = (cudaStream_t)stream;
  BoxDev b = to_dev(h);
  rowbox_seg_kernel<<<1, 32, 0, st>>>(b, ids, n_ptr, n_max);
  PushArgs a{};
  a.ids = ids; a.n_ptr = n_ptr; a.n_max = n_max;
  a.cur = reinterpret_cast<const unsigned char*>(cur);
  a.old = reinterpret_cast<const unsigned char*>(old);
  a.src_ld_bytes = ld * 4; a.scale = scale; a.epoch = epoch;
  if (opt) a.opt = *opt;
  a.row_bytes = (int)(h->map.num_col * 4);
  int grid = grid_of(ctas_per_sm);
  const int64_t need = (n_max + 3) / 4;
  if ((int64_t)grid > need) grid = (int)(need > 0 ? need : 1);
  if (delta) rowbox_push_kernel<true><<<grid, kBoxThreads, 0, st>>>(b, a);
  else rowbox_push_kernel<false><<<grid, kBoxThreads, 0, st>>>(b, a);
  MVB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

// AddDeltaParameter through the mailboxes: slot rows = (cur - old) * scale. `ids` ascending.
extern "C" int mvb_rowbox_push_delta(const MvbRowBox* h, const int* ids, const int* n_ptr, int64_t n_max,
                                     const float* cur, const float* old, int64_t ld, float scale, uint64_t epoch,
                                     const MvbAddOpt* opt, int ctas_per_sm, void* stream) {
  return push_common(h, true, ids, n_ptr, n_max, cur, old, ld, scale, epoch, opt, ctas_per_sm, stream);
}
// row Add of caller-provided values (x scale) through the mailboxes. `ids` ascending.
extern "C" int mvb_rowbox_push_vals(const MvbRowBox* h, const int* ids, const int* n_ptr, int64_t n_max,
                                    const float* vals, int64_t ld, float scale, uint64_t epoch, const MvbAddOpt* opt,
                                    int ctas_per_sm, void* stream) {
  return push_common(h, false, ids, n_ptr, n_max, vals, nullptr, ld, scale, epoch, opt, ctas_per_sm, stream);
}

// Owner side, step 1: for every source (src < 0) or one source decide whether its next push is complete
// (wait = 1: spin for it with the watchdog; wait = 0: just look).  Step 2: mvb_rowbox_apply per source.
extern "C" int mvb_rowbox_poll(const MvbRowBox* h, int src, int wait, void* stream) {
  BoxDev b = to_dev(h);
  rowbox_poll_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(b, wait, src);
  MVB_CUDA_CHECK(cudaGetLastError());
  return 0;
}

// Owner side, step 2: apply the push of worker `src` that the preceding poll found complete (no-op otherwise).
extern "C" int mvb_rowbox_apply(const MvbRowBox* h, int updater, float* shard, float* st0, float* st1,
                                int64_t state_stride, int64_t row_lo, int src, int ctas_per_sm,
                                void* stream) {
  if (h->map.num_col % 4) return -9;
  BoxDev b = to_dev(h);
  ApplyArgs a{};
  a.shard = shard; a.st0 = st0; a.st1 = st1; a.state_stride = state_stride; a.row_lo = row_lo;
  a.src = src; a.row_floats = (int)h->map.num_col;
  const int grid = grid_of(ctas_per_sm);
  cudaStream_t st = (cudaStream_t)stream;
  switch (updater) {
    case MVB_UPD_DEFAULT: rowbox_apply_kernel<MVB_UPD_DEFAULT><<<grid, kBoxThreads, 0, st>>>(b, a); break;
    case MVB_UPD_SGD: rowbox_apply_kernel<MVB_UPD_SGD><<<grid, kBoxThreads, 0, st>>>(b, a); break;
    case MVB_UPD_MOMENTUM: rowbox_apply_kernel<MVB_UPD_MOMENTUM><<<grid, kBoxThreads, 0, st>>>(b, a); break;
    case MVB_UPD_ADAGRAD: rowbox_apply_kernel<MVB_UPD_ADAGRAD><<<grid, kBoxThreads, 0, st>>>(b, a); break;
    case MVB_UPD_DCASGD: rowbox_apply_kernel<MVB_UPD_DCASGD><<<grid, kBoxThreads, 0, st>>>(b, a); break;
    case MVB_UPD_DCASGDA: rowbox_apply_kernel<MVB_UPD_DCASGDA><<<grid, kBoxThreads, 0, st>>>(b, a); break;
    default: return -2;
  }
  MVB_CUDA_CHECK(cudaGetLastError());
  return 0;
}
