// multiverso-b200 :: K7 (window-batched variant) -- skip-gram negative sampling, one CENTRE POSITION
// per consumer warp, row traffic on the bulk-copy (TMA) engine.
//
// The pair-at-a-time kernels (sgns.cu, sgns_tma.cu) move 7 rows in and 7 rows out for every
// (context, centre) pair: the centre row and the K negatives are re-fetched and re-reduced for each
// of the ~6 contexts of a position and every input row is fetched once per centre it is a context of.
// This variant keeps the reference's sample schedule (ParseSentence, wordembedding.cpp:216-257: random
// window shrink, contexts stop at sentence breaks, one sample per context word, target == centre
// skipped) but batches the maths per centre position (the mini-batch formulation of pWord2Vec /
// "Parallelizing Word2Vec in Shared and Distributed Memory": the K negatives are drawn once per
// position and shared by its contexts, scores are a [contexts x D]·[D x (1+K)] tile evaluated on
// pre-update rows):
//
//   producer warps walk the CTA's contiguous token range in order, as two independent streams.  Warp 0,
//                   per position: ONE cp.async.bulk of the input row into a CTA-wide ring (every row is
//                   a context of up to 2W centres, so it crosses L2 once instead of ~6 times); one lane
//                   per position waits for its slot and issues its copy.  Warp 1, per position: 1+K
//                   cp.async.bulk of the centre / negative output rows into the owning consumer's
//                   double-buffered stage.  Negative sampling (hash RNG -> alias table or block pool)
//                   and the id -> cache-slot maps are evaluated 32 positions at a time, one lane per
//                   position, so their dependent loads are paid once per batch.
//   consumer warps  position p -> warp (p mod NW).  Phase 1 (output rows in registers): per context
//                   1+K dots, sigmoid / error terms, the context's input delta sum_k g_k·out_k goes
//                   through a small staging ring and leaves with ONE cp.reduce.async.bulk.add.f32.
//                   Phase 2: the 1+K output deltas sum_c g_ck·in_c accumulate in registers over the
//                   window, are written over the stage in place and leave with one bulk reduction
//                   per row.  Per word: 7 row loads + (1+K+contexts) ~ 12 row reductions instead of
//                   42 + 42.
//
// Synchronisation is mbarrier only (no __syncthreads after init): in_full/in_empty per ring slot
// (empty counts 2W+1 arrivals: the centres whose window covers the slot), out_full/out_empty per
// consumer stage.  Virtual positions [a-2W, b+2W) make every count uniform at the range ends.
#include <cstdlib>
#include "mvb_common.cuh"

namespace {

constexpr int kNW = 10;            // consumer warps per CTA
constexpr int kThreads = 32 * (2 + kNW);   // two producer warps + the consumers
constexpr int kKO = 8;             // max output rows per position (centre + up to 7 negatives)
constexpr int kDinSlots = 3;       // per-warp staging ring of input-row deltas
constexpr int kRelLanes = 8;       // lanes that issue row reductions / arrive on out_empty

struct WinDev {
  const int* tokens;
  int64_t n_tokens;
  float* w_in;
  float* w_out;
  int dim;
  int64_t ld;
  int window, negative;
  float lr;
  const float* alias_prob;
  const int* alias_idx;
  int vocab;
  const int* neg_pool;
  int neg_pool_size;
  const int* neg_pool_size_ptr;
  const int* map_in;
  const int* map_out;
  const float* scale_in;    // optional per-word step scale (hot-row cap), indexed by word id
  const float* scale_out;
  uint64_t seed;
  float* loss_sum;
  unsigned long long* pair_count;
  int row_bytes;     // 4 * dim
  int ring;          // input-row ring slots (>= 2*kNW + 2*window + 1)
  int nw;            // active consumer warps (<= kNW)
  int ko;            // 1 + negative
  int chunk;         // token positions per work unit (chunk c belongs to CTA c mod gridDim.x)
  // direct mode (Hogwild over NVLink): S > 0 => rows are addressed in the row-sharded tables themselves,
  // owner = id / rps (last server takes the remainder), through the peer mappings of the shards
  int S;
  int64_t rps;
  float* in_peer[MVB_MAX_RANKS];
  float* out_peer[MVB_MAX_RANKS];
};

MVB_DEVINL float* row_of(float* const* peers, int S, int64_t rps, int64_t ld, float* local_base, int64_t rid) {
  if (S <= 0) return local_base + rid * ld;
  int64_t o = rid / rps;
  if (o > S - 1) o = S - 1;
  return peers[o] + (rid - o * rps) * ld;
}


struct OutMeta {
  float* ptr[kKO];   // global row addresses (nullptr = row unused)
  float scale[kKO];  // step scale of the row (1 unless the word is in the capped Zipf head)
  long long p;       // token position of this centre (seeds its window shrink / negatives)
  int active;        // 1 train, 0 virtual / sentence-break position, -1 end of work (consumer exits)
  int pad[5];
};
static_assert(sizeof(OutMeta) == 128, "OutMeta size");

MVB_DEVINL uint64_t hash64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
  return x;
}
MVB_DEVINL uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
MVB_DEVINL void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
MVB_DEVINL void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
MVB_DEVINL void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
MVB_DEVINL void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
MVB_DEVINL void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
MVB_DEVINL void bulk_reduce_add_s2g(void* gdst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;"
               ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
MVB_DEVINL void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
MVB_DEVINL void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
MVB_DEVINL void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// packed fp32x2 math (Blackwell FFMA2): halves the FMA instruction count of dots / axpys
MVB_DEVINL float2 ffma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;"
      : "=l"(reinterpret_cast<uint64_t&>(d))
      : "l"(reinterpret_cast<uint64_t&>(a)), "l"(reinterpret_cast<uint64_t&>(b)), "l"(reinterpret_cast<uint64_t&>(c)));
  return d;
}
MVB_DEVINL float2 fmul2(float2 a, float2 b) {
  float2 d;
  asm("mul.rn.f32x2 %0, %1, %2;"
      : "=l"(reinterpret_cast<uint64_t&>(d))
      : "l"(reinterpret_cast<uint64_t&>(a)), "l"(reinterpret_cast<uint64_t&>(b)));
  return d;
}
MVB_DEVINL float sigm_fast(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
MVB_DEVINL float softplus_neg(float x) { return fmaxf(-x, 0.f) + log1pf(__expf(-fabsf(x))); }

struct F4 {
  float2 lo, hi;
};
MVB_DEVINL F4 lds_f4(const unsigned char* p) {
  float4 t = *reinterpret_cast<const float4*>(p);
  F4 r;
  r.lo = make_float2(t.x, t.y);
  r.hi = make_float2(t.z, t.w);
  return r;
}
MVB_DEVINL void sts_f4(unsigned char* p, F4 v) {
  *reinterpret_cast<float4*>(p) = make_float4(v.lo.x, v.lo.y, v.hi.x, v.hi.y);
}

// A row of `dim` floats spread over the 32 lanes of a warp: NF4 chunks of one float4 per lane (the last
// one possibly partial) and, when 16 or fewer float4 remain (dim 300: 11), a TAIL of one float2 per lane over twice
// as many lanes instead of a half-empty float4 chunk -- 10 registers and 5 FFMA2 per 300-float row and dot
// instead of 12 and 6.
template <int NF4, bool TAIL>
struct Row {
  F4 v[NF4 > 0 ? NF4 : 1];
  float2 t;
};
struct RowAct {
  bool v[4];       // lane holds a float4 of chunk j
  bool t;          // lane holds a float2 of the tail
  int tail_off;    // byte offset of the tail inside a row
};
template <int NF4, bool TAIL>
MVB_DEVINL void row_zero(Row<NF4, TAIL>& r) {
#pragma unroll
  for (int j = 0; j < NF4; ++j) r.v[j].lo = r.v[j].hi = make_float2(0.f, 0.f);
  r.t = make_float2(0.f, 0.f);
}
template <int NF4, bool TAIL>
MVB_DEVINL void row_load(Row<NF4, TAIL>& r, const unsigned char* base, int lane, const RowAct& act) {
  row_zero(r);
#pragma unroll
  for (int j = 0; j < NF4; ++j)
    if (act.v[j]) r.v[j] = lds_f4(base + (size_t)(lane + 32 * j) * 16);
  if (TAIL && act.t) r.t = *reinterpret_cast<const float2*>(base + act.tail_off + lane * 8);
}
template <int NF4, bool TAIL>
MVB_DEVINL void row_store(unsigned char* base, const Row<NF4, TAIL>& r, int lane, const RowAct& act) {
#pragma unroll
  for (int j = 0; j < NF4; ++j)
    if (act.v[j]) sts_f4(base + (size_t)(lane + 32 * j) * 16, r.v[j]);
  if (TAIL && act.t) *reinterpret_cast<float2*>(base + act.tail_off + lane * 8) = r.t;
}
template <int NF4, bool TAIL>
MVB_DEVINL float row_dot(const Row<NF4, TAIL>& x, const Row<NF4, TAIL>& y) {
  float2 acc = make_float2(0.f, 0.f);
#pragma unroll
  for (int j = 0; j < NF4; ++j) {
    acc = ffma2(x.v[j].lo, y.v[j].lo, acc);
    acc = ffma2(x.v[j].hi, y.v[j].hi, acc);
  }
  if (TAIL) acc = ffma2(x.t, y.t, acc);
  return acc.x + acc.y;
}
template <int NF4, bool TAIL>
MVB_DEVINL void row_axpy(Row<NF4, TAIL>& y, float g, const Row<NF4, TAIL>& x) {       // y += g * x
  const float2 gg = make_float2(g, g);
#pragma unroll
  for (int j = 0; j < NF4; ++j) {
    y.v[j].lo = ffma2(gg, x.v[j].lo, y.v[j].lo);
    y.v[j].hi = ffma2(gg, x.v[j].hi, y.v[j].hi);
  }
  if (TAIL) y.t = ffma2(gg, x.t, y.t);
}

// smem carve-up (all offsets from the dynamic smem base, which is 128-byte aligned)
struct Layout {
  int in_full, in_empty, out_full, out_empty;   // mbarrier arrays (byte offsets)
  int in_ptr;                                   // float* [ring]
  int in_scale;                                 // float  [ring]
  int out_meta;                                 // OutMeta [kNW*2]
  int batch;                                    // producer batch scratch
  int in_rows, out_rows, din_rows;
  int total;
};
__host__ __device__ inline Layout make_layout(int ring, int nw, int ko, int row_bytes) {
  Layout L;
  int o = 0;
  L.in_full = o;   o += ring * 8;
  L.in_empty = o;  o += ring * 8;
  L.out_full = o;  o += kNW * 2 * 8;
  L.out_empty = o; o += kNW * 2 * 8;
  L.in_ptr = o;    o += ring * 8;
  L.in_scale = o;  o += ring * 4;
  o = (o + 127) / 128 * 128;
  L.out_meta = o;  o += kNW * 2 * (int)sizeof(OutMeta);
  L.batch = o;     o += 32 * 8 + 32 * kKO * 4 + 32 * 4 + 32 * 4 + 32 * kKO * 4;   // in ptr, out ids, active, scales
  o = (o + 127) / 128 * 128;
  L.in_rows = o;   o += ring * row_bytes;
  L.out_rows = o;  o += nw * 2 * ko * row_bytes;
  L.din_rows = o;  o += nw * kDinSlots * row_bytes;
  L.total = o;
  return L;
}

template <int NF4, bool TAIL, int KO>
__global__ void __launch_bounds__(kThreads, 1)
sgns_win_kernel(const __grid_constant__ WinDev a) {
  using RowT = Row<NF4, TAIL>;
  extern __shared__ __align__(128) unsigned char smem[];
  const Layout L = make_layout(a.ring, a.nw, a.ko, a.row_bytes);
  uint64_t* in_full = reinterpret_cast<uint64_t*>(smem + L.in_full);
  uint64_t* in_empty = reinterpret_cast<uint64_t*>(smem + L.in_empty);
  uint64_t* out_full = reinterpret_cast<uint64_t*>(smem + L.out_full);
  uint64_t* out_empty = reinterpret_cast<uint64_t*>(smem + L.out_empty);
  float** in_ptr = reinterpret_cast<float**>(smem + L.in_ptr);
  float* in_scale = reinterpret_cast<float*>(smem + L.in_scale);
  OutMeta* out_meta = reinterpret_cast<OutMeta*>(smem + L.out_meta);
  unsigned char* in_rows = smem + L.in_rows;
  unsigned char* out_rows = smem + L.out_rows;
  unsigned char* din_rows = smem + L.din_rows;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int W = a.window;
  const int R = a.ring;
  const int NW = a.nw;

  // Work units are chunks of `a.chunk` consecutive token positions, chunk c on CTA c mod gridDim.x (one chunk
  // per CTA unless MVB_WIN_CHUNK says otherwise).  Inside a CTA the chunks form ONE
  // stream of virtual positions G = 0, 1, 2, ...: chunk [ra, rb) contributes rb - ra + 4W of them (local index
  // i <-> centre token ra - 2W + i, ring entry i <-> input token ra - W + i, empty beyond the chunk), so ring
  // slots, stages and barrier phases simply continue across chunks.
  const int64_t n_chunks = (a.n_tokens + a.chunk - 1) / a.chunk;

  if (threadIdx.x == 0) {
    for (int s = 0; s < R; ++s) {
      mbar_init(in_full + s, 1);
      mbar_init(in_empty + s, 2 * W + 1);
    }
    for (int s = 0; s < kNW * 2; ++s) {
      mbar_init(out_full + s, 1);
      mbar_init(out_empty + s, kRelLanes);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp < 2) {
    // ================================ PRODUCERS ========================================
    // Two warps, two independent streams over the same virtual positions: warp 0 feeds the input-row ring,
    // warp 1 samples the negatives and feeds the output stages.  (One warp doing both was the bottleneck of
    // the kernel: the consumers spent half of their samples waiting in out_full while the producer never
    // waited on anything.)
    float** b_inptr = reinterpret_cast<float**>(smem + L.batch);
    int* b_rid = reinterpret_cast<int*>(smem + L.batch + 32 * 8);
    int* b_act = b_rid + 32 * kKO;
    float* b_isc = reinterpret_cast<float*>(b_act + 32);
    float* b_osc = b_isc + 32;
    const int pool_n = (warp == 1 && a.neg_pool_size_ptr) ? __ldg(a.neg_pool_size_ptr) : a.neg_pool_size;
    int64_t G0 = 0;                          // virtual positions emitted by earlier chunks
    for (int64_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
      const int64_t ra = c * (int64_t)a.chunk;
      const int64_t rb = (ra + a.chunk < a.n_tokens) ? ra + a.chunk : a.n_tokens;
      const int64_t len = rb - ra;
      const int64_t nj = len + 2 * W;          // ring entries with a token behind them
      const int64_t nv = len + 4 * W;          // virtual centres of this chunk
      if (warp == 0) {
        // ---- input-row ring: every virtual position owns an entry (empty beyond the chunk) -----
        for (int64_t i0 = 0; i0 < nv; i0 += 32) {
          {
            const int64_t i = i0 + lane;
            const int64_t q = ra - W + i;
            float* iptr = nullptr;
            float isc = 1.f;
            if (i < nj && q >= 0 && q < a.n_tokens) {
              const int tq = __ldg(a.tokens + q);
              if (tq >= 0) {
                const int rid = a.map_in ? __ldg(a.map_in + tq) : tq;
                if (rid >= 0) iptr = row_of(a.in_peer, a.S, a.rps, a.ld, a.w_in, rid);
                if (a.scale_in) isc = __ldg(a.scale_in + tq);
              }
            }
            b_inptr[lane] = iptr;
            b_isc[lane] = isc;
          }
          __syncwarp();
          const int nb = (int)((nv - i0) < 32 ? (nv - i0) : 32);
          // one lane per position: each waits for ITS slot and issues ITS copy (slots of a batch are distinct
          // because R >= 32 is not guaranteed -> walk the batch in groups of at most R lanes)
          for (int l0 = 0; l0 < nb; l0 += R) {
            const int l = l0 + lane;
            if (lane < R && l < nb) {
              const int64_t ii = G0 + i0 + l;
              const int slot = (int)(ii % R);
              const uint32_t round = (uint32_t)(ii / R);
              mbar_wait(in_empty + slot, (round & 1u) ^ 1u);
              // a row nobody needed (shrunk window, break) can be released before its copy landed:
              // never re-arm a slot whose previous transaction is still in flight
              if (round > 0) mbar_wait(in_full + slot, (round - 1u) & 1u);
              float* ptr = (i0 + l < nj) ? b_inptr[l] : nullptr;
              in_ptr[slot] = ptr;
              in_scale[slot] = b_isc[l];
              if (ptr) {
                mbar_arrive_expect_tx(in_full + slot, (uint32_t)a.row_bytes);
                bulk_g2s(in_rows + (size_t)slot * a.row_bytes, ptr, (uint32_t)a.row_bytes, in_full + slot);
              } else {
                mbar_arrive(in_full + slot);
              }
            }
            __syncwarp();
          }
          __syncwarp();       // batch scratch is re-written next round
        }
      } else {
        // ---- centre + negatives of every virtual centre, into the stage of the owning consumer ----
        for (int64_t i0 = 0; i0 < nv; i0 += 32) {
          {
            const int64_t i = i0 + lane;
            const int64_t p = ra - 2 * W + i;
            int tp = -1;
            if (i >= 2 * W && i < 2 * W + len) tp = __ldg(a.tokens + p);
            int rid[kKO];
            float osc[kKO];
#pragma unroll
            for (int k = 0; k < kKO; ++k) { rid[k] = -1; osc[k] = 1.f; }
            if (tp >= 0) {
              const uint64_t prng = hash64(a.seed ^ (uint64_t)(p + 1) * 0x9E3779B97F4A7C15ull);
              int t[kKO];
#pragma unroll
              for (int k = 1; k < kKO; ++k) {
                t[k] = -1;
                if (k < a.ko) {
                  const uint64_t r = hash64(prng ^ ((uint64_t)k * 0xD6E8FEB86659FD93ull));
                  if (a.neg_pool) {
                    t[k] = pool_n > 0 ? __ldg(a.neg_pool + (r >> 8) % (uint64_t)pool_n) : -1;
                  } else {
                    const uint32_t idx = (uint32_t)((r >> 32) % (uint64_t)a.vocab);
                    const float u = (float)(r & 0xFFFFFF) * (1.0f / 16777216.0f);
                    t[k] = (u < __ldg(a.alias_prob + idx)) ? (int)idx : __ldg(a.alias_idx + idx);
                  }
                  if (t[k] == tp) t[k] = -1;                 // Parse(): target == word_idx is skipped
                }
              }
              rid[0] = a.map_out ? __ldg(a.map_out + tp) : tp;
              if (a.scale_out) osc[0] = __ldg(a.scale_out + tp);
#pragma unroll
              for (int k = 1; k < kKO; ++k)
                if (t[k] >= 0) {
                  rid[k] = a.map_out ? __ldg(a.map_out + t[k]) : t[k];
                  if (a.scale_out) osc[k] = __ldg(a.scale_out + t[k]);
                }
            }
#pragma unroll
            for (int k = 0; k < kKO; ++k) { b_rid[lane * kKO + k] = rid[k]; b_osc[lane * kKO + k] = osc[k]; }
            b_act[lane] = (tp >= 0 && rid[0] >= 0) ? 1 : 0;
          }
          __syncwarp();
          const int nb = (int)((nv - i0) < 32 ? (nv - i0) : 32);
          // four positions per pass: lane group g = lane / 8 serves position l4 + g, lane k of the group
          // its k-th output row (kKO <= 8)
          // (never two positions of the same consumer in one pass: the second one waits for a stage that is
          // only released once the first one has been handed over)
          const int g = lane >> 3, k = lane & 7;
          const int PW = NW < 4 ? NW : 4;
          for (int l4 = 0; l4 < nb; l4 += PW) {
            const int l = l4 + g;
            const bool on = g < PW && l < nb;
            const int64_t ii = G0 + i0 + l;                 // position in the CTA's virtual stream
            const int cw = (int)(ii % NW);
            const int64_t n = ii / NW;
            const int st = cw * 2 + (int)(n & 1);
            float* optr = nullptr;
            if (on) {
              mbar_wait(out_empty + st, (uint32_t)((n >> 1) & 1) ^ 1u);
              const int act = b_act[l];
              if (k < kKO && act) {
                const int r = b_rid[l * kKO + k];
                if (r >= 0) optr = row_of(a.out_peer, a.S, a.rps, a.ld, a.w_out, r);
              }
              if (k < kKO) {
                out_meta[st].ptr[k] = optr;
                out_meta[st].scale[k] = b_osc[l * kKO + k];
              }
              if (k == 0) {
                out_meta[st].active = act;
                out_meta[st].p = (long long)(ra - 2 * W + i0 + l);
              }
            }
            const uint32_t have = (__ballot_sync(0xffffffffu, optr != nullptr) >> (g * 8)) & 0xffu;
            __syncwarp();     // every lane's meta pointer is written before the release-arrive
            if (on && k == 0) {
              if (have) mbar_arrive_expect_tx(out_full + st, (uint32_t)__popc(have) * (uint32_t)a.row_bytes);
              else mbar_arrive(out_full + st);
            }
            __syncwarp();     // expect_tx is armed before any copy can complete_tx
            if (optr)
              bulk_g2s(out_rows + ((size_t)st * a.ko + k) * a.row_bytes, optr, (uint32_t)a.row_bytes,
                       out_full + st);
          }
          __syncwarp();       // batch scratch is re-written next round
        }
      }
      G0 += nv;
    }
    if (warp == 1) {
      // out of work: one end marker per consumer warp (in stream order)
      for (int k = 0; k < NW; ++k) {
        const int64_t ii = G0 + k;
        const int cw = (int)(ii % NW);
        const int64_t n = ii / NW;
        const int st = cw * 2 + (int)(n & 1);
        if (lane == 0) {
          mbar_wait(out_empty + st, (uint32_t)((n >> 1) & 1) ^ 1u);
          out_meta[st].active = -1;
          mbar_arrive(out_full + st);
        }
      }
    }
    return;
  }

  // ================================ CONSUMERS =========================================
  const int cw = warp - 2;
  if (cw >= NW) return;
  const int nvec = a.dim >> 2;
  unsigned char* my_din = din_rows + (size_t)cw * kDinSlots * a.row_bytes;
  float loss_acc = 0.f;
  unsigned long long pairs_acc = 0ull;
  uint32_t din_count = 0;
  RowAct act;
  {
    const int full4 = TAIL ? NF4 * 32 : nvec;            // float4s covered by the float4 chunks
#pragma unroll
    for (int j = 0; j < 4; ++j) act.v[j] = j < NF4 && (lane + 32 * j) < full4;
    act.tail_off = NF4 * 512;
    act.t = TAIL && lane < 2 * (nvec - NF4 * 32);
  }

  int64_t n = 0;
  for (int64_t i = cw;; i += NW, ++n) {
    const int st = cw * 2 + (int)(n & 1);
    mbar_wait(out_full + st, (uint32_t)((n >> 1) & 1));
    // the previous position's stage: its reductions were committed at the end of the last iteration;
    // by now the engine has read the rows, so the stage can go back to the producer
    if (n > 0 && lane < kRelLanes) {
      bulk_wait_read<0>();
      mbar_arrive(out_empty + (st ^ 1));
    }
    const OutMeta* meta = out_meta + st;
    const int active = meta->active;
    if (active < 0) break;
    if (active) {
      const int64_t p = meta->p;
      unsigned char* rows = out_rows + (size_t)st * a.ko * a.row_bytes;
      float* my_ptr = (lane < KO) ? meta->ptr[lane] : nullptr;
      const float my_osc = (lane < KO) ? meta->scale[lane] : 1.f;
      const uint32_t used = __ballot_sync(0xffffffffu, my_ptr != nullptr);
      // ---- window of this centre (ParseSentence) ---------------------------------------------
      const uint64_t prng = hash64(a.seed ^ (uint64_t)(p + 1) * 0x9E3779B97F4A7C15ull);
      const int off = (int)((prng >> 16) % (uint64_t)W);
      const int hw = W - off;                         // effective half window, 1..W
      // lane t <-> offset d = t - hw (t in [0, 2hw]); ring position j = i - W + d
      float* cptr = nullptr;
      float csc = 1.f;
      int cslot = 0;
      if (lane <= 2 * hw && lane != hw) {
        const int64_t j = i - W + (lane - hw);
        cslot = (int)(j % R);
        mbar_wait(in_full + cslot, (uint32_t)((j / R) & 1));
        cptr = in_ptr[cslot];
        csc = in_scale[cslot];
      }
      const uint32_t isnull = __ballot_sync(0xffffffffu, lane <= 2 * hw && lane != hw && cptr == nullptr);
      // contexts stop at the first break on either side of the centre
      uint32_t ctx = 0;
      {
        const uint32_t left_null = isnull & ((1u << hw) - 1u);            // lanes 0..hw-1
        const int lo = left_null ? (32 - __clz(left_null)) : 0;           // first valid lane on the left
        const uint32_t right_null = isnull >> (hw + 1);                   // lanes hw+1..
        const int hi = right_null ? (hw + __ffs(right_null) - 1) : 2 * hw;   // last valid lane on the right
        const uint32_t span = (hi >= 31 ? 0xffffffffu : ((1u << (hi + 1)) - 1u)) & ~((1u << lo) - 1u);
        ctx = span & ~(1u << hw);
      }
      if (ctx) {
        // ---- phase 1: output rows in registers, per context dots + input-row delta -------------
        RowT O[KO];
#pragma unroll
        for (int k = 0; k < KO; ++k) {
          row_zero(O[k]);
          if ((used >> k) & 1u) row_load(O[k], rows + (size_t)k * a.row_bytes, lane, act);
        }
        float gs[KO];                                   // lane t keeps the error terms of context t
#pragma unroll
        for (int k = 0; k < KO; ++k) gs[k] = 0.f;
        const bool want_loss = a.loss_sum != nullptr && (p & 7) == 0;
        for (uint32_t m = ctx; m; m &= m - 1) {
          const int t = __ffs(m) - 1;
          const int slot = __shfl_sync(0xffffffffu, cslot, t);
          float* gptr = reinterpret_cast<float*>(__shfl_sync(0xffffffffu, (unsigned long long)cptr, t));
          const float isc = __shfl_sync(0xffffffffu, csc, t);
          RowT h;
          row_load(h, in_rows + (size_t)slot * a.row_bytes, lane, act);
          // 1+K dots; the 8 partial sums are reduced together: three exchange steps halve the number of
          // live values (lanes keep the half they will own), two more finish -> 9 shuffles instead of 5 per dot
          float f8[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) f8[k] = k < KO ? row_dot(h, O[k]) : 0.f;
          {
            const bool up16 = lane & 16, up8 = lane & 8, up4 = lane & 4;
            float f4[4], f2[2], f1;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float send = up16 ? f8[k] : f8[k + 4];
              const float keep = up16 ? f8[k + 4] : f8[k];
              f4[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
            }
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              const float send = up8 ? f4[k] : f4[k + 2];
              const float keep = up8 ? f4[k + 2] : f4[k];
              f2[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
            }
            {
              const float send = up4 ? f2[0] : f2[1];
              const float keep = up4 ? f2[1] : f2[0];
              f1 = keep + __shfl_xor_sync(0xffffffffu, send, 4);
            }
            f1 += __shfl_xor_sync(0xffffffffu, f1, 2);
            f1 += __shfl_xor_sync(0xffffffffu, f1, 1);
            // lane L now holds the full dot of output row k(L) = 4*[L&16] + 2*[L&8] + [L&4]
            const int myk = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
            const bool u = (used >> myk) & 1u;
            const float gmine = u ? ((myk == 0 ? 1.f : 0.f) - sigm_fast(f1)) * a.lr : 0.f;
            if (want_loss && u && (lane & 3) == 0) loss_acc += 8.f * softplus_neg(myk == 0 ? f1 : -f1);
            // lane holding row k: (k&4 ? 16 : 0) + (k&2 ? 8 : 0) + (k&1 ? 4 : 0)
#pragma unroll
            for (int k = 0; k < 8; ++k)
              f8[k] = __shfl_sync(0xffffffffu, gmine, ((k >> 2) & 1) * 16 + ((k >> 1) & 1) * 8 + (k & 1) * 4);
          }
#pragma unroll
          for (int k = 0; k < KO; ++k)
            if (lane == t) gs[k] = f8[k];
          // input-row delta of this context -> staging ring -> one bulk reduction
          unsigned char* dst = my_din + (size_t)(din_count % kDinSlots) * a.row_bytes;
          ++din_count;
          if (lane == 0) bulk_wait_read<kDinSlots - 1>();   // the slot's previous reduction has read it
          __syncwarp();
          RowT e;
          row_zero(e);
#pragma unroll
          for (int k = 0; k < KO; ++k) row_axpy(e, f8[k] * isc, O[k]);
          row_store(dst, e, lane, act);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            bulk_reduce_add_s2g(gptr, dst, (uint32_t)a.row_bytes);
            bulk_commit();
          }
          ++pairs_acc;
        }
        // ---- phase 2: output-row deltas sum_c g_ck * in_c accumulate in registers ----------------
#pragma unroll
        for (int k = 0; k < KO; ++k) row_zero(O[k]);
        for (uint32_t m = ctx; m; m &= m - 1) {
          const int t = __ffs(m) - 1;
          const int slot = __shfl_sync(0xffffffffu, cslot, t);
          RowT h;
          row_load(h, in_rows + (size_t)slot * a.row_bytes, lane, act);
#pragma unroll
          for (int k = 0; k < KO; ++k) {
            const float gk = __shfl_sync(0xffffffffu, gs[k], t) * __shfl_sync(0xffffffffu, my_osc, k);
            row_axpy(O[k], gk, h);
          }
        }
#pragma unroll
        for (int k = 0; k < KO; ++k)
          if ((used >> k) & 1u) row_store(rows + (size_t)k * a.row_bytes, O[k], lane, act);
        fence_proxy_async();
        __syncwarp();
        if (lane < KO && my_ptr != nullptr) {
          bulk_reduce_add_s2g(my_ptr, rows + (size_t)lane * a.row_bytes, (uint32_t)a.row_bytes);
          bulk_commit();
        }
      }
    }
    // ---- release the ring slots whose window this centre closes: j in [i-2W, i] -----------------
    __syncwarp();
    if (lane <= 2 * W) {
      const int64_t j = i - 2 * W + lane;
      if (j >= 0) mbar_arrive(in_empty + (int)(j % R));
    }
  }
  if (lane < kRelLanes) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // adds performed
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) loss_acc += __shfl_xor_sync(0xffffffffu, loss_acc, o);   // lanes hold different rows
  if (lane == 0) {
    if (a.loss_sum && loss_acc != 0.f) atomicAdd(a.loss_sum, loss_acc);
    if (a.pair_count && pairs_acc) atomicAdd(a.pair_count, pairs_acc);
  }
}

}  // namespace

// Centre positions the window-batched kernel keeps in flight on this device (every consumer warp owns
// a double-buffered stage): the staleness window the hot-row step cap is computed from.
extern "C" int mvb_sgns_win_inflight(int dim, int negative, int window, int max_ctas) {
  int dev = 0, max_smem = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  int nw = kNW;
  for (; nw >= 2; --nw)
    if (make_layout(2 * nw + 2 * window + 2, nw, 1 + negative, dim * 4).total <= max_smem) break;
  if (nw < 2) return 0;
  int blocks = mvb_num_sms();
  if (max_ctas > 0 && max_ctas < blocks) blocks = max_ctas;
  return blocks * nw * 2;
}

// Window-batched K7. Returns -20/-21 for configurations it does not cover, -22 when the rows
// do not fit the shared-memory rings (the caller falls back to the pair-at-a-time kernels).
extern "C" int mvb_sgns_train_win(const MvbSgns* h, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (h->n_tokens <= 0) return 0;
  if (h->cbow || h->hs || h->use_adagrad || h->negative < 1 || h->negative > kKO - 1) return -20;
  if (h->dim % 4 || h->ld % 4 || h->dim > 512 || h->window < 1 || h->window > 15) return -21;
  if (h->nservers <= 1 &&
      ((reinterpret_cast<uintptr_t>(h->w_in) & 15) || (reinterpret_cast<uintptr_t>(h->w_out) & 15))) return -21;
  WinDev a{};
  a.tokens = h->tokens; a.n_tokens = h->n_tokens; a.w_in = h->w_in; a.w_out = h->w_out;
  a.dim = h->dim; a.ld = h->ld; a.window = h->window; a.negative = h->negative; a.lr = h->lr;
  a.alias_prob = h->alias_prob; a.alias_idx = h->alias_idx; a.vocab = h->vocab;
  a.neg_pool = h->neg_pool; a.neg_pool_size = h->neg_pool_size; a.neg_pool_size_ptr = h->neg_pool_size_ptr;
  a.map_in = h->map_in;
  a.map_out = h->map_out; a.seed = h->seed; a.loss_sum = h->loss_sum; a.pair_count = h->pair_count;
  a.scale_in = h->scale_in; a.scale_out = h->scale_out;
  a.S = 0;
  if (h->nservers > 1) {
    if (h->map_in || h->map_out || h->rows_per_server <= 0) return -21;
    a.S = h->nservers;
    a.rps = h->rows_per_server;
    for (int s = 0; s < MVB_MAX_RANKS; ++s) {
      a.in_peer[s] = s < h->nservers ? (float*)h->w_in_peers[s] : nullptr;
      a.out_peer[s] = s < h->nservers ? (float*)h->w_out_peers[s] : nullptr;
      if (s < h->nservers && (!a.in_peer[s] || !a.out_peer[s])) return -21;
    }
  }
  a.row_bytes = h->dim * 4;
  a.ko = 1 + h->negative;
  int dev = 0, max_smem = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  // as many consumer warps as the rings allow (kNW for dim <= 300 with 5 negatives)
  int nw = kNW;
  Layout L{};
  for (; nw >= 2; --nw) {
    a.ring = 2 * nw + 2 * h->window + 2;
    L = make_layout(a.ring, nw, a.ko, a.row_bytes);
    if (L.total <= max_smem) break;
  }
  if (nw < 2) return -22;
  if (const char* e = getenv("MVB_WIN_NW")) {
    const int v = atoi(e);
    if (v >= 2 && v < nw) {
      nw = v;
      a.ring = 2 * nw + 2 * h->window + 2;
      L = make_layout(a.ring, nw, a.ko, a.row_bytes);
    }
  }
  a.nw = nw;
  int blocks = mvb_num_sms();
  if (h->max_ctas > 0 && h->max_ctas < blocks) blocks = h->max_ctas;
  // tiny inputs use fewer CTAs (>= 64 positions each)
  const int64_t by_len = (h->n_tokens + 63) / 64;
  if ((int64_t)blocks > by_len) blocks = (int)by_len;
  // One chunk per CTA by default (a static partition: every chunk boundary drains and refills the rings, and
  // the measured imbalance between CTAs is smaller than that: 1M tokens on 148 SMs, chunks of 256 / 512 /
  // 1024 tokens handed out by an atomic ticket ran at 103 / 106 / 107 M words/s against 110 with one chunk per
  // CTA).  MVB_WIN_CHUNK=n deals n-token chunks round-robin instead.
  a.chunk = (int)((h->n_tokens + blocks - 1) / blocks);
  if (const char* e = getenv("MVB_WIN_CHUNK")) {
    const int v = atoi(e);
    if (v >= 32 && v <= (1 << 20)) a.chunk = v;
  }
  const int64_t n_chunks = (h->n_tokens + a.chunk - 1) / a.chunk;
  if ((int64_t)blocks > n_chunks) blocks = (int)n_chunks;
  // register layout of a row: full float4 chunks (+ a partial one) or a float2 tail (Row<NF4, TAIL>)
  const int nvec = h->dim / 4, fc = nvec / 32, rem4 = nvec % 32;
  int nf4 = fc, tail = 0;
  if (rem4 > 16) nf4 = fc + 1;
  else if (rem4 > 0) tail = 1;
  const size_t smem = (size_t)L.total;
#define MVB_LAUNCH_WIN(NF, TL, K)                                                                    \
  do {                                                                                               \
    MVB_CUDA_CHECK(cudaFuncSetAttribute(sgns_win_kernel<NF, TL, K>,                                  \
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));    \
    sgns_win_kernel<NF, TL, K><<<blocks, kThreads, smem, st>>>(a);                                   \
  } while (0)
#define MVB_DISPATCH_WIN(K)                                                  \
  do {                                                                       \
    switch (nf4 * 2 + tail) {                                                \
      case 1: MVB_LAUNCH_WIN(0, true, K); break;                             \
      case 2: MVB_LAUNCH_WIN(1, false, K); break;                            \
      case 3: MVB_LAUNCH_WIN(1, true, K); break;                             \
      case 4: MVB_LAUNCH_WIN(2, false, K); break;                            \
      case 5: MVB_LAUNCH_WIN(2, true, K); break;                             \
      case 6: MVB_LAUNCH_WIN(3, false, K); break;                            \
      case 7: MVB_LAUNCH_WIN(3, true, K); break;                             \
      case 8: MVB_LAUNCH_WIN(4, false, K); break;                            \
      default: return -21;                                                   \
    }                                                                        \
  } while (0)
  if (a.ko <= 6) MVB_DISPATCH_WIN(6);
  else MVB_DISPATCH_WIN(8);
#undef MVB_DISPATCH_WIN
#undef MVB_LAUNCH_WIN
  MVB_CUDA_CHECK(cudaGetLastError());
  return 0;
}
