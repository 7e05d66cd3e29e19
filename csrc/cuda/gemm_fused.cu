// multiverso-b200 :: K2-fused, Worker::Get fused with the first consumer GEMM.
//
//   Y[M x N] = X[M x K] * W[N x K]^T        (fp32 in / fp32 out, TF32 tensor-core math)
//
// W is a MatrixTable whose rows are range-sharded over the servers (peer-mapped HBM). The
// reference pulls the table into a host buffer (MatrixWorkerTable::Get -> ProcessReplyGet
// memcpy, src/table/matrix_table.cpp:58-76,316-341) and only then multiplies on the device.
// Here the pulled row block never lands in local HBM: a work item is one 128-row tile of W
// (one server's shard) times up to 512 rows of X.  The W tile is streamed from its owner over
// NVLink with TMA (cp.async.bulk.tensor, 128B swizzle) straight into shared memory and is the
// *A* operand of tcgen05.mma (kind::tf32, M=128, N=256, K=8): the accumulator is the transposed
// tile Y^T[w_row, x_row] -- 128 TMEM lanes x 2 x 256 columns = all of TMEM -- so W crosses NVLink
// ceil(M/512) times while X (the B operand, N=256 per instruction: the shape the tensor pipe runs
// at full rate) is re-read from local L2.  Lanes = consecutive Y columns, so the epilogue's
// tcgen05.ld registers store straight to 128-byte coalesced row segments of Y.
//
//   warp 0      TMA producer: W ring (one tile per k-block) + X ring (<=2 chunks per k-block)
//   warp 1      TMEM alloc + single-thread tcgen05.mma issue; tcgen05.commit frees smem slots
//   warps 2-5   epilogue: tcgen05.ld (32 lanes x 32 columns) -> registers -> global stores
// Persistent CTAs (one per SM); the two 256-column accumulators are handed back to the MMA
// issuer one by one, so the next item's MMAs overlap the rest of the epilogue (items with
// M <= 256 alternate accumulators = full double buffering).
#include <cuda.h>
#include <algorithm>
#include <cstdio>
#include "mvb_common.cuh"

namespace {

constexpr int WM = 128;          // UMMA_M: W rows per tile (TMEM lanes)
constexpr int XN = 256;          // UMMA_N: X rows per MMA (TMEM columns per accumulator)
constexpr int XC_MAX = 2;        // accumulators (X chunks) per item: 2 * 256 = 512 TMEM columns
constexpr int BK = 32;           // fp32 elements per k-block = 128 bytes = one swizzle row
constexpr int UK = 8;            // UMMA_K for tf32
constexpr int W_STAGES = 4;      // 16 KB each
constexpr int X_STAGES = 4;      // 32 KB each
constexpr int W_BYTES = WM * BK * 4;
constexpr int X_BYTES = XN * BK * 4;
constexpr int kEpiWarps = 4;
constexpr int kThreads = 64 + 32 * kEpiWarps;

struct GemmDev {
  float* y;
  int64_t M, N, K;
  int64_t ldy;
  int S;
  int64_t row_begin[MVB_MAX_RANKS + 1];   // global row range of server s
  int tile_begin[MVB_MAX_RANKS + 1];      // first W-tile index of server s
  int tiles_n;                            // total W tiles
  int x_groups;                           // ceil(M / 512)
};

MVB_DEVINL uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
MVB_DEVINL void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
MVB_DEVINL void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
MVB_DEVINL void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
MVB_DEVINL void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\nLAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra LAB_DONE;\nbra LAB_WAIT;\nLAB_DONE:\n}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
MVB_DEVINL void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
// K-major operand, 128B swizzle: 8-row groups are 1024 B apart (SBO), LBO unused, version 1.
MVB_DEVINL uint64_t make_smem_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);          // start address  [0,14)
  d |= (uint64_t)0 << 16;                               // leading byte offset (ignored)
  d |= (uint64_t)(1024 >> 4) << 32;                     // stride byte offset [32,46)
  d |= (uint64_t)1 << 46;                               // descriptor version (sm100)
  d |= (uint64_t)2 << 61;                               // SWIZZLE_128B
  return d;
}
// c=F32, a=b=TF32, both K-major, N>>3 at [17,23), M>>4 at [24,29)
constexpr uint32_t kInstrDesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(XN >> 3) << 17) |
                                ((uint32_t)(WM >> 4) << 24);

MVB_DEVINL void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(kInstrDesc), "r"(accumulate) : "memory");
}
MVB_DEVINL void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
               ::"r"(smem_u32(bar)) : "memory");
}
MVB_DEVINL void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}

struct SmemLayout {
  uint64_t w_full[W_STAGES], w_empty[W_STAGES], x_full[X_STAGES], x_empty[X_STAGES];
  uint64_t acc_full[XC_MAX], acc_empty[XC_MAX];
  uint32_t tmem_base;
};

// One work item = (W tile, group of up to 512 X rows).  Items are dealt round-robin to the
// persistent CTAs with the W tile varying fastest, so CTAs running side by side share the same X
// rows in L2 while each streams its own W tile.
struct Item {
  int s, XC;
  int64_t n_local, n_global, n_valid, x0;
};
MVB_DEVINL Item decode_item(const GemmDev& g, int item) {
  Item it;
  const int xg = item / g.tiles_n, ntile = item - xg * g.tiles_n;
  int s = 0;
  while (s + 1 < g.S && ntile >= g.tile_begin[s + 1]) ++s;
  it.s = s;
  it.n_local = (int64_t)(ntile - g.tile_begin[s]) * WM;       // row inside the shard
  it.n_global = g.row_begin[s] + it.n_local;
  it.n_valid = min((int64_t)WM, g.row_begin[s + 1] - it.n_global);
  it.x0 = (int64_t)xg * (XN * XC_MAX);
  it.XC = (int)min((int64_t)XC_MAX, (g.M - it.x0 + XN - 1) / XN);
  return it;
}

__global__ void __launch_bounds__(kThreads, 1)
get_gemm_fused_kernel(const __grid_constant__ CUtensorMap map_x,
                      const __grid_constant__ CUtensorMap map_w0, const __grid_constant__ CUtensorMap map_w1,
                      const __grid_constant__ CUtensorMap map_w2, const __grid_constant__ CUtensorMap map_w3,
                      const __grid_constant__ CUtensorMap map_w4, const __grid_constant__ CUtensorMap map_w5,
                      const __grid_constant__ CUtensorMap map_w6, const __grid_constant__ CUtensorMap map_w7,
                      const __grid_constant__ GemmDev g) {
  extern __shared__ unsigned char smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment: align the dynamic segment by hand
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  unsigned char* x_tiles = smem;                              // X_STAGES * 32 KB (1024-aligned)
  unsigned char* w_tiles = smem + X_STAGES * X_BYTES;         // W_STAGES * 16 KB
  SmemLayout* sl = reinterpret_cast<SmemLayout*>(w_tiles + W_STAGES * W_BYTES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const int num_items = g.tiles_n * g.x_groups;
  const int num_kb = (int)((g.K + BK - 1) / BK);

  if (threadIdx.x == 0) {
    for (int i = 0; i < W_STAGES; ++i) { mbar_init(&sl->w_full[i], 1); mbar_init(&sl->w_empty[i], 1); }
    for (int i = 0; i < X_STAGES; ++i) { mbar_init(&sl->x_full[i], 1); mbar_init(&sl->x_empty[i], 1); }
    for (int i = 0; i < XC_MAX; ++i) { mbar_init(&sl->acc_full[i], 1); mbar_init(&sl->acc_empty[i], kEpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&sl->tmem_base)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem_base = sl->tmem_base;

  if (warp == 0) {
    if (lane == 0) {
      // ============================== TMA PRODUCER ======================================
      int iw = 0, ix = 0;
      uint32_t pw = 0, px = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        const Item it = decode_item(g, item);
        const int s = it.s;
        const CUtensorMap* map_w = s == 0 ? &map_w0 : s == 1 ? &map_w1 : s == 2 ? &map_w2 : s == 3 ? &map_w3
                                 : s == 4 ? &map_w4 : s == 5 ? &map_w5 : s == 6 ? &map_w6 : &map_w7;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&sl->w_empty[iw], pw ^ 1u);
          mbar_expect_tx(&sl->w_full[iw], W_BYTES);
          tma_load_2d(w_tiles + iw * W_BYTES, map_w, kb * BK, (int)it.n_local, &sl->w_full[iw]);   // peer HBM
          if (++iw == W_STAGES) { iw = 0; pw ^= 1u; }
          for (int xc = 0; xc < it.XC; ++xc) {
            mbar_wait(&sl->x_empty[ix], px ^ 1u);
            mbar_expect_tx(&sl->x_full[ix], X_BYTES);
            tma_load_2d(x_tiles + ix * X_BYTES, &map_x, kb * BK, (int)(it.x0 + xc * XN), &sl->x_full[ix]);
            if (++ix == X_STAGES) { ix = 0; px ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ============================== MMA ISSUER =========================================
      int iw = 0, ix = 0;
      uint32_t pw = 0, px = 0;
      uint32_t q = 0;                                 // running accumulator-chunk counter: slot = q & 1
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        const Item it = decode_item(g, item);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&sl->w_full[iw], pw);
          const uint32_t w_addr = smem_u32(w_tiles + iw * W_BYTES);
          for (int xc = 0; xc < it.XC; ++xc) {
            const uint32_t qc = q + (uint32_t)xc, slot = qc & 1u;
            // the slot must have been drained by the epilogue of its previous use
            if (kb == 0 && qc >= 2) mbar_wait(&sl->acc_empty[slot], ((qc >> 1) - 1u) & 1u);
            mbar_wait(&sl->x_full[ix], px);
            asm volatile("tcgen05.fence::after_thread_sync;");
            const uint32_t x_addr = smem_u32(x_tiles + ix * X_BYTES);
#pragma unroll
            for (int k = 0; k < BK / UK; ++k) {
              // advance 32 bytes inside the 128-byte swizzle row per UMMA_K step
              umma_tf32(tmem_base + slot * XN, make_smem_desc(w_addr + k * UK * 4),
                        make_smem_desc(x_addr + k * UK * 4), (kb | k) ? 1u : 0u);
            }
            umma_commit(&sl->x_empty[ix]);            // X slot is free once these MMAs retire
            if (++ix == X_STAGES) { ix = 0; px ^= 1u; }
          }
          umma_commit(&sl->w_empty[iw]);
          if (++iw == W_STAGES) { iw = 0; pw ^= 1u; }
        }
        for (int xc = 0; xc < it.XC; ++xc) umma_commit(&sl->acc_full[(q + (uint32_t)xc) & 1u]);
        q += (uint32_t)it.XC;
      }
    }
  } else {
    // ================================ EPILOGUE ===========================================
    // lanes = consecutive W rows = consecutive Y columns: register j of the warp is one 128-byte
    // segment of Y row (x0 + ... + j).
    const int quarter = warp & 3;                   // TMEM lane quarter this warp may access
    uint32_t q = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      const Item it = decode_item(g, item);
      const int64_t ncol = it.n_local + quarter * 32 + lane;                 // row inside the W tile's shard
      const bool n_ok = quarter * 32 + lane < it.n_valid;
      float* ycol = g.y + it.n_global + quarter * 32 + lane;
      (void)ncol;
      for (int xc = 0; xc < it.XC; ++xc, ++q) {
        const uint32_t slot = q & 1u;
        mbar_wait(&sl->acc_full[slot], (q >> 1) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;");
        const int64_t m0 = it.x0 + (int64_t)xc * XN;
        const int rows = (int)min((int64_t)XN, g.M - m0);
#pragma unroll 1
        for (int cc = 0; cc < XN; cc += 64) {
          if (cc < rows) {
            uint32_t v0[32], v1[32];
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + slot * XN + (uint32_t)cc;
            tmem_ld_32x32(taddr, v0);
            tmem_ld_32x32(taddr + 32, v1);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (cc + 64 >= rows || cc + 64 >= XN) {
              // last TMEM read of this accumulator: hand the columns back before storing
              asm volatile("tcgen05.fence::before_thread_sync;");
              __syncwarp();
              if (lane == 0) mbar_arrive(&sl->acc_empty[slot]);
            }
            if (n_ok) {
              float* out = ycol + (m0 + cc) * g.ldy;
              if (cc + 64 <= rows) {
#pragma unroll
                for (int j = 0; j < 32; ++j) out[(int64_t)j * g.ldy] = __uint_as_float(v0[j]);
#pragma unroll
                for (int j = 0; j < 32; ++j) out[(int64_t)(32 + j) * g.ldy] = __uint_as_float(v1[j]);
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (cc + j < rows) out[(int64_t)j * g.ldy] = __uint_as_float(v0[j]);
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (cc + 32 + j < rows) out[(int64_t)(32 + j) * g.ldy] = __uint_as_float(v1[j]);
              }
            }
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
  }
}

// ---- host side ---------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D fp32 row-major [rows x cols], box = BK columns x box_rows rows, 128B swizzle
int make_map(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return -30;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -31;
}

}  // namespace

extern "C" int mvb_get_gemm_supported(void) { return get_encode() != nullptr ? 1 : 0; }

extern "C" int mvb_get_gemm_fused(const MvbGetGemm* h, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (h->M <= 0 || h->N <= 0 || h->K <= 0) return 0;
  // TMA needs 16-byte aligned row pitches
  if (h->K % 4 != 0) return -32;
  const MvbRowMap& wm = h->wmap;
  if (wm.nservers < 1 || wm.nservers > MVB_MAX_RANKS || wm.num_col != h->K || wm.num_row != h->N) return -33;
  GemmDev g{};
  g.y = h->y; g.M = h->M; g.N = h->N; g.K = h->K; g.ldy = h->N; g.S = wm.nservers;
  CUtensorMap maps[1 + MVB_MAX_RANKS];
  int rc = make_map(&maps[0], h->x, h->M, h->K, h->K, XN);
  if (rc) return rc;
  int tiles = 0;
  for (int s = 0; s < wm.nservers; ++s) {
    int64_t lo = wm.rows_per_server * s;
    int64_t hi = (s == wm.nservers - 1) ? wm.num_row : wm.rows_per_server * (s + 1);
    if (lo > wm.num_row) lo = wm.num_row;
    if (hi > wm.num_row) hi = wm.num_row;
    g.row_begin[s] = lo;
    g.row_begin[s + 1] = hi;
    g.tile_begin[s] = tiles;
    tiles += (int)((hi - lo + WM - 1) / WM);
    g.tile_begin[s + 1] = tiles;
    rc = make_map(&maps[1 + s], wm.shard_ptrs[s], hi - lo > 0 ? hi - lo : 1, h->K, h->K, WM);
    if (rc) return rc;
  }
  for (int s = wm.nservers; s < MVB_MAX_RANKS; ++s) maps[1 + s] = maps[1];
  g.x_groups = (int)((h->M + XN * XC_MAX - 1) / (XN * XC_MAX));
  g.tiles_n = tiles;
  const size_t smem = X_STAGES * X_BYTES + W_STAGES * W_BYTES + sizeof(SmemLayout) + 1024;
  MVB_CUDA_CHECK(cudaFuncSetAttribute(get_gemm_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t items = (int64_t)tiles * g.x_groups;
  dim3 grid((unsigned)std::min<int64_t>(items, mvb_num_sms()));
  get_gemm_fused_kernel<<<grid, kThreads, smem, st>>>(maps[0], maps[1], maps[2], maps[3], maps[4], maps[5], maps[6],
                                                      maps[7], maps[8], g);
  MVB_CUDA_CHECK(cudaGetLastError());
  return 0;
}
