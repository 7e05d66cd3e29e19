// multiverso-b200 :: K2-fused, Worker::Get fused with the first consumer GEMM.
//
//   Y[M x N] = X[M x K] * W[N x K]^T        (fp32 in / fp32 out, TF32 tensor-core math)
//
// W is a MatrixTable whose rows are range-sharded over the servers (peer-mapped HBM). The
// reference pulls the table into a host buffer (MatrixWorkerTable::Get -> ProcessReplyGet
// memcpy, src/table/matrix_table.cpp:58-76,316-341) and only then multiplies on the device.
// Here the pulled row block never lands in local HBM: each CTA owns one 64-row tile of W
// (one server's shard), streams it ONCE from the owner over NVLink with TMA
// (cp.async.bulk.tensor, 128B swizzle) straight into shared memory, and feeds it as the B
// operand of tcgen05.mma (kind::tf32, M=128, N=64, K=8) against up to 8 row tiles of the
// local X -- 8 fp32 accumulators of 64 TMEM columns each = all 512 columns -- so the NVLink
// traffic is |W| regardless of M while X tiles are re-read from local L2.
//
//   warp 0      TMA producer: B ring (one W tile per k-block) + A ring (MT X tiles per k-block)
//   warp 1      TMEM alloc + single-thread tcgen05.mma issue; tcgen05.commit frees smem slots
//   warps 2-5   epilogue: tcgen05.ld (32 lanes x 32 columns) -> registers -> global stores
#include <cuda.h>
#include <cstdio>
#include "mvb_common.cuh"

namespace {

constexpr int BM = 128;          // UMMA_M
constexpr int BN = 64;           // UMMA_N: W rows per CTA
constexpr int BK = 32;           // fp32 elements per k-block = 128 bytes = one swizzle row
constexpr int UK = 8;            // UMMA_K for tf32
constexpr int MT_MAX = 8;        // X row tiles per CTA (8 * 64 = 512 TMEM columns)
constexpr int A_STAGES = 6;      // 16 KB each
constexpr int B_STAGES = 4;      // 8 KB each
constexpr int A_BYTES = BM * BK * 4;
constexpr int B_BYTES = BN * BK * 4;
constexpr int kThreads = 192;

struct GemmDev {
  float* y;
  int64_t M, N, K;
  int64_t ldy;
  int S;
  int64_t row_begin[MVB_MAX_RANKS + 1];   // global row range of server s
  int tile_begin[MVB_MAX_RANKS + 1];      // first n-tile index of server s
  int mt_groups;                          // ceil(ceil(M/128) / MT_MAX)
};

MVB_DEVINL uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
MVB_DEVINL void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
MVB_DEVINL void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
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
constexpr uint32_t kInstrDesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) |
                                ((uint32_t)(BM >> 4) << 24);

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

struct SmemLayout {
  uint64_t a_full[A_STAGES], a_empty[A_STAGES], b_full[B_STAGES], b_empty[B_STAGES], acc_full;
  uint32_t tmem_base;
};

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
  unsigned char* a_tiles = smem;                              // A_STAGES * 16 KB (1024-aligned)
  unsigned char* b_tiles = smem + A_STAGES * A_BYTES;         // B_STAGES * 8 KB
  SmemLayout* sl = reinterpret_cast<SmemLayout*>(b_tiles + B_STAGES * B_BYTES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- which W tile / which group of X row tiles ---------------------------------------
  const int ntile = blockIdx.x;
  int s = 0;
  while (s + 1 < g.S && ntile >= g.tile_begin[s + 1]) ++s;
  const int64_t n_local = (int64_t)(ntile - g.tile_begin[s]) * BN;       // row inside the shard
  const int64_t n_global = g.row_begin[s] + n_local;
  const int64_t n_valid = min((int64_t)BN, g.row_begin[s + 1] - n_global);
  const int tiles_m = (int)((g.M + BM - 1) / BM);
  const int mt0 = blockIdx.y * MT_MAX;
  const int MT = min(MT_MAX, tiles_m - mt0);
  const int num_kb = (int)((g.K + BK - 1) / BK);
  const CUtensorMap* map_w = s == 0 ? &map_w0 : s == 1 ? &map_w1 : s == 2 ? &map_w2 : s == 3 ? &map_w3
                           : s == 4 ? &map_w4 : s == 5 ? &map_w5 : s == 6 ? &map_w6 : &map_w7;

  if (threadIdx.x == 0) {
    for (int i = 0; i < A_STAGES; ++i) { mbar_init(&sl->a_full[i], 1); mbar_init(&sl->a_empty[i], 1); }
    for (int i = 0; i < B_STAGES; ++i) { mbar_init(&sl->b_full[i], 1); mbar_init(&sl->b_empty[i], 1); }
    mbar_init(&sl->acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    // all 512 columns: MT accumulators of BN columns (power-of-two allocation)
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
      int ia = 0, ib = 0;
      uint32_t pa = 0, pb = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&sl->b_empty[ib], pb ^ 1u);
        mbar_expect_tx(&sl->b_full[ib], B_BYTES);
        tma_load_2d(b_tiles + ib * B_BYTES, map_w, kb * BK, (int)n_local, &sl->b_full[ib]);   // peer HBM
        if (++ib == B_STAGES) { ib = 0; pb ^= 1u; }
        for (int mt = 0; mt < MT; ++mt) {
          mbar_wait(&sl->a_empty[ia], pa ^ 1u);
          mbar_expect_tx(&sl->a_full[ia], A_BYTES);
          tma_load_2d(a_tiles + ia * A_BYTES, &map_x, kb * BK, (mt0 + mt) * BM, &sl->a_full[ia]);
          if (++ia == A_STAGES) { ia = 0; pa ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ============================== MMA ISSUER =========================================
      int ia = 0, ib = 0;
      uint32_t pa = 0, pb = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&sl->b_full[ib], pb);
        const uint32_t b_addr = smem_u32(b_tiles + ib * B_BYTES);
        for (int mt = 0; mt < MT; ++mt) {
          mbar_wait(&sl->a_full[ia], pa);
          asm volatile("tcgen05.fence::after_thread_sync;");
          const uint32_t a_addr = smem_u32(a_tiles + ia * A_BYTES);
#pragma unroll
          for (int k = 0; k < BK / UK; ++k) {
            // advance 32 bytes inside the 128-byte swizzle row per UMMA_K step
            umma_tf32(tmem_base + (uint32_t)(mt * BN), make_smem_desc(a_addr + k * UK * 4),
                      make_smem_desc(b_addr + k * UK * 4), (kb | k) ? 1u : 0u);
          }
          umma_commit(&sl->a_empty[ia]);            // A slot is free once these MMAs retire
          if (++ia == A_STAGES) { ia = 0; pa ^= 1u; }
        }
        umma_commit(&sl->b_empty[ib]);
        if (++ib == B_STAGES) { ib = 0; pb ^= 1u; }
      }
      umma_commit(&sl->acc_full);                   // every accumulator is final
    }
  } else {
    // ================================ EPILOGUE ===========================================
    mbar_wait(&sl->acc_full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;");
    const int quarter = warp & 3;                   // TMEM lane quarter this warp may access
    for (int mt = 0; mt < MT; ++mt) {
      const int64_t m = (int64_t)(mt0 + mt) * BM + quarter * 32 + lane;
#pragma unroll
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t v[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(mt * BN + c0);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
              "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
              "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
              "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (m < g.M) {
          float* out = g.y + m * g.ldy + n_global + c0;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (c0 + j < n_valid) out[j] = __uint_as_float(v[j]);
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
  }
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
  int rc = make_map(&maps[0], h->x, h->M, h->K, h->K, BM);
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
    tiles += (int)((hi - lo + BN - 1) / BN);
    g.tile_begin[s + 1] = tiles;
    rc = make_map(&maps[1 + s], wm.shard_ptrs[s], hi - lo > 0 ? hi - lo : 1, h->K, h->K, BN);
    if (rc) return rc;
  }
  for (int s = wm.nservers; s < MVB_MAX_RANKS; ++s) maps[1 + s] = maps[1];
  const int tiles_m = (int)((h->M + BM - 1) / BM);
  g.mt_groups = (tiles_m + MT_MAX - 1) / MT_MAX;
  const size_t smem = A_STAGES * A_BYTES + B_STAGES * B_BYTES + sizeof(SmemLayout) + 1024;
  MVB_CUDA_CHECK(cudaFuncSetAttribute(get_gemm_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)tiles, (unsigned)g.mt_groups);
  get_gemm_fused_kernel<<<grid, kThreads, smem, st>>>(maps[0], maps[1], maps[2], maps[3], maps[4], maps[5], maps[6],
                                                      maps[7], maps[8], g);
  MVB_CUDA_CHECK(cudaGetLastError());
  return 0;
}
