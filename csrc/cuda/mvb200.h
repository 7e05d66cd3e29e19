// multiverso-b200 :: C ABI of the sm_100a data-plane library (libmvb200.so).
//
// Every entry point takes raw device pointers + a cudaStream_t (as void*) so the
// Python layer can drive it with torch tensors (tensor.data_ptr(), current stream)
// through ctypes, and a C++ host program can link it directly. All functions
// return 0 on success or a cudaError_t / negative library error; the message is
// available from mvb_last_error().
#pragma once
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MVB_MAX_RANKS 8
#define MVB_PAD_CHANNELS 64            /* signal-pad channels per rank           */
#define MVB_PAD_WORDS (MVB_PAD_CHANNELS * MVB_MAX_RANKS)
#define MVB_EPOCH_FIN (1ull << 62)      /* ready-slot value of a worker that finished training */

/* element types */
enum { MVB_F32 = 0, MVB_F64 = 1, MVB_I32 = 2, MVB_I64 = 3, MVB_I8 = 4 };
/* updaters (reference: -updater_type default|sgd|momentum_sgd|adagrad|dcasgd|dcasgda) */
enum {
  MVB_UPD_DEFAULT = 0,
  MVB_UPD_SGD = 1,
  MVB_UPD_MOMENTUM = 2,
  MVB_UPD_ADAGRAD = 3,
  MVB_UPD_DCASGD = 4,
  MVB_UPD_DCASGDA = 5
};

/* Same 20-byte layout as the reference AddOption (updater.h:13-69):
 * {worker_id, momentum, learning_rate, rho, lambda}. */
typedef struct MvbAddOpt {
  int worker_id;
  float momentum;
  float lr;
  float rho;
  float lambda;
} MvbAddOpt;

const char* mvb_last_error(void);
int mvb_device_count(void);
int mvb_set_device(int dev);
int mvb_device_info(int dev, int* sms, int* cc_major, int* cc_minor, int64_t* total_mem);

/* ---- symmetric memory (cudaMalloc + cudaIpc) ------------------------------ */
int mvb_symm_alloc(int64_t bytes, void** out_ptr);
int mvb_symm_free(void* ptr);
int mvb_ipc_get_handle(void* ptr, void* handle64 /* 64 bytes */);
int mvb_ipc_open_handle(const void* handle64, void** out_ptr);
int mvb_ipc_close_handle(void* ptr);
int mvb_enable_peer_access(int peer_dev);
int mvb_can_access_peer(int dev, int peer, int* out);
int mvb_memset_async(void* ptr, int value, int64_t bytes, void* stream);
int mvb_memcpy_async(void* dst, const void* src, int64_t bytes, void* stream);
int mvb_stream_sync(void* stream);
int mvb_host_alloc_pinned(int64_t bytes, void** out);
int mvb_host_free_pinned(void* p);
/* plain device memory, streams, events: for hosts without their own CUDA runtime (csrc/device_rt) */
int mvb_device_malloc(int64_t bytes, void** out);
int mvb_device_free(void* p);
int mvb_device_sync(void);
int mvb_stream_create(void** out);
int mvb_stream_destroy(void* stream);
int mvb_event_create(void** out, int timing);
int mvb_event_record(void* event, void* stream);
int mvb_event_sync(void* event);
int mvb_event_elapsed_ms(void* start, void* stop, float* ms);
int mvb_event_destroy(void* event);

/* ---- signal pads / device barrier (K11) ------------------------------------
 * pads[r] = rank r's pad (uint64[MVB_PAD_WORDS]); slot(channel, src) on rank r is
 * written by rank `src`. Epochs are monotonically increasing per channel. */
int mvb_signal(void* const* pads, int me, int world, int channel, uint64_t epoch, void* stream);
int mvb_wait(void* const* pads, int me, int world, int channel, uint64_t epoch,
             uint32_t src_mask, int* err_flag, double timeout_s, void* stream);
int mvb_barrier(void* const* pads, int me, int world, int channel, uint64_t epoch,
                int* err_flag, double timeout_s, void* stream);

/* ---- dense Add / Get (K1, K2) ----------------------------------------------
 * mvb_add_dense_fused: owner-side reduce-scatter + updater. The owner pulls its
 * slice [shard_off, shard_off+shard_len) from each of `nworkers` staging buffers
 * (delta_ptrs[w], full-table sized, peer-mapped) and applies the updater once per
 * worker in worker order, in registers, writing the shard (and state) once.
 * If pads != NULL the kernel first publishes (channel ch_ready) that this rank's
 * staging buffer is complete and waits for all workers in worker_mask, and on
 * completion publishes ch_done (consumed by mvb_get_dense).  */
typedef struct MvbDenseAdd {
  int dtype;              /* MVB_F32 | MVB_F64 | MVB_I32 */
  int updater;            /* MVB_UPD_*                   */
  void* shard;            /* local shard                 */
  void* state0;           /* updater state slab(s)       */
  void* state1;
  int64_t shard_len;      /* elements                    */
  int64_t shard_off;      /* element offset in staging   */
  int64_t state_stride;   /* elements between per-worker state slabs */
  int nworkers;
  uint32_t worker_mask;   /* which workers contribute    */
  const void* delta_ptrs[MVB_MAX_RANKS]; /* indexed by worker id */
  const void* delta_multicast;  /* optional NVLS multicast mapping of the staging buffers:
                                   default/sgd fp32 adds then reduce in the switch */
  MvbAddOpt opts[MVB_MAX_RANKS];         /* per worker          */
  /* fused Add -> Get (optional): push the updated shard into every rank's full-table replica */
  void* replica_ptrs[MVB_MAX_RANKS];     /* indexed by rank, NULL = off */
  void* replica_multicast;               /* NVLS multicast view of the replicas or NULL */
  float scale;            /* delta pre-scale (1 = none)  */
  float clip;             /* |delta| clip, 0 = off       */
  /* fused signalling (optional) */
  void* const* pads;      /* host array[world] of pad pointers or NULL */
  int me, world;
  int ch_ready, ch_done;
  uint64_t epoch;
  int worker_rank[MVB_MAX_RANKS]; /* worker id -> rank (for flag slots) */
  int is_worker;          /* this rank publishes ch_ready             */
  int* err_flag;          /* device int, set on watchdog timeout      */
  int* fin_flag;          /* device int, set to 1 when every worker has finished */
  unsigned int* done_counter; /* device uint, zero-initialised         */
  double timeout_s;
  /* per-worker AddOptions through symmetric memory (optional): opt_box[r] = rank r's option box
     (2 x MVB_MAX_RANKS MvbAddOpt, peer mapped).  The calling worker publishes opts[my_worker] in every
     owner's box before its ready flag; owners apply worker w's delta with worker w's published option.
     NULL boxes: opts[] are used as given (single process / identical options).                       */
  void* opt_box[MVB_MAX_RANKS];
  int my_worker;          /* this rank's worker id (-1: not a worker)  */
} MvbDenseAdd;
int mvb_add_dense_fused(const MvbDenseAdd* a, void* stream);

/* Local (single-source) updater apply: the stand-alone K9 kernel used by the
 * NCCL comparator path and by row-wise / whole adds with one worker. */
int mvb_updater_apply(int dtype, int updater, void* data, const void* delta, void* state0,
                      void* state1, int64_t n, const MvbAddOpt* opt, float scale, void* stream);

/* mvb_get_dense: all-gather by pull. shard_ptrs[s] = server s's shard (peer-mapped),
 * shard_offs[s]/shard_lens[s] in elements, out = local full-size buffer. If pads
 * != NULL waits for ch_done >= epoch from every server rank first. */
typedef struct MvbDenseGet {
  int dtype;
  void* out;
  int nservers;
  const void* shard_ptrs[MVB_MAX_RANKS];
  int64_t shard_offs[MVB_MAX_RANKS];
  int64_t shard_lens[MVB_MAX_RANKS];
  void* const* pads;
  int me, world;
  int ch_done;
  uint64_t epoch;
  int server_rank[MVB_MAX_RANKS];
  int* err_flag;
  double timeout_s;
} MvbDenseGet;
int mvb_get_dense(const MvbDenseGet* g, void* stream);

/* One-sided async push: red.add of (sign * delta[shard range]) into every owner's
 * shard through the peer mapping. Stateless updaters only (default: +, sgd: -). */
int mvb_push_dense_red(int dtype, const void* delta, int nservers, void* const* shard_ptrs,
                       const int64_t* shard_offs, const int64_t* shard_lens, float sign,
                       void* stream);

/* One-sided async push for STATEFUL updaters: remote read-modify-write of the owners' shard +
 * state slabs (state must be peer-mapped). Per-worker state is exact; data is Hogwild. */
int mvb_push_dense_stateful(int dtype, int updater, const void* delta, int nservers,
                            void* const* shard_ptrs, void* const* state0_ptrs, void* const* state1_ptrs,
                            const int64_t* shard_offs, const int64_t* shard_lens,
                            const int64_t* state_strides, const MvbAddOpt* opt, int me, void* stream);

/* ---- row-sparse Get / Add (K3, K4, K10) ------------------------------------ */
typedef struct MvbRowMap {
  int64_t num_row, num_col;
  int nservers;
  int64_t rows_per_server;            /* reference rule: num_row / nservers      */
  void* shard_ptrs[MVB_MAX_RANKS];    /* server s shard base (rows [s*rps, ...)) */
} MvbRowMap;
int mvb_get_rows(int dtype, const MvbRowMap* m, const int64_t* row_ids, int64_t k, void* out,
                 int64_t out_ld, void* stream);
int mvb_add_rows_red(int dtype, const MvbRowMap* m, const int64_t* row_ids, int64_t k,
                     const void* vals, int64_t vals_ld, float sign, void* stream);
/* fused AddDeltaParameter: red.add of (cur - old) * scale per row (fp32, num_col % 4 == 0) */
int mvb_add_rows_delta(const MvbRowMap* m, const int64_t* row_ids, int64_t k, const float* cur,
                       const float* old, int64_t ld, float scale, void* stream);
/* Stateful row add applied by the OWNER on its local shard for rows in its range:
 * ids/vals may be another rank's (peer-mapped) staging. */
int mvb_add_rows_owner(int dtype, int updater, void* shard, void* state0, void* state1,
                       int64_t row_lo, int64_t row_hi, int64_t num_col, int64_t state_stride,
                       const int64_t* row_ids, int64_t k, const void* vals, int64_t vals_ld,
                       const MvbAddOpt* opt, void* stream);
int mvb_row_nonzero_mask(int dtype, const void* data, int64_t rows, int64_t cols, int64_t ld,
                         uint8_t* mask, void* stream);
/* stale-row bitmap (T4/T5 delta-pull): mark rows stale for all workers / query+clear */
int mvb_stale_mark(uint8_t* stale /*[workers][rows]*/, int64_t rows, int nworkers,
                   const int64_t* row_ids, int64_t k /* k<0: all rows */, void* stream);
int mvb_stale_take(uint8_t* stale_w /*[rows] for one worker*/, int64_t rows,
                   const int64_t* row_ids, int64_t k, uint8_t* out_mask, void* stream);
/* ascending positions of the non-zero bytes of a mask; *count on the device */
int mvb_mask_compact(const uint8_t* mask, int64_t n, int64_t* out_ids, int64_t* count, void* stream);

/* ---- KV hash table (K5) ----------------------------------------------------- */
/* open addressing, keys int64 (empty = INT64_MIN), values 8 bytes (f64 or i64) or 4
 * bytes (f32 / i32). Owner of key = mod(key, nservers). */
typedef struct MvbKV {
  int vtype;                       /* MVB_F32|MVB_F64|MVB_I32|MVB_I64 */
  int nservers;
  int64_t capacity;                /* slots per shard (power of two)  */
  void* keys[MVB_MAX_RANKS];       /* int64[capacity] per server      */
  void* vals[MVB_MAX_RANKS];
} MvbKV;
int mvb_kv_init(void* keys, int64_t capacity, void* stream);
/* growth: live keys of a shard; re-insertion of a shard into a larger (initialised) local table */
int mvb_kv_count(const void* keys, int64_t capacity, int64_t* out_count, void* stream);
int mvb_kv_rehash(int vtype, const void* old_keys, const void* old_vals, int64_t old_cap, void* new_keys,
                  void* new_vals, int64_t new_cap, int* err_flag, void* stream);
int mvb_kv_add(const MvbKV* kv, const int64_t* keys, const void* vals, int64_t n, int* err_flag,
               void* stream);
int mvb_kv_get(const MvbKV* kv, const int64_t* keys, void* out_vals, int64_t n, void* stream);
int mvb_kv_dump(int vtype, const void* keys, const void* vals, int64_t capacity,
                int64_t* out_keys, void* out_vals, int64_t* out_count, void* stream);

/* ---- all-reduce (K6) --------------------------------------------------------- */
/* One-shot P2P: every rank reads all peers' symmetric buffers and reduces locally
 * in fixed rank order (deterministic, identical on all ranks). Two-shot: reduce own
 * slice, then write it to every peer. In-kernel signalling on channels ch..ch+1. */
typedef struct MvbAllreduce {
  int dtype;
  int64_t n;
  void* bufs[MVB_MAX_RANKS];    /* symmetric staging, peer-mapped */
  void* out;                    /* local result (may alias bufs[me] for two-shot) */
  void* const* pads;
  int me, world;
  int ch;
  uint64_t epoch;
  int* err_flag;
  unsigned int* done_counter;
  double timeout_s;
} MvbAllreduce;
int mvb_allreduce_oneshot(const MvbAllreduce* a, void* stream);
int mvb_allreduce_twoshot(const MvbAllreduce* a, void* stream);
/* latency path: stage-in from `src`, handshake on channel a->ch (ready only) and one-shot reduction in
 * ONE launch; bufs[r] + slot_off_bytes is the epoch's half of a double-buffered staging area */
int mvb_allreduce_fused(const MvbAllreduce* a, const void* src, int64_t slot_off_bytes, void* stream);
/* NVLS: in-switch reduction through the multicast mapping of the staging buffers (fp32) */
int mvb_allreduce_nvls(const MvbAllreduce* a, void* multicast_ptr, void* stream);

/* ---- WordEmbedding (K7) -------------------------------------------------------- */
typedef struct MvbSgns {
  const int* tokens;         /* word ids, <0 = sentence break            */
  int64_t n_tokens;
  float* w_in;               /* input embeddings  [rows x ld]             */
  float* w_out;              /* output embeddings [rows x ld]             */
  float* g2_in;              /* AdaGrad G^2 (or NULL)                     */
  float* g2_out;
  int dim;
  int64_t ld;
  int window;
  int negative;              /* K (0 when hs)                             */
  int cbow;                  /* 0 skip-gram, 1 cbow                       */
  int hs;                    /* hierarchical softmax                      */
  int use_adagrad;
  float lr;
  /* negative sampling: alias table over vocab (prob,alias), or an explicit pool */
  const float* alias_prob;
  const int* alias_idx;
  int vocab;
  const int* neg_pool;
  int neg_pool_size;
  /* hierarchical softmax: per word path [codelen<=MVB_MAX_CODE] */
  const int* hs_points;      /* [vocab x max_code] inner-node row ids      */
  const int8_t* hs_codes;    /* [vocab x max_code]                         */
  const int* hs_len;         /* [vocab]                                    */
  int hs_max_code;
  /* optional global-id -> local-slot maps (block mode); NULL = identity   */
  const int* map_in;
  const int* map_out;
  uint64_t seed;
  float* loss_sum;           /* optional: += sum of -log sigma(..)         */
  unsigned long long* pair_count;  /* optional: += trained (input,target) samples */
  int variant;               /* kernel variant: 0 auto, 1..5 = negatives held in registers, 10 = TMA pipeline
                                (pair at a time), 20 = window-batched TMA pipeline (one centre per warp) */
  int max_ctas;              /* 0 = one CTA per SM; > 0 caps the persistent grid (leaves SMs to the
                                row pull / push kernels of the pipelined multi-GPU step)            */
  /* optional per-WORD step scales in (0,1] (variant 20): caps the summed step of the Zipf-head rows
     that collect more concurrent stale updates than plain SGD tolerates; NULL = 1 everywhere      */
  const float* scale_in;
  const float* scale_out;
  const int* neg_pool_size_ptr;  /* optional: pool size read on the DEVICE (mvb_we_prepare counts[2]);
                                    variant 20 only, overrides neg_pool_size                          */
  /* direct mode (variant 20, nservers > 1): train IN the row-sharded tables over NVLink -- rows are loaded
     from and reduced into their owners' shards (owner = id / rows_per_server, last server takes the
     remainder) through the peer mappings; no block cache, no maps (Hogwild across GPUs)                 */
  int nservers;
  int64_t rows_per_server;
  void* w_in_peers[MVB_MAX_RANKS];
  void* w_out_peers[MVB_MAX_RANKS];
} MvbSgns;
int mvb_sgns_train(const MvbSgns* a, void* stream);
int mvb_sgns_train_tma(const MvbSgns* a, void* stream);   /* TMA bulk-copy pipeline variant */
int mvb_sgns_train_win(const MvbSgns* a, void* stream);   /* window-batched TMA pipeline      */
int mvb_sgns_win_inflight(int dim, int negative, int window, int max_ctas);  /* centre positions in flight */
/* ---- key-addressed shards for application-defined tables (keys.cu): the device-side extension point ---- */
typedef struct MvbKeyMap {
  int64_t size;                        /* keys 0 .. size-1                                   */
  int64_t per_server;                  /* size / nservers, the last server takes the remainder */
  int nservers;
  int width;                           /* fp32 values per key (SparseTable 1, FTRLTable 2)    */
  void* shard_ptrs[MVB_MAX_RANKS];     /* server s: [keys of s x width] fp32, peer mapped      */
  void* touched_ptrs[MVB_MAX_RANKS];   /* server s: touched bitmap uint32[ceil(keys of s / 32)] */
} MvbKeyMap;
int mvb_keys_add(const MvbKeyMap* m, const int64_t* keys, const float* vals, int64_t n, float sign, void* stream);
int mvb_keys_get(const MvbKeyMap* m, const int64_t* keys, float* out, int64_t n, void* stream);
int mvb_keys_collect(const MvbKeyMap* m, int64_t* out_keys, float* out_vals, int64_t* count, int64_t cap,
                     void* stream);

/* ---- staleness instrumentation: per-shard version counters (symmetric memory), bumped by every Add,
   recorded by every Get; staleness of an Add = other workers' Adds applied since this worker's last Get */
int mvb_stale_on_add(void* const* version_ptrs, int nservers, unsigned long long* last_get,
                     unsigned int* adds_since, unsigned long long* hist, int nbins, void* stream);
int mvb_stale_on_get(void* const* version_ptrs, int nservers, unsigned long long* last_get,
                     unsigned int* adds_since, void* stream);

/* ---- row mailboxes (rowbox.cu): device-side row Add with owner-side apply ------------- */
typedef struct MvbRowBox {
  MvbRowMap map;             /* the table's shards (fp32, num_col % 4 == 0, every rank worker + server)   */
  int me;                    /* my rank == worker id == server id                                     */
  int64_t cap;               /* rows per (owner, source) slot (>= rows of the largest shard)           */
  int64_t slot_bytes;        /* mvb_rowbox_slot_bytes(cap, num_col)                                    */
  void* box[MVB_MAX_RANKS];  /* rank r's mailbox slab: nservers x mvb_rowbox_slots() slots, peer mapped */
  void* ack[MVB_MAX_RANKS];  /* rank r's ack array uint64[nservers], peer mapped                       */
  int* seg;                  /* device scratch int[nservers + 1]                                       */
  unsigned int* done;        /* device scratch uint[2], zero-initialised                               */
  uint64_t* applied;         /* device uint64[nservers], zero-initialised: epochs applied per source   */
  int* go;                   /* device scratch int[nservers]                                           */
  int* err_flag;
  double timeout_s;
} MvbRowBox;
int64_t mvb_rowbox_slot_bytes(int64_t cap, int64_t num_col);
int mvb_rowbox_slots(void);            /* slots per (owner, source) pair (double-buffered: 2) */
int mvb_rowbox_push_delta(const MvbRowBox* b, const int* ids, const int* n_ptr, int64_t n_max,
                          const float* cur, const float* old, int64_t ld, float scale, uint64_t epoch,
                          const MvbAddOpt* opt, int ctas_per_sm, void* stream);
int mvb_rowbox_push_vals(const MvbRowBox* b, const int* ids, const int* n_ptr, int64_t n_max,
                         const float* vals, int64_t ld, float scale, uint64_t epoch, const MvbAddOpt* opt,
                         int ctas_per_sm, void* stream);
int mvb_rowbox_poll(const MvbRowBox* b, int src, int wait, void* stream);
int mvb_rowbox_apply(const MvbRowBox* b, int updater, float* shard, float* st0, float* st1,
                     int64_t state_stride, int64_t row_lo, int src, int ctas_per_sm, void* stream);

/* WordEmbedding block protocol on the device (we_block.cu): PrepareData without a host round trip,
   row pull / delta push on the bulk-copy engine.  All counts live in device memory. */
typedef struct MvbWePrep {
  const int* tokens;         /* block of word ids, <0 = sentence break                          */
  int64_t n_tokens;
  int vocab;
  int negative;              /* K: negative x |input| pool draws (0 = no pool)                 */
  const float* alias_prob;   /* unigram^0.75 alias table                                      */
  const int* alias_idx;
  uint64_t seed;
  uint32_t* bm_in;           /* scratch bitmaps [ceil(vocab/32)]                                */
  uint32_t* bm_out;
  int* chunk_sums;           /* scratch [ceil(vocab/32768)]                                     */
  int* map_in;               /* out: word id -> slot in the input cache (-1 = absent) [vocab]   */
  int* map_out;
  int* ids_in;               /* out: slot -> word id [cap_in]                                   */
  int* ids_out;              /*                      [cap_out]                                  */
  int* neg_pool;             /* out: the block's negative pool [pool_cap]                       */
  int64_t pool_cap;
  int* counts;               /* out (device): [0] = n_in, [1] = n_out, [2] = n_pool             */
  int64_t cap_in, cap_out;
} MvbWePrep;
int mvb_we_prepare(const MvbWePrep* p, void* stream);
int mvb_we_prepare_launches(int negative);      /* kernels + memset/memcpy nodes it enqueues */
int mvb_rows_pull_bulk(const MvbRowMap* m, int esz, const int* ids, const int* n_ptr, int64_t n_max,
                       void* dst_a, void* dst_b, int64_t dst_ld, int max_ctas, void* stream);
int mvb_rows_push_delta_bulk(const MvbRowMap* m, const int* ids, const int* n_ptr, int64_t n_max,
                             const float* cur, const float* old, int64_t ld, float scale, int max_ctas,
                             void* stream);
int mvb_build_alias_table(const double* weights_host, int n, float* prob_host, int* alias_host);

/* ---- LogisticRegression (K8) ---------------------------------------------------- */
/* sparse CSR minibatch: objective 0 linear(squared) 1 sigmoid 2 softmax(dense only) */
typedef struct MvbLrSparse {
  const int64_t* row_ptr;    /* [n+1]                      */
  const int64_t* keys;       /* [nnz] feature ids          */
  const float* vals;         /* [nnz] (NULL => 1.0)        */
  const float* labels;       /* [n]                        */
  const float* sample_w;     /* [n] or NULL                */
  int64_t n;
  int objective;
  const float* w;            /* dense weight vector (pulled table) [dim*out] */
  int64_t dim;               /* input_size incl. bias      */
  int out;                   /* output classes (1 for binary) */
  float* grad;               /* [dim*out] accumulated += g*x / n  (atomic)  */
  float* loss_sum;           /* scalar                      */
  int* correct;              /* scalar: # correct predictions */
  float* pred;               /* optional [n*out]            */
  float* err;                /* scratch [n*out]: (p - y) * weight / n */
  int compute_grad;          /* 0 = predict only            */
} MvbLrSparse;
int mvb_lr_sparse_fwd_bwd(const MvbLrSparse* a, void* stream);
typedef struct MvbLrDense {
  const float* x;            /* [n x dim] row-major (bias column included) */
  const float* labels;       /* [n] class index or target                   */
  int64_t n, dim;
  int out;
  int objective;
  const float* w;            /* [out x dim]                                 */
  float* grad;               /* [out x dim]  = (1/n) sum (p-y) x^T          */
  float* loss_sum;
  int* correct;
  float* pred;               /* optional [n x out]                          */
  float* err;                /* scratch [n x out]                           */
  int compute_grad;
} MvbLrDense;
int mvb_lr_dense_fwd_bwd(const MvbLrDense* a, void* stream);
/* wide dense path (> 64 classes / GEMM-sized minibatches): logits and gradient run on the tcgen05 kernel
   (mvb_get_gemm_fused); this is the softmax / sigmoid / linear epilogue between the two products: loss,
   accuracy, predictions and the error matrix E^T = ((P - Y) / n)^T as [out x n_pad] (zero padded).     */
int mvb_lr_wide_epilogue(const float* logits, const float* labels, int64_t n, int out, int objective,
                         float* err_t, int64_t n_pad, float* pred, float* loss_sum, int* correct, void* stream);
int mvb_transpose_pad_f32(const float* in, int64_t rows, int64_t cols, int64_t ld_in, float* out,
                          int64_t ld_out, void* stream);
int mvb_axpy_f32(float* y, const float* x, int64_t n, float alpha, void* stream);
/* FTRL-proximal: weights from (z,n); gradient emits (dz,dn) (reference objective.cpp:260-336) */
int mvb_ftrl_weights(const float* z, const float* n, float* w, int64_t len, float alpha, float beta,
                     float l1, float l2, void* stream);
int mvb_ftrl_update(float* z, float* n, const float* w, const float* g, int64_t len, float alpha,
                    void* stream);
/* FTRL through the PS: dz = -(g - sigma w), dn = -g^2 with sigma = (sqrt(n + g^2) - sqrt(n)) / alpha */
int mvb_ftrl_delta(const float* n, const float* w, const float* g, float* dz, float* dn, int64_t len,
                   float alpha, void* stream);
/* regulariser add: 1 = L1 sign(w)*c, 2 = L2 w*c */
int mvb_regularize(float* grad, const float* w, int64_t len, int type, float coef, void* stream);

/* ---- fused Get + GEMM (tcgen05 / TMEM / TMA) -------------------------------------- */
/* Y[M x N] = X[M x K] * W[N x K]^T where W's rows are sharded over servers (row-range
 * partition, peer-mapped); optionally also materialises W into w_cache (the Get). */
typedef struct MvbGetGemm {
  const float* x;            /* [M x K] fp32 row-major, local     */
  float* y;                  /* [M x N] fp32                      */
  float* w_cache;            /* reserved (must be NULL)           */
  int64_t M, N, K;
  MvbRowMap wmap;            /* N rows x K cols over servers      */
  int local_server;          /* index of the shard in local HBM, -1 if none */
} MvbGetGemm;
int mvb_get_gemm_fused(const MvbGetGemm* g, void* stream);
int mvb_get_gemm_supported(void);
int mvb_get_gemm_last_config(void);  /* ctas*1000 + grid of the last launch */

#ifdef __cplusplus
}
#endif
