// Matrix<T> with sparse delta-pull and optional wire compression
// (see include/multiverso/table/matrix.h and table/sparse_matrix_table.h).
#include <algorithm>
#include "multiverso/table/matrix.h"
#include "multiverso/multiverso.h"
#include "multiverso/table/sparse_matrix_table.h"
#include "multiverso/util/configure.h"
#include "multiverso/util/log.h"
#include "multiverso/util/parallel_for.h"
#include "multiverso/util/quantization_util.h"

namespace multiverso {

MV_DECLARE_int(omp_threads);

namespace {
const integer_t kWholeTable = -1;
inline bool IsWhole(const Blob& keys) { return keys.size<integer_t>() == 1 && keys.As<integer_t>(0) == kWholeTable; }
}  // namespace

// ======================================= worker =========================================
template <typename T>
MatrixWorker<T>::MatrixWorker(integer_t num_row, integer_t num_col, bool is_sparse, bool compress)
    : MatrixWorkerTable<T>(num_row, num_col), is_sparse_(is_sparse), compress_(compress) {}

// Whole-table Add on a sparse table: ship only the rows that are not all-zero
// (src/table/matrix.cpp:146-182).
template <typename T>
int MatrixWorker<T>::SubmitWholeAdd(T* data, size_t size, const AddOption* opt) {
  if (!is_sparse_) return MatrixWorkerTable<T>::SubmitWholeAdd(data, size, opt);
  const integer_t R = this->num_row_, C = this->num_col_;
  std::vector<integer_t> rows;
  for (integer_t r = 0; r < R; ++r) {
    const T* row = data + r * C;
    bool nz = false;
    for (integer_t c = 0; c < C && !nz; ++c) nz = row[c] != T(0);
    if (nz) rows.push_back(r);
  }
  {
    // Every server must SEE this whole-table Add, also the ones none of the non-zero rows belongs to: under
    // -sync the SyncServer counts one Add per worker per step (src/server.cpp:141-163), and a worker that
    // skips a server would leave the other workers' Gets parked there until FinishTrain.  Like the
    // reference's dummy zero row (src/table/matrix.cpp:166-170), servers without a non-zero row get one
    // all-zero placeholder row (their first row: it IS all zero in `data`, adding it changes nothing).
    std::vector<char> seen(static_cast<size_t>(this->part_.num_servers), 0);
    for (integer_t r : rows) seen[static_cast<size_t>(this->part_.ServerOf(r))] = 1;
    bool added = false;
    for (int s = 0; s < this->part_.num_servers; ++s)
      if (!seen[static_cast<size_t>(s)]) { rows.push_back(this->part_.row_begin[s]); added = true; }
    if (added) std::sort(rows.begin(), rows.end());
  }
  Blob vals(rows.size() * C * sizeof(T));
  ParallelFor(static_cast<int64_t>(rows.size()), rows.size() >= 2048 ? std::max(1, MV_CONFIG(omp_threads)) : 1,
              [&](int64_t lo, int64_t hi) {
                for (int64_t i = lo; i < hi; ++i)
                  std::memcpy(vals.data() + i * C * sizeof(T), data + rows[i] * C, C * sizeof(T));
              });
  return WorkerTable::AddAsync(Blob(rows.data(), rows.size() * sizeof(integer_t)), std::move(vals), opt);
}

template <typename T>
void MatrixWorker<T>::FilterOutgoing(std::vector<Blob>* blobs) {
  if (!compress_) return;
  // SparseFilter(clip = 0, skip_option_blob): (index,value) pairs when < half is non-zero
  const bool has_opt = blobs->size() == 3;
  SparseFilter<T, int32_t> filter(0.0, has_opt);
  std::vector<Blob> out;
  filter.FilterIn(*blobs, &out);
  *blobs = std::move(out);
}

// ======================================= server =========================================
template <typename T>
MatrixServer<T>::MatrixServer(integer_t num_row, integer_t num_col, bool is_sparse, bool is_pipeline,
                              bool compress)
    : MatrixServerTable<T>(num_row, num_col), is_sparse_(is_sparse), compress_(compress) {
  slots_ = MV_NumWorkers() * (is_pipeline ? 2 : 1);
  if (is_sparse_)
    stale_.assign(static_cast<size_t>(slots_), std::vector<unsigned char>(static_cast<size_t>(this->my_num_row_), 1));
}

template <typename T>
void MatrixServer<T>::MarkStale(const integer_t* rows, size_t n, bool all) {
  for (auto& per : stale_) {
    if (all) std::fill(per.begin(), per.end(), 1);
    else for (size_t i = 0; i < n; ++i) per[static_cast<size_t>(rows[i] - this->row_offset_)] = 1;
  }
}

template <typename T>
void MatrixServer<T>::ProcessAdd(const std::vector<Blob>& data) {
  std::vector<Blob> plain;
  const std::vector<Blob>* in = &data;
  if (compress_) {
    SparseFilter<T, int32_t> filter(0.0, true);
    filter.FilterOut(data, &plain);
    in = &plain;
  }
  MatrixServerTable<T>::ProcessAdd(*in);
  if (is_sparse_) {
    const Blob& keys = (*in)[0];
    if (IsWhole(keys)) MarkStale(nullptr, 0, true);
    else MarkStale(&keys.As<integer_t>(0), keys.size<integer_t>(), false);
  }
}

template <typename T>
void MatrixServer<T>::ProcessGet(const std::vector<Blob>& data, std::vector<Blob>* result) {
  if (!is_sparse_) {
    MatrixServerTable<T>::ProcessGet(data, result);
    return;
  }
  const Blob& keys = data[0];
  int worker = -1;
  if (data.size() >= 2 && data[1].size() >= sizeof(int)) worker = GetOption(data[1].data(), data[1].size()).worker_id();
  if (!IsWhole(keys)) {
    // explicit rows are always served; they become up to date for that worker
    MatrixServerTable<T>::ProcessGet(data, result);
    if (worker >= 0 && worker < slots_)
      for (size_t i = 0; i < keys.size<integer_t>(); ++i)
        stale_[worker][static_cast<size_t>(keys.As<integer_t>(i) - this->row_offset_)] = 0;
    return;
  }
  if (worker < 0 || worker >= slots_) {   // worker_id == -1: return everything
    MatrixServerTable<T>::ProcessGet(data, result);
    return;
  }
  // collect this worker's stale rows (and mark them fresh), then gather them in parallel
  auto& mine = stale_[worker];
  size_t n_stale = 0;
  for (unsigned char f : mine) n_stale += f;
  Blob ids(n_stale * sizeof(integer_t));
  size_t k = 0;
  for (integer_t r = 0; r < this->my_num_row_ && k < n_stale; ++r)
    if (mine[static_cast<size_t>(r)]) {
      ids.As<integer_t>(k++) = r + this->row_offset_;
      mine[static_cast<size_t>(r)] = 0;
    }
  const size_t row_bytes = static_cast<size_t>(this->num_col_) * sizeof(T);
  Blob vals(n_stale * row_bytes);
  const char* base = reinterpret_cast<const char*>(this->storage_.data());
  ParallelFor(static_cast<int64_t>(n_stale), n_stale >= 2048 ? std::max(1, MV_CONFIG(omp_threads)) : 1,
              [&](int64_t lo, int64_t hi) {
                for (int64_t i = lo; i < hi; ++i)
                  std::memcpy(vals.data() + i * row_bytes,
                              base + static_cast<size_t>(ids.As<integer_t>(i) - this->row_offset_) * row_bytes, row_bytes);
              });
  // an explicit (possibly empty) row list -- no row-0 placeholder
  result->push_back(std::move(ids));
  result->push_back(std::move(vals));
  result->emplace_back(&this->server_id_, sizeof(int));
}

template class MatrixWorker<float>;
template class MatrixWorker<double>;
template class MatrixWorker<int>;
template class MatrixServer<float>;
template class MatrixServer<double>;
template class MatrixServer<int>;
template class SparseMatrixWorkerTable<float>;
template class SparseMatrixWorkerTable<double>;
template class SparseMatrixWorkerTable<int>;
template class SparseMatrixServerTable<float>;
template class SparseMatrixServerTable<double>;
template class SparseMatrixServerTable<int>;

}  // namespace multiverso
