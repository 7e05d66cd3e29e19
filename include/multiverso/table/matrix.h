// Matrix<T> (MatrixWorker / MatrixServer / MatrixOption): the sparse-aware successor of
// MatrixTable (counterpart of include/multiverso/table/matrix.h:14-123, src/table/matrix.cpp).
// With is_sparse: (a) a whole-table Add ships only rows that are not all-zero; (b) the
// server keeps an up-to-date bitmap per (worker[, pipeline slot], row): every Add marks the
// touched rows stale for ALL workers (T4 behaviour, Q13), a whole-table Get returns only the
// rows stale for the requesting worker -- an explicit empty reply when nothing is stale
// (the reference sends row 0 as a placeholder, Q12); GetOption.worker_id == -1 returns all.
#ifndef MULTIVERSO_TABLE_MATRIX_H_
#define MULTIVERSO_TABLE_MATRIX_H_
#include "multiverso/table/matrix_table.h"

namespace multiverso {

template <typename T> class MatrixWorker;
template <typename T> class MatrixServer;

template <typename T>
struct MatrixOption {
  integer_t num_row = 0, num_col = 0;
  bool is_sparse = false;
  bool is_pipeline = false;
  DEFINE_TABLE_TYPE(T, MatrixWorker, MatrixServer);
};

template <typename T>
class MatrixWorker : public MatrixWorkerTable<T> {
 public:
  MatrixWorker(integer_t num_row, integer_t num_col, bool is_sparse = false, bool compress = false);
  explicit MatrixWorker(const MatrixOption<T>& o) : MatrixWorker(o.num_row, o.num_col, o.is_sparse) {}
  bool is_sparse() const { return is_sparse_; }

 protected:
  int SubmitWholeAdd(T* data, size_t size, const AddOption* opt) override;
  void FilterOutgoing(std::vector<Blob>* blobs) override;
  bool is_sparse_, compress_;
};

template <typename T>
class MatrixServer : public MatrixServerTable<T> {
 public:
  MatrixServer(integer_t num_row, integer_t num_col, bool is_sparse = false, bool is_pipeline = false,
               bool compress = false);
  explicit MatrixServer(const MatrixOption<T>& o)
      : MatrixServer(o.num_row, o.num_col, o.is_sparse, o.is_pipeline) {}
  void ProcessAdd(const std::vector<Blob>& data) override;
  void ProcessGet(const std::vector<Blob>& data, std::vector<Blob>* result) override;

 private:
  void MarkStale(const integer_t* rows, size_t n, bool all);
  bool is_sparse_, compress_;
  int slots_;                                      // workers (x2 when pipelined)
  std::vector<std::vector<unsigned char>> stale_;  // [slot][local row]
};

}  // namespace multiverso
#endif
