// SparseMatrixTable<T>: the older sparse variant = stale-row delta-pull plus SparseFilter
// wire compression of every outgoing Add partition (counterpart of
// include/multiverso/table/sparse_matrix_table.h:14-71, src/table/sparse_matrix_table.cpp).
#ifndef MULTIVERSO_TABLE_SPARSE_MATRIX_TABLE_H_
#define MULTIVERSO_TABLE_SPARSE_MATRIX_TABLE_H_
#include "multiverso/table/matrix.h"

namespace multiverso {

template <typename T> class SparseMatrixWorkerTable;
template <typename T> class SparseMatrixServerTable;

template <typename T>
struct SparseMatrixTableOption {
  SparseMatrixTableOption(integer_t r, integer_t c, bool pipeline = false)
      : num_row(r), num_col(c), using_pipeline(pipeline) {}
  integer_t num_row, num_col;
  bool using_pipeline;
  DEFINE_TABLE_TYPE(T, SparseMatrixWorkerTable, SparseMatrixServerTable);
};

template <typename T>
class SparseMatrixWorkerTable : public MatrixWorker<T> {
 public:
  SparseMatrixWorkerTable(integer_t num_row, integer_t num_col)
      : MatrixWorker<T>(num_row, num_col, true, true) {}
  explicit SparseMatrixWorkerTable(const SparseMatrixTableOption<T>& o)
      : SparseMatrixWorkerTable(o.num_row, o.num_col) {}
};

template <typename T>
class SparseMatrixServerTable : public MatrixServer<T> {
 public:
  SparseMatrixServerTable(integer_t num_row, integer_t num_col, bool using_pipeline)
      : MatrixServer<T>(num_row, num_col, true, using_pipeline, true) {}
  explicit SparseMatrixServerTable(const SparseMatrixTableOption<T>& o)
      : SparseMatrixServerTable(o.num_row, o.num_col, o.using_pipeline) {}
};

}  // namespace multiverso
#endif
