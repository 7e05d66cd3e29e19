// MatrixTable<T>: dense 2-D, row-addressable table (counterpart of
// include/multiverso/table/matrix_table.h:16-127, src/table/matrix_table.cpp). Rows are
// range-partitioned: num_row/num_servers rows each, the last server takes the remainder;
// with fewer rows than servers the first num_row servers own one row each.
// Worker API, each in sync + async flavour: whole table / one row / row-id vector with
// scattered destination pointers / row-id array with one contiguous buffer.
#ifndef MULTIVERSO_TABLE_MATRIX_TABLE_H_
#define MULTIVERSO_TABLE_MATRIX_TABLE_H_
#include <utility>
#include <vector>
#include "multiverso/table_interface.h"

namespace multiverso {

template <typename T> class MatrixWorkerTable;
template <typename T> class MatrixServerTable;

template <typename T>
struct MatrixTableOption {
  MatrixTableOption(integer_t r, integer_t c) : num_row(r), num_col(c) {}
  MatrixTableOption(integer_t r, integer_t c, T lo, T hi)
      : num_row(r), num_col(c), random_init(true), min_value(lo), max_value(hi) {}
  integer_t num_row, num_col;
  bool random_init = false;
  T min_value = T(), max_value = T();
  DEFINE_TABLE_TYPE(T, MatrixWorkerTable, MatrixServerTable);
};

// Row -> server mapping shared by worker and server halves.
struct RowPartition {
  RowPartition() = default;
  RowPartition(integer_t num_row, int num_servers);
  int ServerOf(integer_t row) const;
  integer_t num_row = 0;
  int num_servers = 1;        // "actual" servers (<= configured)
  integer_t rows_each = 1;
  std::vector<integer_t> row_begin;   // size num_servers + 1
};

template <typename T>
class MatrixWorkerTable : public WorkerTable {
 public:
  MatrixWorkerTable(integer_t num_row, integer_t num_col);
  explicit MatrixWorkerTable(const MatrixTableOption<T>& o) : MatrixWorkerTable(o.num_row, o.num_col) {}

  // ---- Get ----
  void Get(T* data, size_t size, const GetOption* opt = nullptr);
  void Get(integer_t row_id, T* data, size_t size, const GetOption* opt = nullptr);
  void Get(const std::vector<integer_t>& row_ids, const std::vector<T*>& data_vec, size_t size,
           const GetOption* opt = nullptr);
  void Get(T* data, size_t size, integer_t* row_ids, int row_ids_size, const GetOption* opt = nullptr);
  int GetAsync(T* data, size_t size, const GetOption* opt = nullptr);
  int GetAsync(integer_t row_id, T* data, size_t size, const GetOption* opt = nullptr);
  int GetAsync(const std::vector<integer_t>& row_ids, const std::vector<T*>& data_vec, size_t size,
               const GetOption* opt = nullptr);
  int GetAsync(T* data, size_t size, integer_t* row_ids, int row_ids_size, const GetOption* opt = nullptr);
  // ---- Add ----
  void Add(T* data, size_t size, const AddOption* opt = nullptr);
  void Add(integer_t row_id, T* data, size_t size, const AddOption* opt = nullptr);
  void Add(const std::vector<integer_t>& row_ids, const std::vector<T*>& data_vec, size_t size,
           const AddOption* opt = nullptr);
  void Add(T* data, size_t size, integer_t* row_ids, int row_ids_size, const AddOption* opt = nullptr);
  int AddAsync(T* data, size_t size, const AddOption* opt = nullptr);
  int AddAsync(integer_t row_id, T* data, size_t size, const AddOption* opt = nullptr);
  int AddAsync(const std::vector<integer_t>& row_ids, const std::vector<T*>& data_vec, size_t size,
               const AddOption* opt = nullptr);
  int AddAsync(T* data, size_t size, integer_t* row_ids, int row_ids_size, const AddOption* opt = nullptr);

  integer_t num_row() const { return num_row_; }
  integer_t num_col() const { return num_col_; }

  int Partition(const std::vector<Blob>& kv, MsgType partition_type,
                std::unordered_map<int, std::vector<Blob>>* out) override;
  void ProcessReplyGet(std::vector<Blob>& reply_data, int msg_id) override;

 protected:
  void OnRequestDone(int msg_id) override;
  // Hooks for the sparse subclasses.
  virtual void FilterOutgoing(std::vector<Blob>* /*blobs*/) {}
  virtual int SubmitWholeAdd(T* data, size_t size, const AddOption* opt);

  // Destination of one in-flight Get. Row requests keep (row id, destination) pairs sorted by
  // row id -- built with one sort (none when the ids arrive sorted), looked up by binary search --
  // instead of a node-per-row hash map: a WordEmbedding block pulls 10^5..10^6 rows per request.
  struct GetRecord {
    T* whole = nullptr;                              // destination of a whole-table Get
    std::vector<std::pair<integer_t, T*>> rows;      // sorted by row id
    void AddRow(integer_t row, T* dst) { rows.emplace_back(row, dst); }
    void Seal();                                     // sort if needed
    T* Find(integer_t row) const;
  };
  int SubmitGet(GetRecord&& rec, Blob keys, const GetOption* opt);
  integer_t num_row_, num_col_;
  RowPartition part_;
  std::mutex rec_mu_;
  std::unordered_map<int, GetRecord> records_;
};

template <typename T>
class MatrixServerTable : public ServerTable {
 public:
  MatrixServerTable(integer_t num_row, integer_t num_col);
  MatrixServerTable(integer_t num_row, integer_t num_col, T min_value, T max_value);
  explicit MatrixServerTable(const MatrixTableOption<T>& o);
  ~MatrixServerTable() override;
  void ProcessAdd(const std::vector<Blob>& data) override;
  void ProcessGet(const std::vector<Blob>& data, std::vector<Blob>* result) override;
  void Store(Stream* s) override;
  void Load(Stream* s) override;
  T* raw() { return storage_.data(); }
  integer_t my_num_row() const { return my_num_row_; }
  integer_t row_offset() const { return row_offset_; }

 protected:
  void Init(integer_t num_row, integer_t num_col);
  int server_id_;
  integer_t num_col_, my_num_row_, row_offset_;
  std::vector<T> storage_;
  Updater<T>* updater_;
};

}  // namespace multiverso
#endif
