/* C API of the host runtime: ABI-compatible with the reference (include/multiverso/c_api.h:
 * 14-54: float-only array and matrix tables, TableHandler = void*), plus 64-bit entry
 * points (the reference's int sizes cap a table at 2^31 elements) and the
 * extras the Python package needs (flags, KV table, aggregate, checkpoint, roles). */
#ifndef MULTIVERSO_C_API_H_
#define MULTIVERSO_C_API_H_
#include <stdint.h>

#if defined _WIN32
#define DllExport __declspec(dllexport)
#else
#define DllExport __attribute__((visibility("default")))
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef void* TableHandler;

DllExport void MV_Init(int* argc, char* argv[]);
DllExport void MV_ShutDown();
DllExport void MV_Barrier();
DllExport int MV_NumWorkers();
DllExport int MV_WorkerId();
DllExport int MV_ServerId();

/* Array table */
DllExport void MV_NewArrayTable(int size, TableHandler* out);
DllExport void MV_GetArrayTable(TableHandler handler, float* data, int size);
DllExport void MV_AddArrayTable(TableHandler handler, float* data, int size);
DllExport void MV_AddAsyncArrayTable(TableHandler handler, float* data, int size);

/* Matrix table */
DllExport void MV_NewMatrixTable(int num_row, int num_col, TableHandler* out);
DllExport void MV_GetMatrixTableAll(TableHandler handler, float* data, int size);
DllExport void MV_AddMatrixTableAll(TableHandler handler, float* data, int size);
DllExport void MV_AddAsyncMatrixTableAll(TableHandler handler, float* data, int size);
DllExport void MV_GetMatrixTableByRows(TableHandler handler, float* data, int size, int row_ids[],
                                       int row_ids_n);
DllExport void MV_AddMatrixTableByRows(TableHandler handler, float* data, int size, int row_ids[],
                                       int row_ids_n);
DllExport void MV_AddAsyncMatrixTableByRows(TableHandler handler, float* data, int size,
                                            int row_ids[], int row_ids_n);

/* ---- extensions ------------------------------------------------------------------------- */
DllExport void MV_ShutDownEx(int finalize_net);
DllExport int MV_Rank();
DllExport int MV_Size();
DllExport int MV_NumServers();
DllExport int MV_WorkerIdToRank(int worker_id);
DllExport int MV_ServerIdToRank(int server_id);
DllExport int MV_SetFlagInt(const char* name, int value);
DllExport int MV_SetFlagBool(const char* name, int value);
DllExport int MV_SetFlagDouble(const char* name, double value);
DllExport int MV_SetFlagString(const char* name, const char* value);
DllExport int MV_NetBindC(int rank, const char* endpoint);
DllExport int MV_NetConnectC(int* ranks, const char* endpoints[], int size);
DllExport void MV_NetFinalizeC();
DllExport void MV_AggregateFloat(float* data, int64_t size);
DllExport void MV_AggregateDouble(double* data, int64_t size);
DllExport void MV_AggregateInt(int* data, int64_t size);
DllExport void MV_AggregateChar(char* data, int64_t size);

/* typed tables: dtype 0 = float, 1 = double, 2 = int */
DllExport void MV_NewArrayTable64(int64_t size, int dtype, TableHandler* out);
DllExport void MV_GetArrayTable64(TableHandler h, int dtype, void* data, int64_t size);
DllExport void MV_AddArrayTable64(TableHandler h, int dtype, void* data, int64_t size,
                                  const void* add_option20, int async);
DllExport void MV_NewMatrixTable64(int64_t num_row, int64_t num_col, int dtype, int is_sparse,
                                   int is_pipeline, int random_init, double min_value,
                                   double max_value, TableHandler* out);
DllExport void MV_GetMatrixTable64(TableHandler h, int dtype, void* data, int64_t size,
                                   const int64_t* row_ids, int64_t row_ids_n, int worker_id_opt);
DllExport void MV_AddMatrixTable64(TableHandler h, int dtype, void* data, int64_t size,
                                   const int64_t* row_ids, int64_t row_ids_n,
                                   const void* add_option20, int async);
DllExport void MV_NewKVTable(int val_dtype, TableHandler* out);   /* keys int64 */
DllExport void MV_KVAdd(TableHandler h, int val_dtype, const int64_t* keys, const void* vals,
                        int64_t n);
DllExport void MV_KVGet(TableHandler h, int val_dtype, const int64_t* keys, void* vals, int64_t n);
DllExport int MV_TableId(TableHandler h);
DllExport int MV_SaveTableC(int table_id, const char* uri);
DllExport int MV_LoadTableC(int table_id, const char* uri);
DllExport void MV_DashboardDisplay();
DllExport const char* MV_Version();

#ifdef __cplusplus
}
#endif
#endif
