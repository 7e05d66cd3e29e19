/* Minimal single-node MPI subset -- TOOLING for the reference arm of bench.py, not product code.
 *
 * The image has no MPI, so the unmodified Microsoft/multiverso sources cannot be built as
 * shipped (CMakeLists.txt:11 find_package(MPI REQUIRED)). The reference only touches 15 MPI
 * symbols (include/multiverso/net/mpi_net.h); this header + mpi_shim.cpp provide exactly
 * those over a local TCP mesh (ranks from RANK/WORLD_SIZE or MV_SHIM_RANK/MV_SHIM_SIZE) so
 * the reference's own code path -- actors, MPINetWrapper serialisation, tables, updaters,
 * WordEmbedding app -- runs unchanged. */
#ifndef MV_SHIM_MPI_H_
#define MV_SHIM_MPI_H_
#ifdef __cplusplus
extern "C" {
#endif

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Request;
typedef struct MPI_Status {
  int MPI_SOURCE;
  int MPI_TAG;
  int MPI_ERROR;
  int count_bytes;
} MPI_Status;

#define MPI_SUCCESS 0
#define MPI_COMM_WORLD 0
#define MPI_ANY_SOURCE (-1)
#define MPI_ANY_TAG (-1)
#define MPI_IN_PLACE ((void*)1)
#define MPI_BYTE 1
#define MPI_CHAR 2
#define MPI_INT 3
#define MPI_FLOAT 4
#define MPI_DOUBLE 5
#define MPI_SUM 1
#define MPI_THREAD_SINGLE 0
#define MPI_THREAD_FUNNELED 1
#define MPI_THREAD_SERIALIZED 2
#define MPI_THREAD_MULTIPLE 3
#define MPI_STATUS_IGNORE ((MPI_Status*)0)
#define MPI_STATUSES_IGNORE ((MPI_Status*)0)

int MPI_Initialized(int* flag);
int MPI_Init_thread(int* argc, char*** argv, int required, int* provided);
int MPI_Query_thread(int* provided);
int MPI_Comm_rank(MPI_Comm comm, int* rank);
int MPI_Comm_size(MPI_Comm comm, int* size);
int MPI_Barrier(MPI_Comm comm);
int MPI_Finalize(void);
int MPI_Allreduce(const void* sendbuf, void* recvbuf, int count, MPI_Datatype type, MPI_Op op, MPI_Comm comm);
int MPI_Isend(const void* buf, int count, MPI_Datatype type, int dest, int tag, MPI_Comm comm, MPI_Request* req);
int MPI_Recv(void* buf, int count, MPI_Datatype type, int source, int tag, MPI_Comm comm, MPI_Status* status);
int MPI_Iprobe(int source, int tag, MPI_Comm comm, int* flag, MPI_Status* status);
int MPI_Get_count(const MPI_Status* status, MPI_Datatype type, int* count);
int MPI_Wait(MPI_Request* req, MPI_Status* status);
int MPI_Waitall(int count, MPI_Request reqs[], MPI_Status statuses[]);
int MPI_Testall(int count, MPI_Request reqs[], int* flag, MPI_Status statuses[]);

#ifdef __cplusplus
}
#endif
#endif
