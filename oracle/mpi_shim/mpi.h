/* TEST INFRASTRUCTURE ONLY (oracle/): single-rank stand-in for <mpi.h>.
 *
 * The reference (lanl/vpic) requires MPI at configure time
 * (CMakeLists.txt:39) and none is installed in this image.  Its whole MPI
 * surface is the 17 calls in src/util/mp/DMPPolicy.h:103-343.  This header
 * plus mpi_shim.c give those calls rank-0-of-1 semantics so the UNMODIFIED
 * reference sources compile into oracle/_ref/ and act as the parity oracle.
 *
 * One behaviour matters: a single-rank periodic run still "sends to itself"
 * (begin_send_port only skips dst<0||dst>=world_size, src/grid/grid_comm.cc:46-52)
 * so Issend/Irecv pairs are matched by tag, in either posting order.
 */
#ifndef VPIC_B200_ORACLE_MPI_SHIM_H
#define VPIC_B200_ORACLE_MPI_SHIM_H

#ifdef __cplusplus
extern "C" {
#endif

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;

typedef struct shim_mpi_request *MPI_Request;
typedef struct { int byte_count; } MPI_Status;

#define MPI_SUCCESS        0
#define MPI_COMM_WORLD     1
#define MPI_COMM_SELF      2
#define MPI_STATUS_IGNORE  ((MPI_Status *)0)

enum { MPI_BYTE = 1, MPI_CHAR = 2, MPI_INT = 3, MPI_LONG_LONG = 4, MPI_DOUBLE = 5 };
enum { MPI_SUM = 1 };

int MPI_Init(int *argc, char ***argv);
int MPI_Finalize(void);
int MPI_Abort(MPI_Comm comm, int code);
int MPI_Comm_dup(MPI_Comm comm, MPI_Comm *out);
int MPI_Comm_free(MPI_Comm *comm);
int MPI_Comm_rank(MPI_Comm comm, int *rank);
int MPI_Comm_size(MPI_Comm comm, int *size);
int MPI_Barrier(MPI_Comm comm);
int MPI_Allreduce(const void *src, void *dst, int n, MPI_Datatype t, MPI_Op op, MPI_Comm comm);
int MPI_Allgather(const void *src, int ns, MPI_Datatype ts, void *dst, int nd, MPI_Datatype td, MPI_Comm comm);
int MPI_Gather(const void *src, int ns, MPI_Datatype ts, void *dst, int nd, MPI_Datatype td, int root, MPI_Comm comm);
int MPI_Send(const void *buf, int n, MPI_Datatype t, int dst, int tag, MPI_Comm comm);
int MPI_Recv(void *buf, int n, MPI_Datatype t, int src, int tag, MPI_Comm comm, MPI_Status *st);
int MPI_Irecv(void *buf, int n, MPI_Datatype t, int src, int tag, MPI_Comm comm, MPI_Request *req);
int MPI_Issend(const void *buf, int n, MPI_Datatype t, int dst, int tag, MPI_Comm comm, MPI_Request *req);
int MPI_Wait(MPI_Request *req, MPI_Status *st);
int MPI_Get_count(const MPI_Status *st, MPI_Datatype t, int *count);

#ifdef __cplusplus
}
#endif
#endif
