/* Type-and-prototype-only stand-in for <mpi.h>, used ONLY to compile the
 * unmodified reference sources (/root/reference/src) as a single-rank CPU
 * oracle.  TEST INFRASTRUCTURE - never linked into
 * the product library. */
#ifndef HBT_ORACLE_MPI_STUB_H
#define HBT_ORACLE_MPI_STUB_H
#include <stddef.h>
#define MPI_VERSION 3
typedef int MPI_Datatype;
typedef int MPI_Comm;
typedef int MPI_Op;
typedef int MPI_Request;
typedef long MPI_Aint;
typedef struct MPI_Status { int MPI_SOURCE, MPI_TAG, MPI_ERROR; } MPI_Status;
#define MPI_COMM_WORLD 0
#define MPI_MAX_PROCESSOR_NAME 256
#define MPI_SUCCESS 0
enum { MPI_CHAR = 1, MPI_INT, MPI_LONG, MPI_FLOAT, MPI_DOUBLE, MPI_2INT, MPI_LONG_INT,
       MPI_UNSIGNED, MPI_UNSIGNED_LONG, MPI_BYTE, MPI_LONG_LONG, MPI_C_BOOL, MPI_DATATYPE_NULL };
enum { MPI_SUM = 1, MPI_MIN, MPI_MAX, MPI_MAXLOC, MPI_MINLOC, MPI_LOR, MPI_BOR, MPI_LAND, MPI_BAND };
#define MPI_BOTTOM ((void *)0)
#define MPI_IN_PLACE ((void *)1)
#define MPI_STATUS_IGNORE ((MPI_Status *)0)
#define MPI_STATUSES_IGNORE ((MPI_Status *)0)
#define MPI_ANY_SOURCE (-1)
#define MPI_ANY_TAG (-1)
#ifdef __cplusplus
extern "C" {
#endif
int MPI_Init(int *, char ***);
int MPI_Finalize(void);
int MPI_Finalized(int *);
int MPI_Abort(MPI_Comm, int);
int MPI_Comm_size(MPI_Comm, int *);
int MPI_Comm_rank(MPI_Comm, int *);
int MPI_Get_processor_name(char *, int *);
int MPI_Comm_split(MPI_Comm, int, int, MPI_Comm *);
int MPI_Comm_free(MPI_Comm *);
int MPI_Barrier(MPI_Comm);
int MPI_Bcast(void *, int, MPI_Datatype, int, MPI_Comm);
int MPI_Allreduce(const void *, void *, int, MPI_Datatype, MPI_Op, MPI_Comm);
int MPI_Reduce(const void *, void *, int, MPI_Datatype, MPI_Op, int, MPI_Comm);
int MPI_Scan(const void *, void *, int, MPI_Datatype, MPI_Op, MPI_Comm);
int MPI_Allgather(const void *, int, MPI_Datatype, void *, int, MPI_Datatype, MPI_Comm);
int MPI_Gather(const void *, int, MPI_Datatype, void *, int, MPI_Datatype, int, MPI_Comm);
int MPI_Alltoall(const void *, int, MPI_Datatype, void *, int, MPI_Datatype, MPI_Comm);
int MPI_Alltoallv(const void *, const int *, const int *, MPI_Datatype, void *, const int *,
                  const int *, MPI_Datatype, MPI_Comm);
int MPI_Alltoallw(const void *, const int *, const int *, const MPI_Datatype *, void *,
                  const int *, const int *, const MPI_Datatype *, MPI_Comm);
int MPI_Scatterv(const void *, const int *, const int *, MPI_Datatype, void *, int, MPI_Datatype,
                 int, MPI_Comm);
int MPI_Isend(const void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Request *);
int MPI_Send(const void *, int, MPI_Datatype, int, int, MPI_Comm);
int MPI_Recv(void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Status *);
int MPI_Probe(int, int, MPI_Comm, MPI_Status *);
int MPI_Get_count(const MPI_Status *, MPI_Datatype, int *);
int MPI_Waitall(int, MPI_Request *, MPI_Status *);
int MPI_Get_address(const void *, MPI_Aint *);
int MPI_Address(const void *, MPI_Aint *);
int MPI_Type_create_struct(int, const int *, const MPI_Aint *, const MPI_Datatype *, MPI_Datatype *);
int MPI_Type_create_resized(MPI_Datatype, MPI_Aint, MPI_Aint, MPI_Datatype *);
int MPI_Type_create_hindexed(int, const int *, const MPI_Aint *, MPI_Datatype, MPI_Datatype *);
int MPI_Type_commit(MPI_Datatype *);
int MPI_Type_free(MPI_Datatype *);
#ifdef __cplusplus
}
#endif
#endif
