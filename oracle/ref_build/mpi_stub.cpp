/* Single-rank no-op definitions of the MPI / HDF5 symbols the reference objects
 * reference.  TEST INFRASTRUCTURE ONLY: lets the unmodified
 * reference sources link into the CPU oracle; none of these is ever reached by the
 * unbinding path (it makes no MPI or I/O call). */
#include <cstring>
#include "mpi.h"
#include "hdf5.h"

extern "C" {
hid_t H5T_NATIVE_INT = 1, H5T_NATIVE_LONG = 2, H5T_NATIVE_FLOAT = 3, H5T_NATIVE_DOUBLE = 4;
herr_t H5Tclose(hid_t) { return 0; }
hid_t H5Dget_space(hid_t) { return 0; }
int H5Sget_simple_extent_dims(hid_t, hsize_t *, hsize_t *) { return 0; }
herr_t H5Sclose(hid_t) { return 0; }
herr_t H5Dvlen_reclaim(hid_t, hid_t, hid_t, void *) { return 0; }
hid_t H5Dopen2(hid_t, const char *, hid_t) { return -1; }
herr_t H5Dread(hid_t, hid_t, hid_t, hid_t, hid_t, void *) { return -1; }
long H5Iget_name(hid_t, char *, size_t) { return 0; }
long H5Fget_name(hid_t, char *, size_t) { return 0; }
herr_t H5Dclose(hid_t) { return 0; }
hid_t H5Aopen_by_name(hid_t, const char *, const char *, hid_t, hid_t) { return -1; }
herr_t H5Aread(hid_t, hid_t, void *) { return -1; }
herr_t H5Aclose(hid_t) { return 0; }

int MPI_Init(int *, char ***) { return 0; }
int MPI_Finalize(void) { return 0; }
int MPI_Finalized(int *f) { *f = 0; return 0; }
int MPI_Abort(MPI_Comm, int) { return 0; }
int MPI_Comm_size(MPI_Comm, int *n) { *n = 1; return 0; }
int MPI_Comm_rank(MPI_Comm, int *r) { *r = 0; return 0; }
int MPI_Get_processor_name(char *name, int *len) { strcpy(name, "oracle"); *len = 6; return 0; }
int MPI_Comm_split(MPI_Comm, int, int, MPI_Comm *c) { *c = 1; return 0; }
int MPI_Comm_free(MPI_Comm *) { return 0; }
int MPI_Barrier(MPI_Comm) { return 0; }
int MPI_Bcast(void *, int, MPI_Datatype, int, MPI_Comm) { return 0; }
int MPI_Allreduce(const void *, void *, int, MPI_Datatype, MPI_Op, MPI_Comm) { return 0; }
int MPI_Reduce(const void *, void *, int, MPI_Datatype, MPI_Op, int, MPI_Comm) { return 0; }
int MPI_Scan(const void *, void *, int, MPI_Datatype, MPI_Op, MPI_Comm) { return 0; }
int MPI_Allgather(const void *, int, MPI_Datatype, void *, int, MPI_Datatype, MPI_Comm) { return 0; }
int MPI_Gather(const void *, int, MPI_Datatype, void *, int, MPI_Datatype, int, MPI_Comm) { return 0; }
int MPI_Alltoall(const void *, int, MPI_Datatype, void *, int, MPI_Datatype, MPI_Comm) { return 0; }
int MPI_Alltoallv(const void *, const int *, const int *, MPI_Datatype, void *, const int *, const int *,
                  MPI_Datatype, MPI_Comm) { return 0; }
int MPI_Alltoallw(const void *, const int *, const int *, const MPI_Datatype *, void *, const int *,
                  const int *, const MPI_Datatype *, MPI_Comm) { return 0; }
int MPI_Scatterv(const void *, const int *, const int *, MPI_Datatype, void *, int, MPI_Datatype, int,
                 MPI_Comm) { return 0; }
int MPI_Isend(const void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Request *) { return 0; }
int MPI_Send(const void *, int, MPI_Datatype, int, int, MPI_Comm) { return 0; }
int MPI_Recv(void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Status *) { return 0; }
int MPI_Probe(int, int, MPI_Comm, MPI_Status *) { return 0; }
int MPI_Get_count(const MPI_Status *, MPI_Datatype, int *c) { *c = 0; return 0; }
int MPI_Waitall(int, MPI_Request *, MPI_Status *) { return 0; }
int MPI_Get_address(const void *p, MPI_Aint *a) { *a = (MPI_Aint)p; return 0; }
int MPI_Address(const void *p, MPI_Aint *a) { *a = (MPI_Aint)p; return 0; }
int MPI_Type_create_struct(int, const int *, const MPI_Aint *, const MPI_Datatype *, MPI_Datatype *t) { *t = 100; return 0; }
int MPI_Type_create_resized(MPI_Datatype, MPI_Aint, MPI_Aint, MPI_Datatype *t) { *t = 101; return 0; }
int MPI_Type_create_hindexed(int, const int *, const MPI_Aint *, MPI_Datatype, MPI_Datatype *t) { *t = 102; return 0; }
int MPI_Type_commit(MPI_Datatype *) { return 0; }
int MPI_Type_free(MPI_Datatype *) { return 0; }
}
