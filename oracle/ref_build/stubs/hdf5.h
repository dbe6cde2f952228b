/* Type-and-prototype-only stand-in for <hdf5.h>; see mpi.h in this directory.
 * Only what src/hdf_wrapper.h's inline helpers mention.  TEST INFRASTRUCTURE. */
#ifndef HBT_ORACLE_HDF5_STUB_H
#define HBT_ORACLE_HDF5_STUB_H
#include <stddef.h>
typedef long hid_t;
typedef unsigned long long hsize_t;
typedef int herr_t;
typedef long ssize_t_h5;
typedef struct { size_t len; void *p; } hvl_t;
#define H5P_DEFAULT 0
#define H5S_ALL 0
#ifdef __cplusplus
extern "C" {
#endif
extern hid_t H5T_NATIVE_INT, H5T_NATIVE_LONG, H5T_NATIVE_FLOAT, H5T_NATIVE_DOUBLE;
hid_t H5Dget_space(hid_t);
int H5Sget_simple_extent_dims(hid_t, hsize_t *, hsize_t *);
herr_t H5Sclose(hid_t);
herr_t H5Dvlen_reclaim(hid_t, hid_t, hid_t, void *);
hid_t H5Dopen2(hid_t, const char *, hid_t);
herr_t H5Dread(hid_t, hid_t, hid_t, hid_t, hid_t, void *);
long H5Iget_name(hid_t, char *, size_t);
long H5Fget_name(hid_t, char *, size_t);
herr_t H5Dclose(hid_t);
hid_t H5Aopen_by_name(hid_t, const char *, const char *, hid_t, hid_t);
herr_t H5Aread(hid_t, hid_t, void *);
herr_t H5Aclose(hid_t);
herr_t H5Tclose(hid_t);
#ifdef __cplusplus
}
#endif
#endif
