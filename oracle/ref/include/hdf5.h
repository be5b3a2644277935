/*
 * hdf5.h -- declaration of the ~30-call subset of the HDF5 C API that the reference
 * (AdvancedPhotonSource/xpcs-eigen) uses, so that its UNMODIFIED sources compile in an image
 * without libhdf5 (SURVEY.md Appendix B.3 lists the call sites: configuration.cpp:85-552,
 * h5_result.cpp:56-427, io/hdf5.cpp:62-146).
 *
 * TEST INFRASTRUCTURE (oracle/): implemented by oracle/ref/h5dir.cpp as a directory-backed
 * container -- "file.hdf5" is a directory, groups are sub-directories, a dataset is one file
 * with a 64-byte header followed by the raw elements.  HDF5 is pure I/O for this path (no
 * arithmetic), so substituting it does not affect parity.  Written from the public HDF5 API
 * documentation; no HDF5 or reference code is copied.
 */
#ifndef XPCS_ORACLE_HDF5_SHIM_H
#define XPCS_ORACLE_HDF5_SHIM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int64_t hid_t;
typedef int herr_t;
typedef int htri_t;
typedef unsigned long long hsize_t;
typedef long long hssize_t;

#define H5F_ACC_RDONLY 0u
#define H5F_ACC_RDWR 1u
#define H5P_DEFAULT ((hid_t)0)
#define H5S_ALL ((hid_t)0)
#define H5E_DEFAULT ((hid_t)0)
#define H5T_VARIABLE ((size_t)(-1))

typedef enum { H5S_SELECT_SET = 0 } H5S_seloper_t;
typedef enum { H5T_DIR_DEFAULT = 0, H5T_DIR_ASCEND = 1, H5T_DIR_DESCEND = 2 } H5T_direction_t;

/* predefined native types: fixed ids below 0x100 */
#define H5T_NATIVE_INT ((hid_t)0x11)
#define H5T_NATIVE_LONG ((hid_t)0x12)
#define H5T_NATIVE_FLOAT ((hid_t)0x13)
#define H5T_NATIVE_DOUBLE ((hid_t)0x14)
#define H5T_NATIVE_UINT32 ((hid_t)0x15)
#define H5T_NATIVE_UINT16 ((hid_t)0x16)
#define H5T_NATIVE_SHORT ((hid_t)0x17)
#define H5T_NATIVE_ULONG ((hid_t)0x18)
#define H5T_C_S1 ((hid_t)0x19)

#define H5P_DATASET_CREATE ((hid_t)0x31)

hid_t H5Fopen(const char *name, unsigned flags, hid_t fapl);
herr_t H5Fclose(hid_t f);

hid_t H5Gopen2(hid_t loc, const char *name, hid_t gapl);
hid_t H5Gcreate2(hid_t loc, const char *name, hid_t lcpl, hid_t gcpl, hid_t gapl);
herr_t H5Gclose(hid_t g);
#define H5Gcreate H5Gcreate2

hid_t H5Dopen2(hid_t loc, const char *name, hid_t dapl);
hid_t H5Dcreate2(hid_t loc, const char *name, hid_t type, hid_t space, hid_t lcpl, hid_t dcpl, hid_t dapl);
herr_t H5Dclose(hid_t d);
hid_t H5Dget_space(hid_t d);
hid_t H5Dget_type(hid_t d);
hsize_t H5Dget_storage_size(hid_t d);
herr_t H5Dread(hid_t d, hid_t mem_type, hid_t mem_space, hid_t file_space, hid_t plist, void *buf);
herr_t H5Dwrite(hid_t d, hid_t mem_type, hid_t mem_space, hid_t file_space, hid_t plist, const void *buf);
#define H5Dopen H5Dopen2
#define H5Dcreate H5Dcreate2

hid_t H5Screate_simple(int rank, const hsize_t *dims, const hsize_t *maxdims);
herr_t H5Sclose(hid_t s);
int H5Sget_simple_extent_dims(hid_t s, hsize_t *dims, hsize_t *maxdims);
int H5Sget_simple_extent_ndims(hid_t s);
herr_t H5Sselect_hyperslab(hid_t s, H5S_seloper_t op, const hsize_t *start, const hsize_t *stride,
                           const hsize_t *count, const hsize_t *block);

htri_t H5Tis_variable_str(hid_t t);
hid_t H5Tget_native_type(hid_t t, H5T_direction_t dir);
size_t H5Tget_size(hid_t t);
htri_t H5Tequal(hid_t a, hid_t b);
herr_t H5Tclose(hid_t t);

hid_t H5Pcreate(hid_t cls);
herr_t H5Pclose(hid_t p);
herr_t H5Pset_chunk(hid_t p, int ndims, const hsize_t *dims);
herr_t H5Pset_deflate(hid_t p, unsigned level);

typedef herr_t (*H5E_auto2_t)(hid_t estack, void *client_data);
herr_t H5Eset_auto2(hid_t estack, H5E_auto2_t func, void *client_data);
#define H5Eset_auto H5Eset_auto2

#ifdef __cplusplus
}
#endif
#endif
