/* tests/h5stub/hdf5.h -- TEST INFRASTRUCTURE: a minimal stand-in for the part of the HDF5 C API that the reference's
 * HDF5 filter plugin (tools/H5Z-SZ3/src/H5Z_SZ3.cpp, include/H5Z_SZ3.hpp) touches.  HDF5 itself is not in this image,
 * so the plugin source -- UNMODIFIED, read where it lies under /root/reference -- is compiled against this header and
 * this repo's drop-in SZ3 headers, and driven the way libhdf5 drives a filter (tests/test_h5z_filter.py): plugin info,
 * set_local, then the filter callback forwards and in reverse.  Only the names, types and call shapes are modelled; the
 * functions are implemented by the test driver (h5stub_driver.cpp).  Not a substitute for testing inside libhdf5. */
#ifndef SZ3B_TEST_H5STUB_HDF5_H
#define SZ3B_TEST_H5STUB_HDF5_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef int64_t hid_t;
typedef int herr_t;
typedef int htri_t;
typedef unsigned long long hsize_t;
typedef int H5Z_filter_t;
#define H5S_MAX_RANK 32
#define H5Z_CLASS_T_VERS 1
#define H5Z_FLAG_MANDATORY 0x0000
#define H5Z_FLAG_OPTIONAL 0x0001
#define H5Z_FLAG_REVERSE 0x0100
#define H5E_DEFAULT ((hid_t)0)
#define H5E_PLINE ((hid_t)101)
#define H5E_ARGS ((hid_t)102)
#define H5E_BADTYPE ((hid_t)201)
#define H5E_BADVALUE ((hid_t)202)
typedef enum { H5T_NO_CLASS = -1, H5T_INTEGER = 0, H5T_FLOAT = 1, H5T_STRING = 3 } H5T_class_t;
typedef enum { H5T_SGN_ERROR = -1, H5T_SGN_NONE = 0, H5T_SGN_2 = 1 } H5T_sign_t;
typedef htri_t (*H5Z_can_apply_func_t)(hid_t dcpl_id, hid_t type_id, hid_t space_id);
typedef herr_t (*H5Z_set_local_func_t)(hid_t dcpl_id, hid_t type_id, hid_t space_id);
typedef size_t (*H5Z_func_t)(unsigned int flags, size_t cd_nelmts, const unsigned int cd_values[], size_t nbytes, size_t *buf_size,
                             void **buf);
typedef struct H5Z_class2_t {
    int version;
    H5Z_filter_t id;
    unsigned encoder_present;
    unsigned decoder_present;
    const char *name;
    H5Z_can_apply_func_t can_apply;
    H5Z_set_local_func_t set_local;
    H5Z_func_t filter;
} H5Z_class2_t;
htri_t H5Zfilter_avail(H5Z_filter_t id);
herr_t H5Pset_filter(hid_t plist_id, H5Z_filter_t filter, unsigned int flags, size_t cd_nelmts, const unsigned int cd_values[]);
herr_t H5Pmodify_filter(hid_t plist_id, H5Z_filter_t filter, unsigned int flags, size_t cd_nelmts, const unsigned int cd_values[]);
herr_t H5Pget_filter_by_id(hid_t plist_id, H5Z_filter_t filter_id, unsigned int flags, size_t *cd_nelmts, unsigned int cd_values[],
                           size_t namelen, char name[], unsigned int *filter_config);
H5T_class_t H5Tget_class(hid_t type_id);
size_t H5Tget_size(hid_t type_id);
H5T_sign_t H5Tget_sign(hid_t type_id);
int H5Sget_simple_extent_dims(hid_t space_id, hsize_t dims[], hsize_t maxdims[]);
herr_t H5Epush(hid_t err_stack, const char *file, const char *func, unsigned line, hid_t cls_id, hid_t maj_id, hid_t min_id,
               const char *msg, ...);
#ifdef __cplusplus
}
#endif
#endif
