/* tests/h5stub/H5PLextern.h -- TEST INFRASTRUCTURE, see hdf5.h in this directory. */
#ifndef SZ3B_TEST_H5STUB_H5PLEXTERN_H
#define SZ3B_TEST_H5STUB_H5PLEXTERN_H
#include "hdf5.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef enum { H5PL_TYPE_ERROR = -1, H5PL_TYPE_FILTER = 0 } H5PL_type_t;
H5PL_type_t H5PLget_plugin_type(void);
const void *H5PLget_plugin_info(void);
#ifdef __cplusplus
}
#endif
#endif
