/* include/vbz_plugin.h - HDF5 dynamically-loaded filter entry points of libvbz_b200.so.
 *
 * Replaces the reference's libvbz_hdf_plugin.so (vbz_plugin/vbz_plugin.cpp:231-261, filter id and
 * cd_values layout vbz_plugin/vbz_plugin.h:5-10): libhdf5 dlopen()s the library found under
 * HDF5_PLUGIN_PATH, calls H5PLget_plugin_type() / H5PLget_plugin_info() and from then on calls
 * H5Z_class2_t::filter once per chunk.  No libhdf5 symbol is needed on Linux (the reference does
 * not link it either).
 */
#ifndef VBZ_B200_PLUGIN_H
#define VBZ_B200_PLUGIN_H

#include <stddef.h>
#include "vbz.h"

#if defined(__cplusplus)
extern "C" {
#endif

#define FILTER_VBZ_ID 32020
/* cd_values[] layout (reference vbz_plugin/vbz_plugin.h:7-10) */
#define FILTER_VBZ_VERSION_OPTION 0
#define FILTER_VBZ_INTEGER_SIZE_OPTION 1
#define FILTER_VBZ_USE_DELTA_ZIG_ZAG_COMPRESSION 2
#define FILTER_VBZ_ZSTD_COMPRESSION_LEVEL_OPTION 3

#define VBZ_H5Z_FLAG_REVERSE 0x0100 /* H5Z_FLAG_REVERSE: decode */

/* H5Z_func_t / H5Z_class2_t as libhdf5 declares them (H5Zpublic.h; the reference vendors the same
 * declarations in third_party/hdf5/hdf5_plugin_types.h:49-75) */
typedef size_t (*vbz_h5z_func_t)(unsigned int flags, size_t cd_nelmts, const unsigned int cd_values[], size_t nbytes,
                                 size_t* buf_size, void** buf);
typedef struct vbz_h5z_class2_t {
    int version;              /* H5Z_CLASS_T_VERS = 1 */
    int id;                   /* 32020 */
    unsigned encoder_present; /* 1 */
    unsigned decoder_present; /* 1 */
    const char* name;         /* "vbz" */
    void* can_apply;          /* NULL */
    void* set_local;          /* NULL */
    vbz_h5z_func_t filter;
} vbz_h5z_class2_t;

/* reference vbz_plugin/vbz_plugin.cpp:248-251: H5PL_TYPE_FILTER (0) */
VBZ_EXPORT int H5PLget_plugin_type(void);
/* reference vbz_plugin/vbz_plugin.cpp:254-261 */
VBZ_EXPORT const void* H5PLget_plugin_info(void);
/* reference vbz_plugin/vbz_plugin.cpp:242-245 */
VBZ_EXPORT const void* vbz_plugin_info(void);
/* The filter callback itself (reference vbz_plugin/vbz_plugin.cpp:97-229; there it has C++ linkage and
 * is reachable only through the struct - exported here under a C name as well for tests). */
VBZ_EXPORT size_t vbz_b200_h5z_filter(unsigned int flags, size_t cd_nelmts, const unsigned int cd_values[], size_t nbytes,
                                      size_t* buf_size, void** buf);

#if defined(__cplusplus)
}
#endif
#endif
