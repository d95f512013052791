// tests/h5stub/h5stub_driver.cpp -- TEST INFRASTRUCTURE: drives the reference's HDF5 filter plugin
// (tools/H5Z-SZ3/src/H5Z_SZ3.cpp, compiled UNMODIFIED against this repo's drop-in SZ3 headers and the stub hdf5.h of
// this directory) the way libhdf5 drives a dynamically loaded filter: H5PLget_plugin_type / H5PLget_plugin_info, the
// user's set_SZ3_conf_to_H5, the "set local" callback with the chunk's datatype and dataspace, then the filter callback
// on a chunk, forwards and with H5Z_FLAG_REVERSE.  The HDF5 calls the plugin makes land in the in-memory mock below.
//
//   h5z_roundtrip            -> plugin info + set_local only (no GPU needed)
//   h5z_roundtrip gpu        -> additionally compresses / decompresses chunks through the filter callback
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <vector>

#include "H5Z_SZ3.hpp"

// ---- mock of the HDF5 calls ------------------------------------------------------------------------------------------
namespace {
struct MockPlist {
    bool has_filter = false;
    std::vector<unsigned int> cd;
} g_plist;
// type id = 0x1000 | class << 8 | sign << 4 | size;  space id = index into g_spaces + 0x2000
std::vector<std::vector<hsize_t>> g_spaces;
hid_t make_type(H5T_class_t cls, size_t size, bool is_signed) { return 0x1000 | (static_cast<int>(cls) << 8) | ((is_signed ? 1 : 0) << 4) | static_cast<int>(size); }
hid_t make_space(std::vector<hsize_t> dims) {
    g_spaces.push_back(std::move(dims));
    return 0x2000 + static_cast<hid_t>(g_spaces.size() - 1);
}
}  // namespace
extern "C" {
htri_t H5Zfilter_avail(H5Z_filter_t id) { return id == H5Z_FILTER_SZ3 && g_plist.has_filter ? 1 : 0; }
herr_t H5Pset_filter(hid_t, H5Z_filter_t filter, unsigned int, size_t n, const unsigned int cd[]) {
    if (filter != H5Z_FILTER_SZ3) return -1;
    g_plist.has_filter = true;
    g_plist.cd.assign(cd, cd + n);
    return 0;
}
herr_t H5Pmodify_filter(hid_t p, H5Z_filter_t filter, unsigned int flags, size_t n, const unsigned int cd[]) {
    return H5Pset_filter(p, filter, flags, n, cd);
}
herr_t H5Pget_filter_by_id(hid_t, H5Z_filter_t filter, unsigned int, size_t *n, unsigned int cd[], size_t, char[], unsigned int *) {
    if (filter != H5Z_FILTER_SZ3 || !g_plist.has_filter) return -1;
    const size_t k = std::min(*n, g_plist.cd.size());
    memcpy(cd, g_plist.cd.data(), k * sizeof(unsigned int));
    *n = k;
    return 0;
}
H5T_class_t H5Tget_class(hid_t t) { return static_cast<H5T_class_t>((t >> 8) & 0xf); }
size_t H5Tget_size(hid_t t) { return static_cast<size_t>(t & 0xf); }
H5T_sign_t H5Tget_sign(hid_t t) { return ((t >> 4) & 1) ? H5T_SGN_2 : H5T_SGN_NONE; }
int H5Sget_simple_extent_dims(hid_t s, hsize_t dims[], hsize_t[]) {
    const auto &d = g_spaces.at(static_cast<size_t>(s - 0x2000));
    for (size_t i = 0; i < d.size(); i++) dims[i] = d[i];
    return static_cast<int>(d.size());
}
herr_t H5Epush(hid_t, const char *file, const char *func, unsigned line, hid_t, hid_t, hid_t, const char *msg, ...) {
    fprintf(stderr, "[h5stub] H5Epush %s:%u %s: %s\n", file, line, func, msg);
    return 0;
}
}

// ---- the driver ------------------------------------------------------------------------------------------------------
template <class T>
static int roundtrip(const H5Z_class2_t *plugin, H5T_class_t cls, std::vector<hsize_t> dims, double abs_eb, double amp) {
    g_plist = MockPlist();
    // the application: an SZ3 Config with its error bound into the dataset creation property list
    SZ3::Config user;
    user.errorBoundMode = SZ3::EB_ABS;
    user.absErrorBound = abs_eb;
    if (set_SZ3_conf_to_H5(1, user) != 1) return 1;
    // libhdf5, at dataset creation: set_local with the chunk's type and space
    const hid_t type_id = make_type(cls, sizeof(T), std::is_signed<T>::value), space_id = make_space(dims);
    if (plugin->set_local(1, type_id, space_id) != 1) return 2;
    size_t num = 1;
    for (hsize_t d : dims) num *= d;
    std::vector<T> data(num);
    for (size_t i = 0; i < num; i++)
        data[i] = static_cast<T>(amp * (std::sin(0.013 * static_cast<double>(i % 4099)) + 0.3 * std::cos(0.0007 * static_cast<double>(i))));
    // libhdf5, writing a chunk: the filter gets a malloc'ed buffer it may replace
    size_t buf_size = num * sizeof(T);
    void *buf = malloc(buf_size);
    memcpy(buf, data.data(), buf_size);
    const size_t csize = plugin->filter(0, g_plist.cd.size(), g_plist.cd.data(), buf_size, &buf_size, &buf);
    if (csize == 0 || csize >= num * sizeof(T)) return 3;
    // ... and reading it back
    const size_t dsize = plugin->filter(H5Z_FLAG_REVERSE, g_plist.cd.size(), g_plist.cd.data(), csize, &buf_size, &buf);
    if (dsize != num * sizeof(T)) return 4;
    double worst = 0;
    const T *dec = static_cast<const T *>(buf);
    for (size_t i = 0; i < num; i++) worst = std::fmax(worst, std::fabs(static_cast<double>(dec[i]) - static_cast<double>(data[i])));
    free(buf);
    if (!(worst <= abs_eb)) return 5;
    printf("h5z filter %zu-byte %s chunk of %zu elements: ratio %.2f, max error %.3g (bound %.3g)\n", sizeof(T),
           cls == H5T_FLOAT ? "float" : "integer", num, num * sizeof(T) / static_cast<double>(csize), worst, abs_eb);
    return 0;
}

int main(int argc, char **argv) {
    if (H5PLget_plugin_type() != H5PL_TYPE_FILTER) return 10;
    const H5Z_class2_t *plugin = static_cast<const H5Z_class2_t *>(H5PLget_plugin_info());
    if (!plugin || plugin->id != H5Z_FILTER_SZ3 || !plugin->filter || !plugin->set_local || !plugin->encoder_present) return 11;
    {   // set_local: the Config that reaches cd_values carries the chunk's dims and element type
        g_plist = MockPlist();
        SZ3::Config user;
        user.errorBoundMode = SZ3::EB_REL;
        user.relErrorBound = 1e-3;
        if (set_SZ3_conf_to_H5(1, user) != 1) return 12;
        if (plugin->set_local(1, make_type(H5T_INTEGER, 8, true), make_space({20, 30, 40})) != 1) return 13;
        SZ3::Config got;
        const unsigned char *p = reinterpret_cast<const unsigned char *>(g_plist.cd.data());
        got.load(p);
        if (got.N != 3 || got.num != 24000 || got.dataType != SZ_INT64 || got.errorBoundMode != SZ3::EB_REL || got.relErrorBound != 1e-3)
            return 14;
    }
    if (argc < 2 || strcmp(argv[1], "gpu")) {
        printf("h5z plugin info and set_local ok\n");
        return 0;
    }
    int rc;
    if ((rc = roundtrip<float>(plugin, H5T_FLOAT, {64, 64, 64}, 1e-3, 1.0))) return 20 + rc;
    if ((rc = roundtrip<double>(plugin, H5T_FLOAT, {48, 80, 40}, 1e-6, 1.0))) return 30 + rc;
    if ((rc = roundtrip<int32_t>(plugin, H5T_INTEGER, {40, 50, 60}, 2.0, 3000.0))) return 40 + rc;
    if ((rc = roundtrip<int64_t>(plugin, H5T_INTEGER, {200, 300}, 500.0, 1.0e9))) return 50 + rc;
    // an element type outside the GPU path is refused loudly (no CPU fallback behind the boundary)
    try {
        roundtrip<uint16_t>(plugin, H5T_INTEGER, {32, 32, 32}, 2.0, 100.0);
        return 60;
    } catch (const std::runtime_error &e) {
        printf("uint16 chunk refused: %s\n", e.what());
    }
    printf("h5z roundtrip ok\n");
    return 0;
}
