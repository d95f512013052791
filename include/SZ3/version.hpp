// SZ3/version.hpp -- drop-in replacement header (sz3_b200).  The reference generates this file from
// include/SZ3/version.hpp.in:1-40; the values below are those of the v3.3.2 source tree this library is parity-tested
// against (stream format "data version" 3.3.2).
#ifndef SZ3_VERSION_HPP
#define SZ3_VERSION_HPP
#include <cstdint>
#include <sstream>
#include <string>

#define SZ3_NAME "SZ3"
#define SZ3_VER "3.3.2"
#define SZ3_VER_MAJOR 3
#define SZ3_VER_MINOR 3
#define SZ3_VER_PATCH 2
#define SZ3_VER_TWEAK 0
#define SZ3_DATA_VER "3.3.2"
#define SZ3_MAGIC_NUMBER 0xF342F310u

// (global functions, as in the reference: callers such as tools/H5Z-SZ3/src/H5Z_SZ3.cpp:147 use them unqualified)
// "a.b.c" -> a<<24 | b<<16 | c<<8
inline uint32_t versionInt(const std::string &version) {
    uint32_t major = 0, minor = 0, patch = 0;
    char dot;
    std::stringstream ss(version);
    ss >> major >> dot >> minor >> dot >> patch;
    return (major << 24) | (minor << 16) | (patch << 8);
}
inline std::string versionStr(uint32_t version) {
    return std::to_string((version >> 24) & 0xFF) + "." + std::to_string((version >> 16) & 0xFF) + "." +
           std::to_string((version >> 8) & 0xFF);
}
#endif
