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

namespace SZ3 {
// "a.b.c" -> a<<24 | b<<16 | c<<8
inline uint32_t versionInt(const std::string &version) {
    uint32_t v[4] = {0, 0, 0, 0};
    std::istringstream ss(version);
    std::string part;
    for (int i = 0; i < 4 && std::getline(ss, part, '.'); i++) v[i] = static_cast<uint32_t>(std::stoul(part));
    return (v[0] << 24) | (v[1] << 16) | (v[2] << 8) | v[3];
}
inline std::string versionStr(uint32_t v) {
    std::ostringstream ss;
    ss << ((v >> 24) & 0xff) << "." << ((v >> 16) & 0xff) << "." << ((v >> 8) & 0xff) << "." << (v & 0xff);
    return ss.str();
}
}  // namespace SZ3
#endif
