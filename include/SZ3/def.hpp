// SZ3/def.hpp -- drop-in replacement header (sz3_b200).  Same names as the reference's include/SZ3/def.hpp:1-21.
#ifndef SZ3_DEF_HPP
#define SZ3_DEF_HPP

namespace SZ3 {
typedef unsigned int uint;
typedef unsigned char uchar;
}  // namespace SZ3

#define SZ3_ERROR_COMP_BUFFER_NOT_LARGE_ENOUGH "The buffer for compressed data is not large enough."

#if defined(__GNUC__) || defined(__clang__)
#define ALWAYS_INLINE inline __attribute__((always_inline))
#else
#define ALWAYS_INLINE inline
#endif

#endif
