// sz3_b200/csrc/decompress.cu -- SZ_decompress path (placeholder until the recover kernels land).
#include "pipeline.hpp"
namespace sz3b {
template <class T>
void decompress_any(Workspace &, sz3b_config &, const uint8_t *, size_t, T *, int) {
    fail(SZ3B_E_UNSUPPORTED, "decompression is not on the GPU path yet");
}
template void decompress_any<float>(Workspace &, sz3b_config &, const uint8_t *, size_t, float *, int);
template void decompress_any<double>(Workspace &, sz3b_config &, const uint8_t *, size_t, double *, int);
}
