// sz3_b200/csrc/blockwise.cu -- BlockwiseDecomposition path (placeholder until the kernels land).
#include "pipeline.hpp"
namespace sz3b {
template <class T>
size_t blockwise_compress(Workspace &, const sz3b_config &, const T *, uint8_t *, size_t, int) {
    fail(SZ3B_E_UNSUPPORTED, "ALGO_LORENZO_REG is not on the GPU path yet");
}
template <class T>
void blockwise_decompose_stage(Workspace &, const sz3b_config &, double, const T *, int, int32_t *,
                               std::vector<uint8_t> &) {
    fail(SZ3B_E_UNSUPPORTED, "ALGO_LORENZO_REG is not on the GPU path yet");
}
template size_t blockwise_compress<float>(Workspace &, const sz3b_config &, const float *, uint8_t *, size_t, int);
template size_t blockwise_compress<double>(Workspace &, const sz3b_config &, const double *, uint8_t *, size_t, int);
template void blockwise_decompose_stage<float>(Workspace &, const sz3b_config &, double, const float *, int, int32_t *, std::vector<uint8_t> &);
template void blockwise_decompose_stage<double>(Workspace &, const sz3b_config &, double, const double *, int, int32_t *, std::vector<uint8_t> &);
}
