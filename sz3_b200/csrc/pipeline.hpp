// sz3_b200/csrc/pipeline.hpp -- internal C++ interface between the C ABI (api.cpp) and the GPU pipelines.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/sz3b.h"
#include "workspace.hpp"

namespace sz3b {

struct Error {
    int code;
    std::string msg;
};
[[noreturn]] inline void fail(int code, const std::string &m) { throw Error{code, m}; }

// dispatcher level (SZ_compress_impl / SZ_compress_dispatcher)
template <class T>
size_t compress_any(Workspace &ws, sz3b_config &conf, const T *data, int loc, uint8_t *cmp, size_t cap);
template <class T>
void decompress_any(Workspace &ws, sz3b_config &conf, const uint8_t *cmp, size_t cmp_size, T *out, int loc);

// stage level
template <class T>
void interp_decompose_stage(Workspace &ws, const sz3b_config &conf, double eb, const T *data, int loc, int schedule,
                            int32_t *quant_out, std::vector<uint8_t> &blob);
template <class T>
void blockwise_decompose_stage(Workspace &ws, const sz3b_config &conf, double eb, const T *data, int loc,
                               int32_t *quant_out, std::vector<uint8_t> &blob);
void huffman_encode_stage(Workspace &ws, const int32_t *q, size_t n, int loc, std::vector<uint8_t> &out,
                          size_t *tree_len);
void huffman_decode_stage(Workspace &ws, const uint8_t *in, size_t in_len, size_t tree_len, size_t n, int32_t *out);
size_t lossless_gpu_stage(Workspace &ws, const uint8_t *src, size_t len, int loc, uint8_t *out, size_t cap);
template <class T>
void tune_stage(Workspace &ws, sz3b_config &conf, const T *data, int loc);
template <class T>
double abs_eb_stage(Workspace &ws, const sz3b_config &conf, const T *data, int loc);
template <class T>
void minmax_stage(Workspace &ws, const T *data, int loc, size_t num, double *mn, double *mx);
template <class T>
size_t compress_slab(Workspace &ws, sz3b_config &slab_conf, const T *slab, int loc, double range, uint8_t *payload,
                     size_t cap);

template <class T>
size_t compress_slab_placed(Workspace &ws, sz3b_config &slab_conf, const T *slab, int loc, double range,
                            void *(*place)(void *user, size_t size), void *user);

// devices one SZ_compress call with conf.openmp spreads its slabs over (pipeline.cu: omp_compress)
void set_device_fanout(int n);
int device_fanout();

}  // namespace sz3b
