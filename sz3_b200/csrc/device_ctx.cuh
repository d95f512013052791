// sz3_b200/csrc/device_ctx.cuh -- per-thread execution context handed to the kernel bodies on the device:
// thread ids, the CTA barrier and the fused quantization-index histogram (replaces the frequency count of
// HuffmanEncoder::init, reference include/SZ3/encoder/HuffmanEncoder.hpp:516-527).
//
// Histogram strategy: smooth data puts most indices exactly on `radius`; those are counted in a register and
// reduced once per warp at the end (no atomics on the hot bin).  Indices inside a kWindow-wide window around the
// radius go to a shared-memory histogram, everything else (rare) straight to the global 64-bit histogram.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sz3b {

constexpr int kHistWindow = 1024;
constexpr int kTileThreads = 512;

struct DevCtx {
    unsigned *shist;
    unsigned long long *ghist;
    int center;
    int lo;
    unsigned center_cnt;

    __device__ __forceinline__ DevCtx(unsigned *sh, unsigned long long *gh, int radius)
        : shist(sh), ghist(gh), center(radius), lo(radius - kHistWindow / 2), center_cnt(0) {}

    __device__ __forceinline__ uint32_t tid() const { return threadIdx.x; }
    __device__ __forceinline__ uint32_t nthreads() const { return blockDim.x; }
    __device__ __forceinline__ void sync() const { __syncthreads(); }

    __device__ __forceinline__ void clear() {
        for (int i = threadIdx.x; i < kHistWindow; i += blockDim.x) shist[i] = 0;
        __syncthreads();
    }
    __device__ __forceinline__ void hist_add(int sym, bool active) {
        if (!active) return;
        if (sym == center) {
            center_cnt++;
            return;
        }
        unsigned k = static_cast<unsigned>(sym - lo);
        if (k < static_cast<unsigned>(kHistWindow))
            atomicAdd(&shist[k], 1u);
        else
            atomicAdd(&ghist[sym], 1ull);
    }
    __device__ __forceinline__ void pass_end() {}
    __device__ __forceinline__ void flush() {
        unsigned c = center_cnt;
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        if ((threadIdx.x & 31) == 0 && c) atomicAdd(&shist[center - lo], c);
        __syncthreads();
        for (int i = threadIdx.x; i < kHistWindow; i += blockDim.x) {
            unsigned v = shist[i];
            if (v) atomicAdd(&ghist[lo + i], static_cast<unsigned long long>(v));
        }
    }
};

// Context of the line-walker kernel: the 8 bins around the radius are counted in one packed 64-bit register
// (8 bits per bin; a thread handles far fewer than 255 points per pass) and reduced across the warp at pass_end().
struct DevCtx2 {
    unsigned *shist;
    unsigned long long *ghist;
    int lo, lo8;
    unsigned long long packed;

    __device__ __forceinline__ DevCtx2(unsigned *sh, unsigned long long *gh, int radius)
        : shist(sh), ghist(gh), lo(radius - kHistWindow / 2), lo8(radius - 4), packed(0) {}

    __device__ __forceinline__ uint32_t tid() const { return threadIdx.x; }
    __device__ __forceinline__ uint32_t nthreads() const { return blockDim.x; }
    __device__ __forceinline__ void sync() const { __syncthreads(); }

    __device__ __forceinline__ void clear() {
        for (int i = threadIdx.x; i < kHistWindow; i += blockDim.x) shist[i] = 0;
        __syncthreads();
    }
    __device__ __forceinline__ void hist_add(int sym, bool active) {
        if (!active) return;
        const unsigned k8 = static_cast<unsigned>(sym - lo8);
        if (k8 < 8u) {
            packed += 1ull << (k8 * 8u);
            return;
        }
        hist_slow(shist, ghist, lo, sym);
    }
    // rare path (index outside the 8 register bins), out of line so the hot loops carry only a branch; static, so
    // that the context itself never has its address taken (it must stay in registers)
    static __device__ __noinline__ void hist_slow(unsigned *shist, unsigned long long *ghist, int lo, int sym) {
        const unsigned k = static_cast<unsigned>(sym - lo);
        if (k < static_cast<unsigned>(kHistWindow))
            atomicAdd(&shist[k], 1u);
        else
            atomicAdd(&ghist[sym], 1ull);
    }
    // all threads of the CTA, converged.  The eight 8-bit counters are widened to 16-bit fields (even / odd bytes of
    // the two halves: a warp's sum of a field is at most 32 * 255), four warp reductions add them up, and lane b
    // (b < 8) adds bin b to the shared window.
    __device__ __forceinline__ void pass_end() {
        const unsigned plo = static_cast<unsigned>(packed), phi = static_cast<unsigned>(packed >> 32);
        const unsigned e_lo = __reduce_add_sync(0xffffffffu, plo & 0x00ff00ffu);          // bins 0, 2
        const unsigned o_lo = __reduce_add_sync(0xffffffffu, (plo >> 8) & 0x00ff00ffu);   // bins 1, 3
        const unsigned e_hi = __reduce_add_sync(0xffffffffu, phi & 0x00ff00ffu);          // bins 4, 6
        const unsigned o_hi = __reduce_add_sync(0xffffffffu, (phi >> 8) & 0x00ff00ffu);   // bins 5, 7
        const unsigned lane = threadIdx.x & 31u;
        if (lane < 8u) {
            const unsigned word = (lane & 4u) ? ((lane & 1u) ? o_hi : e_hi) : ((lane & 1u) ? o_lo : e_lo);
            const unsigned v = (lane & 2u) ? word >> 16 : word & 0xffffu;
            if (v) atomicAdd(&shist[lo8 - lo + static_cast<int>(lane)], v);
        }
        packed = 0;
    }
    __device__ __forceinline__ void flush() {
        __syncthreads();
        for (int i = threadIdx.x; i < kHistWindow; i += blockDim.x) {
            unsigned v = shist[i];
            if (v) atomicAdd(&ghist[lo + i], static_cast<unsigned long long>(v));
        }
    }
};

}  // namespace sz3b
