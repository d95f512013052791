// sz3_b200/csrc/interp_box.cu -- __global__ wrapper + launcher of the box schedule (interp_box.cuh): the fused
// interpolation-predict + LinearQuantizer kernel of the finest level for float data, pass order z, y, x.
//
// Staging: every raw z-plane of a tile (33 rows x 36 floats) comes in by one TMA box copy (cp.async.bulk.tensor.3d,
// completion on the warp's mbarrier); the warp that owns the plane re-arms its slot for its next plane as soon as
// the rows of the current plane sit in registers, so the copy runs under the pass-2 arithmetic.  Two CTAs of eight
// warps per SM (88 KB of shared memory each) cover each other's phase-A latency.
#include <cuda.h>
#include <cuda_runtime.h>
#include <string.h>

#include "device_ctx.cuh"
#include "interp_box.cuh"
#include "launch.hpp"

namespace sz3b {

namespace {

#ifndef SZ3B_BOX_CWARPS
#define SZ3B_BOX_CWARPS 8
#endif
constexpr int kBoxCWarps = SZ3B_BOX_CWARPS;      // warps of a compressing CTA (k_interp_box)
constexpr int kBoxCThreads = kBoxCWarps * 32;
constexpr bool kBoxAllLines = kBoxCThreads >= kBoxEEPlane;   // every z-line of phase A has a thread of its own
constexpr int kBoxEEPad = 9568;   // EE floats rounded up to a multiple of 32 (128 B)
constexpr size_t kBoxSmem = sizeof(float) * (kBoxCWarps * kBoxSlotStride + kBoxEEPad) + sizeof(uint16_t) * kBoxCWarps * kBoxStageU16 +
                            sizeof(uint64_t) * kBoxCWarps;
struct BoxNoCtx {};   // the box schedule keeps no per-thread context (its histogram is taken afterwards: k_hist_u16)
constexpr unsigned kBoxPlaneBytes = kBoxSlotElems * sizeof(float);   // what one TMA box delivers (zero fill included)

__device__ __forceinline__ unsigned smem_u32(const void *p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
    const unsigned a = smem_u32(bar);
    unsigned done = 0;
    while (!done) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(a), "r"(parity)
            : "memory");
    }
}
__device__ __forceinline__ void tma_plane(float *dst, const CUtensorMap *map, uint64_t *bar, int x, int y, int z) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z)
        : "memory");
}

__device__ __forceinline__ void gather_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }

#ifdef SZ3B_BOX_TIMING
__device__ unsigned long long g_box_timing[64];
__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define BOX_TICK(i) do { if (blockIdx.x == gridDim.x / 2 && threadIdx.x == 0) g_box_timing[i] = gtime(); } while (0)
#else
#define BOX_TICK(i) do { } while (0)
#endif

template <bool CUBIC>
__global__ void __launch_bounds__(kBoxCThreads, 2) k_interp_box(const __grid_constant__ CUtensorMap tmap, BoxArgs A, BoxSrc S) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *const slots = reinterpret_cast<float *>(smem_raw);
    float *const EE = slots + kBoxCWarps * kBoxSlotStride;
    uint16_t *const stages = reinterpret_cast<uint16_t *>(EE + kBoxEEPad);
    uint64_t *const bars = reinterpret_cast<uint64_t *>(stages + kBoxCWarps * kBoxStageU16);
    __shared__ BoxTile T;

    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
    const uint32_t split = S.split, part = blockIdx.x % split;
    const uint32_t tile = blockIdx.x / split + A.tile0;
    const uint32_t zstep = kBoxCWarps * split;
    BOX_TICK(0);
    BoxOrigin o;
    box_origin(A, S, tile, o);
    float *const slot = slots + warp * kBoxSlotStride;
    uint16_t *const stage = stages + warp * kBoxStageU16;
    uint64_t *const bar = bars + warp;
    const uint32_t nz = o.n[0];
    const bool tma = S.tma != 0, write2 = A.s >= 2;
    uint32_t z = (o.begin[0] ? 1u : 0u) + part * kBoxCWarps + warp;   // planes of this warp: z, z + zstep, ...
    const int x0 = static_cast<int>(o.begin[2] / S.odiv), y0 = static_cast<int>(o.begin[1] / S.odiv),
              z0 = static_cast<int>(o.begin[0] / S.odiv);
    if (tma && lane == 0) {
        mbar_init(bar, 1);
        fence_barrier_init();
        if (z < nz) {   // the first plane travels while phase A runs
            mbar_expect_tx(bar, kBoxPlaneBytes);
            tma_plane(slot, &tmap, bar, x0, y0, z0 + static_cast<int>(z));
        }
    }
    BoxNoCtx ctx;
    if (tid == 0) box_tile_setup<CUBIC>(A, tile, o, T);
    // ---- phase A: EE, pass 0 ----------------------------------------------------------------------------------------
    box_fill_column(A, S, o, tid, EE);
    if (!kBoxAllLines && tid < 33) box_fill_column(A, S, o, 256 + tid, EE);
    BOX_TICK(1);
    fill_copy_wait();
    __syncthreads();
    BOX_TICK(2);
    if (!tma && z < nz) box_gather_plane(S, T, lane, z, slot);   // (needs T: after the barrier)
    box_pass0_line<CUBIC>(A, S, ctx, T, tid, EE, part == 0);
    if (!kBoxAllLines)
        for (uint32_t e = tid; e < 33u * 16u; e += kBoxCThreads) box_pass0_left<CUBIC>(A, ctx, T, e, EE, part == 0);
    BOX_TICK(3);
    __syncthreads();
    BOX_TICK(4);
    // ---- phase B: the warp's planes -----------------------------------------------------------------------------------
    const uint32_t lowy = T.low[1], c1y = T.c1[1];
    unsigned parity = 0;
    for (; z < nz; z += zstep) {
        if (tma) {
            mbar_wait(bar, parity);
            parity ^= 1u;
        } else {
            gather_wait();
            __syncwarp();
        }
        BOX_TICK(5);
        const float *const EEz = EE + z * kBoxEEPlane;
        box_merge(T, lane, EEz, slot);
        box_pass1_lane<CUBIC>(A, S, ctx, T, lane, z, EEz, slot);
        box_pass1_left<CUBIC>(A, ctx, T, lane, z, EEz, slot);
        __syncwarp();
        BOX_TICK(6);
        float v[36];
        float *const my_row = slot + (lane + lowy) * kBoxPitch;
        if (lane < c1y) {
            const float4 *row = reinterpret_cast<const float4 *>(my_row);
#pragma unroll
            for (int c = 0; c < 9; c++) {
                const float4 f = row[c];
                v[4 * c] = f.x;
                v[4 * c + 1] = f.y;
                v[4 * c + 2] = f.z;
                v[4 * c + 3] = f.w;
            }
        }
        box_pass2_left<CUBIC>(A, ctx, T, lane, z, slot, stage, write2);
        const bool more = z + zstep < nz;
        if (!write2) {   // slot free: the next plane comes in under the arithmetic of this one
            __syncwarp();
            if (more) {
                if (tma) {
                    if (lane == 0) {
                        fence_proxy_async();
                        mbar_expect_tx(bar, kBoxPlaneBytes);
                        tma_plane(slot, &tmap, bar, x0, y0, z0 + static_cast<int>(z + zstep));
                    }
                } else {
                    box_gather_plane(S, T, lane, z + zstep, slot);
                }
            }
        }
        if (lane < c1y) box_pass2_row<CUBIC>(A, S, ctx, T, lane, z, v, stage, write2 ? my_row : nullptr);
        __syncwarp();
        BOX_TICK(7);
        if (write2) {   // the plane's reconstructions feed the next finer level; the slot is re-armed right after that,
                        // so that the next plane travels while the indices of this one leave
            box_plane_out(A, T, lane, z, slot);
            __syncwarp();
            if (more) {
                if (tma) {
                    if (lane == 0) {
                        fence_proxy_async();
                        mbar_expect_tx(bar, kBoxPlaneBytes);
                        tma_plane(slot, &tmap, bar, x0, y0, z0 + static_cast<int>(z + zstep));
                    }
                } else {
                    box_gather_plane(S, T, lane, z + zstep, slot);
                }
            }
        }
        box_copy_out(A, T, lane, z, stage);
        __syncwarp();
        BOX_TICK(8);
    }
    BOX_TICK(9);
}

// recover, last pass of the finest level (interp_box.cuh: box_recover_row): one CTA per tile, warp per plane; `out` is
// both the source of the planes (tensor map over it) and the destination of the odd-x points.
constexpr size_t kBoxRecSmem = sizeof(float) * (kBoxWarps * kBoxSlotStride) + sizeof(uint16_t) * kBoxWarps * kBoxStageU16 +
                               sizeof(uint64_t) * kBoxWarps;
template <bool CUBIC>
__global__ void __launch_bounds__(kBoxThreads, 4) k_box_recover_x(const __grid_constant__ CUtensorMap tmap, BoxArgs A, BoxSrc S, float *out) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *const slots = reinterpret_cast<float *>(smem_raw);
    uint16_t *const stages = reinterpret_cast<uint16_t *>(slots + kBoxWarps * kBoxSlotStride);
    uint64_t *const bars = reinterpret_cast<uint64_t *>(stages + kBoxWarps * kBoxStageU16);
    __shared__ BoxTile T;
    const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
    const uint32_t tile = blockIdx.x + A.tile0;
    BoxOrigin o;
    box_origin(A, S, tile, o);
    float *const slot = slots + warp * kBoxSlotStride;
    uint16_t *const stage = stages + warp * kBoxStageU16;
    uint64_t *const bar = bars + warp;
    const uint32_t nz = o.n[0];
    uint32_t z = (o.begin[0] ? 1u : 0u) + warp;
    const int x0 = static_cast<int>(o.begin[2]), y0 = static_cast<int>(o.begin[1]), z0 = static_cast<int>(o.begin[0]);
    if (lane == 0) {
        mbar_init(bar, 1);
        fence_barrier_init();
        if (z < nz) {
            mbar_expect_tx(bar, kBoxPlaneBytes);
            tma_plane(slot, &tmap, bar, x0, y0, z0 + static_cast<int>(z));
        }
    }
    if (tid == 0) box_tile_setup<CUBIC>(A, tile, o, T);
    __syncthreads();
    const uint32_t lowy = T.low[1], c1y = T.c1[1];
    unsigned parity = 0;
    for (; z < nz; z += kBoxWarps) {
        box_copy_in(A, T, lane, z, stage);
        mbar_wait(bar, parity);
        parity ^= 1u;
        __syncwarp();
        float v[36];
        float *const my_row = slot + (lane + lowy) * kBoxPitch;
        if (lane < c1y) {
            const float4 *row = reinterpret_cast<const float4 *>(my_row);
#pragma unroll
            for (int c = 0; c < 9; c++) {
                const float4 f = row[c];
                v[4 * c] = f.x;
                v[4 * c + 1] = f.y;
                v[4 * c + 2] = f.z;
                v[4 * c + 3] = f.w;
            }
        }
        box_recover_left<CUBIC>(A, T, lane, z, slot, stage);
        if (lane < c1y) box_recover_row<CUBIC>(A, T, lane, z, v, stage, my_row);
        __syncwarp();
        box_recover_out(S, T, lane, z, slot, out);
        __syncwarp();
        if (z + kBoxWarps < nz && lane == 0) {   // the slot is free again: the warp's next plane
            fence_proxy_async();
            mbar_expect_tx(bar, kBoxPlaneBytes);
            tma_plane(slot, &tmap, bar, x0, y0, z0 + static_cast<int>(z + kBoxWarps));
        }
    }
}

// Compact copies of the coarse lattices: dst_k[z][y][x] = src[z * s_k][y * s_k][x * s_k] for up to three strides
// s_0 < s_1 < s_2 (each twice the previous).  A single tile of a coarse level would gather its ~36 k scattered sectors
// with one SM (tens of microseconds of pure latency); here the whole GPU does it once, and the levels then read dense
// arrays.  One thread per element of the finest of the compact lattices.
struct CompactArgs {
    const float *src;
    uint64_t sstride[3];
    uint32_t s0;
    uint32_t cd[3][3];     // dims of compact array k
    float *dst[3];
    int n;
};
__global__ void __launch_bounds__(256) k_box_compact(CompactArgs C) {
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, z = blockIdx.z;
    if (x >= C.cd[0][2]) return;
    const float v = C.src[(static_cast<uint64_t>(z) * C.sstride[0] + static_cast<uint64_t>(y) * C.sstride[1] + x) * C.s0];
    C.dst[0][(static_cast<uint64_t>(z) * C.cd[0][1] + y) * C.cd[0][2] + x] = v;
    if (C.n > 1 && !((x | y | z) & 1u))
        C.dst[1][(static_cast<uint64_t>(z >> 1) * C.cd[1][1] + (y >> 1)) * C.cd[1][2] + (x >> 1)] = v;
    if (C.n > 2 && !((x | y | z) & 3u))
        C.dst[2][(static_cast<uint64_t>(z >> 2) * C.cd[2][1] + (y >> 2)) * C.cd[2][2] + (x >> 2)] = v;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) != cudaSuccess ||
            qr != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            p = nullptr;
        }
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

}  // namespace

// Can the level with stride A.s run on the box schedule?  (float / 16-bit indices by type; pass order z, y, x; every
// tile 32 or 33 points wide.)
bool interp_box_applicable(const BoxArgs &A) {
    if (A.sh.N != 3 || A.sh.perm[0] != 0 || A.sh.perm[1] != 1 || A.sh.perm[2] != 2) return false;
    if ((reinterpret_cast<uintptr_t>(A.q) & 15u)) return false;
    for (int d = 0; d < 3; d++) {
        // tiles are 33 points wide except the last one
        const uint32_t B = kInterpBlock * A.s;
        const uint32_t last_begin = ((A.sh.dims[d] - 1) / B) * B;
        const uint32_t n_last = (A.sh.dims[d] - 1 - last_begin) / A.s + 1;
        if (n_last != 32 && n_last != 33) return false;
    }
    return true;
}

void interp_launch_compact(const float *src, const uint32_t dims[3], const uint64_t stride[3], uint32_t s0, int n, float *const dst[3],
                           cudaStream_t st) {
    CompactArgs C;
    C.src = src;
    C.s0 = s0;
    C.n = n;
    for (int d = 0; d < 3; d++) C.sstride[d] = stride[d];
    for (int k = 0; k < 3; k++) {
        for (int d = 0; d < 3; d++) C.cd[k][d] = k < n ? (dims[d] - 1) / (s0 << k) + 1 : 1;
        C.dst[k] = k < n ? dst[k] : nullptr;
    }
    const dim3 grid((C.cd[0][2] + 255) / 256, C.cd[0][1], C.cd[0][0]);
    k_box_compact<<<grid, 256, 0, st>>>(C);
}

// Launches tiles [A.tile0, A.tile0 + ntiles) of the level.  `S` says where the level's values come from; with S.tma
// the planes arrive by TMA box copies through a tensor map over S.p (dims `sdims`, dense rows).  Returns false when
// the tensor map cannot be encoded (the caller then takes the gather fill or the line-walker kernel).
bool interp_launch_box(const BoxArgs &A, const BoxSrc &S, const uint32_t sdims[3], uint64_t ntiles, cudaStream_t st) {
    CUtensorMap map;
    memset(&map, 0, sizeof(map));
    if (S.tma) {
        EncodeTiledFn enc = encode_tiled_fn();
        if (!enc) return false;
        if ((reinterpret_cast<uintptr_t>(S.p) & 15u) || (sdims[2] & 3u) || S.st[2] != 1) return false;
        const cuuint64_t gdim[3] = {sdims[2], sdims[1], sdims[0]};
        const cuuint64_t gstr[2] = {S.st[1] * sizeof(float), S.st[0] * sizeof(float)};
        const cuuint32_t box[3] = {static_cast<cuuint32_t>(kBoxPitch), 33u, 1u};
        const cuuint32_t estr[3] = {1u, 1u, 1u};
        if (enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(S.p), gdim, gstr, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return false;
    }
    static std::atomic<unsigned long long> attr_set{0};
    once_per_device(attr_set, [&] {
        cudaFuncSetAttribute(k_interp_box<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kBoxSmem));
        cudaFuncSetAttribute(k_interp_box<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kBoxSmem));
    });
    const dim3 grid(static_cast<unsigned>(ntiles * S.split));
    if (A.sh.cubic)
        k_interp_box<true><<<grid, kBoxCThreads, kBoxSmem, st>>>(map, A, S);
    else
        k_interp_box<false><<<grid, kBoxCThreads, kBoxSmem, st>>>(map, A, S);
    return true;
}

// Last pass (along x) of the finest level in recover mode, tiles [A.tile0, A.tile0 + ntiles): A.q = the decoded
// indices, A.unpred_tmp = the stored values by stream position, `out` = the output array (dims A.sh.dims), whose even-x
// points of the level are final.  Returns false where the kernel does not apply (the caller keeps its per-pass kernel).
bool interp_launch_box_recover_x(const BoxArgs &A, float *out, uint64_t ntiles, cudaStream_t st) {
    if (A.s != 1 || !interp_box_applicable(A)) return false;
    if ((A.sh.dims[2] & 3u) || (reinterpret_cast<uintptr_t>(out) & 15u)) return false;
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return false;
    BoxSrc S;
    memset(&S, 0, sizeof(S));
    S.p = out;
    for (int d = 0; d < 3; d++) S.ost[d] = S.st[d] = A.sh.stride[d];
    S.odiv = 1;
    S.tma = 1;
    S.split = 1;
    CUtensorMap map;
    memset(&map, 0, sizeof(map));
    const cuuint64_t gdim[3] = {A.sh.dims[2], A.sh.dims[1], A.sh.dims[0]};
    const cuuint64_t gstr[2] = {A.sh.stride[1] * sizeof(float), A.sh.stride[0] * sizeof(float)};
    const cuuint32_t box[3] = {static_cast<cuuint32_t>(kBoxPitch), 33u, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    if (enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, out, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return false;
    static std::atomic<unsigned long long> attr_set{0};
    once_per_device(attr_set, [&] {
        cudaFuncSetAttribute(k_box_recover_x<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kBoxRecSmem));
        cudaFuncSetAttribute(k_box_recover_x<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kBoxRecSmem));
    });
    const dim3 grid(static_cast<unsigned>(ntiles));
    if (A.sh.cubic)
        k_box_recover_x<true><<<grid, kBoxThreads, kBoxRecSmem, st>>>(map, A, S, out);
    else
        k_box_recover_x<false><<<grid, kBoxThreads, kBoxRecSmem, st>>>(map, A, S, out);
    return true;
}

#ifdef SZ3B_BOX_TIMING
extern "C" void sz3b_debug_box_timing(unsigned long long *out) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, g_box_timing, sizeof(unsigned long long) * 64);
}
#endif

}  // namespace sz3b
