// Microbenchmark: issue cost of the float <-> double conversions of the quantizer (F2F on the conversion unit) against
// an integer-built widening.  nvcc -arch=sm_100a -O3 -o build/f2f_bench tools/exp/f2f_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ double widen_int(float f) {   // exact for normal floats
    const unsigned u = __float_as_uint(f);
    const unsigned hi = (u & 0x80000000u) | (((u & 0x7fffffffu) >> 3) + 0x38000000u);
    return __hiloint2double(hi, u << 29);
}
template <int MODE>
__global__ void k(float *out, int iters) {
    float a[8];
#pragma unroll
    for (int j = 0; j < 8; j++) a[j] = 1.0f + threadIdx.x * 1e-3f + j;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
            if (MODE == 0) {   // widen + narrow: 2 F2F, 1 DADD
                double d = static_cast<double>(a[j]);
                a[j] = static_cast<float>(d + 1e-9);
            } else if (MODE == 1) {   // integer widen + F2F narrow
                double d = widen_int(a[j]);
                a[j] = static_cast<float>(d + 1e-9);
            } else if (MODE == 2) {   // no conversions at all: DADD chain only (baseline)
                double d = __hiloint2double(__float_as_int(a[j]), i);
                d = d + 1e-9;
                a[j] = __int_as_float(__double2hiint(d));
            } else {   // float only
                a[j] = a[j] * 1.0000001f + 1e-9f;
            }
        }
    }
    float s = 0;
    for (int j = 0; j < 8; j++) s += a[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(const char *name) {
    float *out;
    cudaMalloc(&out, 148 * 4 * 512 * 4);
    const int iters = 4096;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<MODE><<<148 * 4, 512>>>(out, iters);
    cudaEventRecord(e0);
    k<MODE><<<148 * 4, 512>>>(out, iters);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    // per SM: 4 CTAs * 16 warps = 64 warps; each does iters * 8 units
    const double units_per_smsp = 16.0 * iters * 8;   // warp-units per scheduler
    printf("%-40s %.3f ms  -> %.2f ns per warp-unit per scheduler = %.1f cycles @1.965GHz\n", name, ms, ms * 1e6 / units_per_smsp,
           ms * 1e6 / units_per_smsp * 1.965);
    cudaFree(out);
}
int main() {
    run<0>("F2F widen + DADD + F2F narrow");
    run<1>("int widen + DADD + F2F narrow");
    run<2>("DADD only (bit casts)");
    run<3>("FMUL+FADD only");
    return 0;
}
