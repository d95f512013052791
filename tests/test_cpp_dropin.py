"""The drop-in C++ surface: a program written against the reference's own API (SZ3/api/sz.hpp, SZ3::Config, sz3c.h)
must compile unchanged against include/ and link against sz3_b200/lib.  CPU part: compile + link + the Config-only
behaviour (no GPU needed).  GPU part: run the round trip."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = r'''
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>
#include "SZ3/api/sz.hpp"
#include "sz3c.h"
int main(int argc, char **argv) {
    SZ3::Config conf(40, 50, 60);
    if (conf.N != 3 || conf.num != 40u * 50u * 60u || conf.blockSize != 6 || conf.cmprAlgo != SZ3::ALGO_INTERP_LORENZO) return 10;
    conf.load_ini("[GlobalSettings]\nCmprAlgo = ALGO_INTERP\nErrorBoundMode = REL\nRelErrorBound = 1e-3\n[AlgoSettings]\nInterpolationAlgo = INTERP_ALGO_LINEAR\n");
    if (conf.cmprAlgo != SZ3::ALGO_INTERP || conf.errorBoundMode != SZ3::EB_REL || conf.interpAlgo != SZ3::INTERP_ALGO_LINEAR) return 11;
    unsigned char blob[256];
    unsigned char *p = blob;
    size_t n = conf.save(p);
    SZ3::Config back;
    const unsigned char *q = blob;
    back.load(q);
    if (n == 0 || back.N != 3 || back.dims != conf.dims || back.relErrorBound != 1e-3 || back.cmprAlgo != SZ3::ALGO_INTERP) return 12;
    SZ3::Config one(1, 7, 1);
    if (one.N != 1 || one.dims[0] != 7) return 13;
    if (argc < 2 || strcmp(argv[1], "gpu")) { printf("config ok\n"); return 0; }
    // round trip through SZ_compress / SZ_decompress and the C shim
    SZ3::Config c2(40, 50, 60);
    c2.absErrorBound = 1e-3;
    std::vector<float> data(c2.num);
    for (size_t i = 0; i < data.size(); i++) data[i] = std::sin(0.01f * i) + 0.001f * (i % 7);
    size_t cmpSize = 0;
    char *cmp = SZ_compress<float>(c2, data.data(), cmpSize);
    SZ3::Config dc;
    float *dec = nullptr;
    SZ_decompress<float>(dc, cmp, cmpSize, dec);
    if (dc.num != c2.num || dc.N != 3) return 20;
    double worst = 0;
    for (size_t i = 0; i < data.size(); i++) worst = std::fmax(worst, std::fabs((double)dec[i] - data[i]));
    if (!(worst <= 1e-3)) return 21;
    delete[] dec;
    delete[] cmp;
    try {
        char small[8];
        SZ_compress<float>(c2, data.data(), small, sizeof small);
        return 22;
    } catch (const std::invalid_argument &) {}
    size_t outSize = 0;
    unsigned char *c = SZ_compress_args(SZ_FLOAT, data.data(), &outSize, ABS, 1e-3, 0, 0, 0, 0, 40, 50, 60);
    float *d2 = (float *)SZ_decompress(SZ_FLOAT, c, outSize, 0, 0, 40, 50, 60);
    worst = 0;
    for (size_t i = 0; i < data.size(); i++) worst = std::fmax(worst, std::fabs((double)d2[i] - data[i]));
    free_buf(c);
    free_buf(d2);
    if (!(worst <= 1e-3)) return 23;
    // conf.openmp: the slab container, one slab per GPU of the box, all driven from this one call
    SZ3::Config c3(64, 50, 60);
    c3.absErrorBound = 1e-3;
    c3.openmp = true;
    std::vector<float> big(c3.num);
    for (size_t i = 0; i < big.size(); i++) big[i] = std::sin(0.01f * i) + 0.001f * (i % 7);
    size_t cmpSize3 = 0;
    char *cmp3 = SZ_compress<float>(c3, big.data(), cmpSize3);
    int nslab = 0;
    memcpy(&nslab, cmp3 + 16, 4);
    SZ3::Config dc3;
    float *dec3 = SZ_decompress<float>(dc3, cmp3, cmpSize3);
    worst = 0;
    for (size_t i = 0; i < big.size(); i++) worst = std::fmax(worst, std::fabs((double)dec3[i] - big[i]));
    delete[] dec3;
    delete[] cmp3;
    if (!(worst <= 1e-3) || !dc3.openmp) return 24;
    if (nslab != sz3b_device_count()) return 25;
    // the integer element types the reference's CLI instantiates (tools/sz3/sz3.cpp:458-461)
    {
        SZ3::Config ci(40, 50, 60);
        ci.absErrorBound = 2.0;
        std::vector<int32_t> a32(ci.num);
        std::vector<int64_t> a64(ci.num);
        for (size_t i = 0; i < a32.size(); i++) {
            a32[i] = static_cast<int32_t>(1000.0 * std::sin(0.01 * i)) + static_cast<int32_t>(i % 5);
            a64[i] = static_cast<int64_t>(a32[i]) * 1000003;
        }
        size_t s32 = 0, s64 = 0;
        char *k32 = SZ_compress<int32_t>(ci, a32.data(), s32);
        ci.absErrorBound = 2.0e6;
        char *k64 = SZ_compress<int64_t>(ci, a64.data(), s64);
        SZ3::Config d32, d64;
        int32_t *r32 = SZ_decompress<int32_t>(d32, k32, s32);
        int64_t *r64 = SZ_decompress<int64_t>(d64, k64, s64);
        double w32 = 0, w64 = 0;
        for (size_t i = 0; i < a32.size(); i++) {
            w32 = std::fmax(w32, std::fabs((double)r32[i] - (double)a32[i]));
            w64 = std::fmax(w64, std::fabs((double)r64[i] - (double)a64[i]));
        }
        delete[] r32;
        delete[] r64;
        delete[] k32;
        delete[] k64;
        if (!(w32 <= 2.0) || !(w64 <= 2.0e6) || s32 >= a32.size() * 4 || s64 >= a64.size() * 8) return 26;
    }
    printf("roundtrip ok ratio %.2f, openmp container: %d slabs on %d GPUs\n", data.size() * 4.0 / cmpSize, nslab, sz3b_device_count());
    return 0;
}
'''


def _build(tmp_path):
    src = tmp_path / "dropin.cpp"
    src.write_text(SRC)
    exe = tmp_path / "dropin"
    lib = os.path.join(ROOT, "sz3_b200", "lib")
    subprocess.run(["g++", "-std=c++17", "-O1", str(src), "-I", os.path.join(ROOT, "include"), "-L", lib, "-lSZ3c", "-lsz3b200",
                    f"-Wl,-rpath,{lib}", "-o", str(exe)], check=True)
    return str(exe)


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "sz3_b200", "lib", "libSZ3c.so")), reason="libraries not built")
def test_dropin_headers_compile_and_config_roundtrip(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)


@pytest.mark.gpu
def test_dropin_roundtrip_on_gpu(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe, "gpu"], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "roundtrip ok" in r.stdout
