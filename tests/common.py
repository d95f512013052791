"""Shared test plumbing: ctypes views of the three libraries and the seeded synthetic fields of SURVEY.md section 8(d).

  REF   oracle/_ref/libsz3ref.so   the unmodified reference compiled by oracle/Makefile (checker only)
  EMUL  tests/emul/_build/libemul.so  the kernel bodies run with host threads (indexing check without a GPU)
  LIB   sz3_b200/lib/libsz3b200.so  the product (CUDA); compute entry points need a GPU
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class Config(C.Structure):
    """POD mirror of SZ3::Config; layout of sz3b_config (include/sz3b.h). ref_config/orc_config are its prefix."""
    _fields_ = [
        ("N", C.c_int32), ("dims", C.c_uint64 * 4), ("cmprAlgo", C.c_int32), ("errorBoundMode", C.c_int32),
        ("absErrorBound", C.c_double), ("relErrorBound", C.c_double), ("psnrErrorBound", C.c_double),
        ("l2normErrorBound", C.c_double), ("openmp", C.c_int32), ("quantbinCnt", C.c_int32),
        ("blockSize", C.c_int32), ("lorenzo", C.c_int32), ("lorenzo2", C.c_int32), ("regression", C.c_int32),
        ("regression2", C.c_int32), ("interpAlgo", C.c_int32), ("interpDirection", C.c_int32),
        ("interpAnchorStride", C.c_int32), ("interpAlpha", C.c_double), ("interpBeta", C.c_double),
        ("dataType", C.c_int32), ("predDim", C.c_int32),
    ]


EB_ABS, EB_REL, EB_PSNR, EB_L2NORM, EB_ABS_AND_REL, EB_ABS_OR_REL = range(6)
ALGO_LORENZO_REG, ALGO_INTERP_LORENZO, ALGO_INTERP, ALGO_NOPRED, ALGO_LOSSLESS = range(5)


def make_config(shape, **kw):
    dims = [d for d in shape if d > 1] or [1]
    c = Config()
    c.N = len(dims)
    for i, d in enumerate(dims):
        c.dims[i] = d
    c.cmprAlgo = ALGO_INTERP_LORENZO
    c.errorBoundMode = EB_ABS
    c.absErrorBound = 1e-3
    c.quantbinCnt = 65536
    c.blockSize = {1: 128, 2: 16}.get(c.N, 6)
    c.lorenzo, c.lorenzo2, c.regression, c.regression2 = 1, 0, 1, 0
    c.interpAlgo, c.interpDirection, c.interpAnchorStride = 1, 0, -1
    c.interpAlpha, c.interpBeta = 1.25, 2.0
    c.dataType, c.predDim = 0, c.N
    for k, v in kw.items():
        if not hasattr(c, k):
            raise AttributeError(k)
        setattr(c, k, v)
    return c


def dtype_code(a):
    return {np.dtype(np.float32): 0, np.dtype(np.float64): 1}[a.dtype]


def _load(path):
    return C.CDLL(path) if os.path.exists(path) else None


def ref_lib():
    lib = _load(os.path.join(ROOT, "oracle", "_ref", "libsz3ref.so"))
    if lib is None:
        return None
    lib.ref_compress.restype = C.c_longlong
    lib.ref_size_bound.restype = C.c_size_t
    lib.ref_interp_decompose.restype = C.c_longlong
    lib.ref_blockwise_decompose.restype = C.c_longlong
    lib.ref_huffman_encode.restype = C.c_longlong
    lib.ref_huffman_decode.restype = C.c_longlong
    lib.ref_abs_eb.restype = C.c_double
    lib.ref_config_save.restype = C.c_size_t
    return lib


def port_lib():
    path = os.path.join(ROOT, "oracle", "libsz3oracle.so")
    if not os.path.exists(path) and os.path.exists(os.path.join(ROOT, "oracle", "sz3_oracle.c")):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "port"], check=False, capture_output=True)
    lib = _load(path)
    if lib is None:
        return None
    lib.orc_compress.restype = C.c_longlong
    lib.orc_size_bound.restype = C.c_size_t
    lib.orc_interp_decompose.restype = C.c_longlong
    lib.orc_blockwise_decompose.restype = C.c_longlong
    lib.orc_huffman_encode.restype = C.c_longlong
    lib.orc_huffman_decode.restype = C.c_longlong
    lib.orc_abs_eb.restype = C.c_double
    lib.orc_config_save.restype = C.c_size_t
    return lib


def emul_lib():
    src = os.path.join(ROOT, "tests", "emul", "emul.cpp")
    out = os.path.join(ROOT, "tests", "emul", "_build", "libemul.so")
    deps = [src] + [os.path.join(ROOT, "sz3_b200", "csrc", f) for f in ("core.cuh", "interp_body.cuh", "interp_fast.cuh", "interp_line.cuh", "interp_lean.cuh", "interp_plan.hpp", "blockwise.cuh", "lorenzo.cuh")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.run(["g++", "-O2", "-std=c++20", "-fPIC", "-ffp-contract=off", "-shared", "-pthread", src, "-o", out],
                       check=True)
    return C.CDLL(out)


def product_lib():
    lib = _load(os.path.join(ROOT, "sz3_b200", "lib", "libsz3b200.so"))
    if lib is None:
        return None
    lib.sz3b_last_error.restype = C.c_char_p
    lib.sz3b_version.restype = C.c_char_p
    lib.sz3b_compress_bound.restype = C.c_size_t
    lib.sz3b_config_save.restype = C.c_size_t
    lib.sz3b_omp_header_size.restype = C.c_size_t
    return lib


# ---- synthetic fields --------------------------------------------------------------------------
def field_g3(shape, dtype=np.float32, seed=1234):
    nz, ny, nx = shape
    z = np.arange(nz, dtype=dtype)[:, None, None]
    y = np.arange(ny, dtype=dtype)[None, :, None]
    x = np.arange(nx, dtype=dtype)[None, None, :]
    two_pi = dtype(2 * np.pi)
    a = (np.sin(two_pi * x / dtype(64)) * np.cos(two_pi * y / dtype(96)) + dtype(0.5) * np.sin(two_pi * z / dtype(128) + dtype(0.3))
         + dtype(0.25) * np.sin(two_pi * (x + y + z) / dtype(37)))
    noise = np.random.default_rng(seed).standard_normal(shape, dtype=dtype)
    return np.ascontiguousarray((a + dtype(0.002) * noise).astype(dtype))


def field_g1(n, seed=1234):
    t = np.arange(n, dtype=np.float64)
    a = np.sin(2 * np.pi * t / 4096) + 0.1 * np.sin(2 * np.pi * t / 97)
    return (a + 1e-4 * np.random.default_rng(seed).standard_normal(n)).astype(np.float32)


def field_g4(shape, seed=1234):
    T, Z, Y, X = shape
    t = np.arange(T, dtype=np.float32)[:, None, None, None]
    z = np.arange(Z, dtype=np.float32)[None, :, None, None]
    y = np.arange(Y, dtype=np.float32)[None, None, :, None]
    x = np.arange(X, dtype=np.float32)[None, None, None, :]
    a = (280 + 30 * np.cos(np.pi * y / Y) + 5 * np.sin(2 * np.pi * x / X + 0.2 * t) + 0.1 * z * np.sin(2 * np.pi * y / 32))
    noise = np.random.default_rng(seed).standard_normal(shape, dtype=np.float32)
    return np.ascontiguousarray((a + 0.05 * noise).astype(np.float32))


def field_nd(shape, dtype=np.float32, seed=7):
    """Smooth-ish field of any rank for edge-shape parity cases."""
    rng = np.random.default_rng(seed)
    grids = np.meshgrid(*[np.arange(n, dtype=np.float64) for n in shape], indexing="ij")
    a = np.zeros(shape, dtype=np.float64)
    for k, g in enumerate(grids):
        a += np.sin(2 * np.pi * g / (23.0 + 11 * k)) * (1.0 + 0.3 * k)
    a += 0.01 * rng.standard_normal(shape)
    return np.ascontiguousarray(a.astype(dtype))


# ---- reference calls ---------------------------------------------------------------------------------------------------
def ref_interp(lib, data, conf, eb, prefix="ref"):
    work = data.copy()
    n = data.size
    q = np.empty(n, dtype=np.int32)
    blob = np.empty(n * data.itemsize + 4096, dtype=np.uint8)
    blen = C.c_size_t(0)
    fn = getattr(lib, prefix + "_interp_decompose")
    r = fn(dtype_code(data), C.byref(conf), C.c_double(eb), work.ctypes.data_as(C.c_void_p), q.ctypes.data_as(C.c_void_p),
           blob.ctypes.data_as(C.c_void_p), C.byref(blen))
    assert r == n, r
    return q, bytes(blob[:blen.value]), work


def ref_blockwise(lib, data, conf, eb, prefix="ref"):
    """BlockwiseDecomposition::compress + save of the checker: (indices, blob)."""
    work = data.copy()
    n = data.size
    q = np.empty(n, dtype=np.int32)
    blob = np.empty(2 * n * data.itemsize + (1 << 20), dtype=np.uint8)
    blen = C.c_size_t(0)
    fn = getattr(lib, prefix + "_blockwise_decompose")
    r = fn(dtype_code(data), C.byref(conf), C.c_double(eb), work.ctypes.data_as(C.c_void_p), q.ctypes.data_as(C.c_void_p),
           blob.ctypes.data_as(C.c_void_p), C.byref(blen))
    assert r == n, r
    return q, bytes(blob[:blen.value])


def interp_blob_unpred(blob, N, dtype):
    """Splits InterpolationDecomposition::save output into (header bytes, unpred array)."""
    hdr = 8 * N + 4 + 4 + 4 + 8 + 8 + 8 + 1 + 8 + 4 + 8
    nun = int(np.frombuffer(blob[hdr - 8:hdr], dtype=np.uint64)[0])
    un = np.frombuffer(blob[hdr:hdr + nun * np.dtype(dtype).itemsize], dtype=dtype)
    return blob[:hdr], un
