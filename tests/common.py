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
    return {np.dtype(np.float32): 0, np.dtype(np.float64): 1, np.dtype(np.int32): 7, np.dtype(np.int64): 9}[a.dtype]


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
    deps = [src] + [os.path.join(ROOT, "sz3_b200", "csrc", f) for f in ("core.cuh", "interp_body.cuh", "interp_line.cuh", "interp_lean.cuh", "interp_box.cuh", "interp_plan.hpp", "blockwise.cuh", "lorenzo.cuh")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.run(["g++", "-O2", "-std=c++20", "-fPIC", "-ffp-contract=off", "-shared", "-pthread", src, "-o", out],
                       check=True)
    return C.CDLL(out)


def huffman_host_emul_lib():
    """The product's host Huffman stage (tree, blob, code table) with sequential stand-ins for its two GPU kernels
    (tests/emul/huffman_host_emul.cpp; test infrastructure)."""
    src = os.path.join(ROOT, "tests", "emul", "huffman_host_emul.cpp")
    out = os.path.join(ROOT, "tests", "emul", "_build", "libhuffman_host_emul.so")
    deps = [src] + [os.path.join(ROOT, "sz3_b200", "csrc", f) for f in ("huffman_host.cpp", "huffman_host.hpp")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", src, "-o", out], check=True)
    lib = C.CDLL(out)
    lib.emul_huffman_encode.restype = C.c_longlong
    lib.emul_huffman_encode.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p]
    return lib


def zhuf_emul_lib():
    """Sequential encoder of the GPU lossless stage's frames (tests/emul/zhuf_emul.cpp; test infrastructure)."""
    src = os.path.join(ROOT, "tests", "emul", "zhuf_emul.cpp")
    out = os.path.join(ROOT, "tests", "emul", "_build", "libzhuf_emul.so")
    deps = [src, os.path.join(ROOT, "sz3_b200", "csrc", "zhuf.cuh"), os.path.join(ROOT, "sz3_b200", "csrc", "core.cuh"),
            os.path.join(ROOT, "sz3_b200", "csrc", "zhuf_dec.cuh")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", src, "-o", out], check=True)
    lib = C.CDLL(out)
    lib.zhuf_emul_compress.restype = C.c_longlong
    lib.zhuf_emul_compress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p]
    lib.zhuf_emul_decompress.restype = C.c_longlong
    lib.zhuf_emul_decompress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int]
    return lib


def zstd_lib():
    """The image's libzstd runtime: its decoder is the checker of every zstd frame this repo writes."""
    z = C.CDLL("libzstd.so.1")
    z.ZSTD_decompress.restype = C.c_size_t
    z.ZSTD_decompress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
    z.ZSTD_compress.restype = C.c_size_t
    z.ZSTD_compress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int]
    z.ZSTD_compressBound.restype = C.c_size_t
    z.ZSTD_compressBound.argtypes = [C.c_size_t]
    z.ZSTD_isError.argtypes = [C.c_size_t]
    z.ZSTD_getErrorName.restype = C.c_char_p
    z.ZSTD_getErrorName.argtypes = [C.c_size_t]
    return z


def zhuf_cases():
    """Byte buffers for the lossless-stage tests: (name, uint8 array)."""
    rng = np.random.default_rng(1)
    cases = []
    # a Huffman-coded index stream like the one the stage sees: geometric symbols through a canonical-ish bit packing
    sym = np.minimum(rng.geometric(0.35, size=3_000_000), 15).astype(np.uint8)
    bits = np.unpackbits(((1 << sym.astype(np.uint16)) - 2).astype(">u2").view(np.uint8).reshape(-1, 2), axis=1)
    keep = np.arange(16)[None, :] >= (16 - sym[:, None])
    cases.append(("huffman-coded stream", np.packbits(bits[keep])))
    p = np.exp(-np.arange(256) / 20.0)
    cases.append(("geometric over 256 symbols", rng.choice(256, size=2_500_000, p=p / p.sum()).astype(np.uint8)))
    p = np.exp(-np.arange(256) / 4.0)
    cases.append(("steep (length-limited code)", rng.choice(256, size=1_200_000, p=p / p.sum()).astype(np.uint8)))
    cases.append(("40 symbols uniform", rng.choice(40, size=900_000).astype(np.uint8)))
    cases.append(("3 symbols", rng.choice(3, size=600_000, p=[0.9, 0.09, 0.01]).astype(np.uint8)))
    cases.append(("two symbols", rng.choice([7, 200], size=300_000).astype(np.uint8)))
    cases.append(("uniform random (raw blocks)", rng.integers(0, 256, size=1_500_000).astype(np.uint8)))
    cases.append(("all zero (raw blocks)", np.zeros(400_000, np.uint8)))
    cases.append(("mixed", np.concatenate([rng.choice(256, size=1 << 20, p=p / p.sum()), rng.integers(0, 256, size=1 << 20)]).astype(np.uint8)))
    cases.append(("one block + short tail", rng.choice(256, size=131072 + 500, p=p / p.sum()).astype(np.uint8)))
    cases.append(("5000 bytes", rng.choice(30, size=5000).astype(np.uint8)))
    cases.append(("100 bytes", rng.choice(30, size=100).astype(np.uint8)))
    cases.append(("1 byte", np.array([42], np.uint8)))
    for k in range(12):
        n = int(rng.integers(1, 400000))
        a = float(rng.uniform(0.5, 60))
        m = int(rng.integers(2, 257))
        pp = np.exp(-np.arange(m) / a)
        cases.append((f"random alphabet m={m} a={a:.1f} n={n}", rng.permutation(256)[:m][rng.choice(m, size=n, p=pp / pp.sum())].astype(np.uint8)))
    return cases


def product_lib():
    # (SZ3B_LIB_PATH: a differently built copy of the library, for A/B measurements of kernel variants)
    lib = _load(os.environ.get("SZ3B_LIB_PATH") or os.path.join(ROOT, "sz3_b200", "lib", "libsz3b200.so"))
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
