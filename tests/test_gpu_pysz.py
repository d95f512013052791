"""The reference's own Python binding (tools/pysz: sz.pyx / sz.pxd, UNMODIFIED) built against this repo's drop-in
headers and libsz3b200 (tools/build_pysz.py -> build/pysz): what a pysz user calls, with the GPU path behind it.
Same calls as the reference's tools/pysz/tests/test_pysz.py; on top of its checks, the bytes pysz returns are compared
with the unmodified reference's SZ_compress for the same Config, for every element type pysz accepts."""
import os
import sys

import numpy as np
import pytest

from common import ROOT, EB_ABS, EB_REL, field_nd, make_config, ref_lib

PYSZ = os.path.join(ROOT, "build", "pysz")
built = any(f.startswith("sz.") and f.endswith(".so") for f in os.listdir(os.path.join(PYSZ, "pysz"))) if os.path.isdir(os.path.join(PYSZ, "pysz")) else False
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not built, reason="build/pysz not built (tools/build_pysz.py needs /root/reference)")]


@pytest.fixture(scope="module")
def pysz():
    sys.path.insert(0, PYSZ)
    import pysz as m
    return m


def test_reference_test_script_calls(pysz):
    sz, szConfig, mode = pysz.sz, pysz.szConfig, pysz.szErrorBoundMode
    rng = np.random.default_rng(0)
    data = rng.standard_normal((100, 100)).astype(np.float32)
    config = szConfig()
    config.errorBoundMode = mode.ABS
    config.absErrorBound = 0.01
    compressed, ratio = sz.compress(data, config)
    assert compressed.dtype == np.uint8 and ratio == data.nbytes / compressed.size
    decompressed, dec_config = sz.decompress(compressed, data.dtype, data.shape)
    assert decompressed.shape == data.shape and dec_config.dims == data.shape
    max_error, psnr, nrmse = sz.verify(data, decompressed)
    assert max_error <= 0.01
    d64 = rng.standard_normal((50, 50))
    c64 = szConfig()
    c64.errorBoundMode = mode.ABS
    c64.absErrorBound = 1e-6
    cmp64, _ = sz.compress(d64, c64)
    dec64, _ = sz.decompress(cmp64, d64.dtype, d64.shape)
    assert sz.verify(d64, dec64)[0] <= 1e-6
    d3 = rng.standard_normal((20, 30, 40)).astype(np.float32)
    c3 = szConfig()
    c3.errorBoundMode = mode.REL
    c3.relErrorBound = 0.001
    cmp3, _ = sz.compress(d3, c3)
    dec3, _ = sz.decompress(cmp3, d3.dtype, d3.shape)
    assert sz.verify(d3, dec3)[0] <= 0.001 * float(d3.max() - d3.min()) * (1 + 1e-6)


@pytest.mark.skipif(ref_lib() is None, reason="oracle/_ref/libsz3ref.so not built")
@pytest.mark.parametrize("dtype,scale,eb", [(np.float32, 1.0, 1e-3), (np.float64, 1.0, 1e-6), (np.int32, 2000.0, 2.0), (np.int64, 1.0e9, 500.0)])
def test_pysz_bytes_match_reference(pysz, dtype, scale, eb):
    from test_gpu_compress import ref_compress, ref_decompress
    g = field_nd((48, 60, 72), np.float64)
    data = np.ascontiguousarray(np.rint(g * scale).astype(dtype) if np.issubdtype(dtype, np.integer) else g.astype(dtype))
    config = pysz.szConfig()
    config.errorBoundMode = pysz.szErrorBoundMode.ABS
    config.absErrorBound = eb
    compressed, _ = pysz.sz.compress(data, config)
    # the reference with the capacity pysz offers (2 x the array, sz.pyx:230): the capacity decides whether a nearly
    # incompressible array stays on the lossy path or falls back to the lossless one (SZDispatcher.hpp:55-61)
    import ctypes as C
    from common import dtype_code
    R = ref_lib()
    rconf = make_config(data.shape, errorBoundMode=EB_ABS, absErrorBound=eb)
    buf = np.empty(data.nbytes * 2, dtype=np.uint8)
    R.ref_compress.restype = C.c_longlong
    n = R.ref_compress(dtype_code(data), C.byref(rconf), data.ctypes.data_as(C.c_void_p), buf.ctypes.data_as(C.c_char_p), C.c_size_t(buf.size))
    assert n > 0
    theirs = buf[:n].copy()
    assert compressed.size == theirs.size and np.array_equal(compressed, theirs)
    dec, conf = pysz.sz.decompress(theirs, dtype, data.shape)
    dec_ref, conf_ref = ref_decompress(theirs, data)
    assert np.array_equal(dec, dec_ref)
    assert conf.absErrorBound == eb and conf.cmprAlgo == conf_ref.cmprAlgo   # (the Config the stream carries)


def test_pysz_rejects_what_the_reference_rejects(pysz):
    with pytest.raises(TypeError):
        pysz.sz.compress(np.zeros((8, 8), dtype=np.int16), pysz.szConfig())
    with pytest.raises(ValueError):   # std::invalid_argument: not an SZ3 stream
        pysz.sz.decompress(np.zeros(64, dtype=np.uint8), np.float32, (4, 4))
