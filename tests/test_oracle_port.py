"""Pins the plain-C restatement (oracle/sz3_oracle.c + sz3_oracle_t.inc) against the unmodified reference
(oracle/_ref/libsz3ref.so, built from /root/reference in this container): indices, decomposition blobs, Huffman
streams and whole compressed streams byte for byte, plus decode round trips.  CPU only."""
import ctypes as C

import numpy as np
import pytest

from common import (ALGO_INTERP, ALGO_LORENZO_REG, EB_REL, Config, dtype_code, field_g1, field_g3, field_nd, make_config, port_lib, ref_blockwise,
                    ref_interp, ref_lib)

pytestmark = pytest.mark.skipif(ref_lib() is None or port_lib() is None, reason="needs oracle/_ref and oracle/libsz3oracle.so")


@pytest.mark.parametrize("shape,dtype,kw", [
    ((40, 50, 70), np.float32, dict(interpAlgo=1, interpDirection=0)),
    ((33, 65, 97), np.float32, dict(interpAlgo=0, interpDirection=5)),
    ((20, 37, 66), np.float64, dict(interpAlgo=1, interpDirection=3, interpAlpha=2.0, interpBeta=3.0)),
    ((9, 12, 20, 18), np.float32, dict(interpAlgo=1, interpAnchorStride=16)),
    ((100, 333), np.float32, dict(interpAlgo=1, interpAnchorStride=128)),
    ((5000,), np.float32, dict(interpAlgo=0, interpAnchorStride=4096)),
])
def test_port_interp_matches_reference(shape, dtype, kw):
    data = field_nd(shape, dtype)
    kw.setdefault("interpAnchorStride", 32)
    conf = make_config(shape, cmprAlgo=ALGO_INTERP, **kw)
    q_ref, blob_ref, _ = ref_interp(ref_lib(), data, conf, 1e-2)
    q, blob, _ = ref_interp(port_lib(), data, conf, 1e-2, prefix="orc")
    assert np.array_equal(q, q_ref)
    assert blob == blob_ref


@pytest.mark.parametrize("shape,dtype,kw", [
    ((24, 30, 36), np.float64, dict(lorenzo=0, regression=1)),
    ((20, 33, 47), np.float32, dict(lorenzo=0, regression=1)),
    ((24, 30, 36), np.float32, dict(lorenzo=1, regression=0)),
    ((20, 33, 47), np.float32, dict(lorenzo=1, regression=1)),
    ((30, 31, 32), np.float64, dict(lorenzo=1, lorenzo2=1, regression=1)),
    ((40, 45), np.float32, dict(lorenzo=1, regression=1, blockSize=16)),
    ((3000,), np.float32, dict(lorenzo=1, lorenzo2=1, regression=0, blockSize=128)),
])
def test_port_blockwise_matches_reference(shape, dtype, kw):
    data = field_nd(shape, dtype)
    conf = make_config(shape, cmprAlgo=ALGO_LORENZO_REG, **kw)
    q_ref, blob_ref = ref_blockwise(ref_lib(), data, conf, 1e-3)
    q, blob = ref_blockwise(port_lib(), data, conf, 1e-3, prefix="orc")
    assert np.array_equal(q, q_ref)
    assert blob == blob_ref


def test_port_huffman_matches_reference():
    rng = np.random.default_rng(5)
    for q in (rng.integers(32700, 32830, 20000).astype(np.int32), np.full(1000, 7, np.int32),
              (32768 + np.round(rng.standard_normal(50000) * 3)).astype(np.int32), rng.integers(0, 65536, 30000).astype(np.int32)):
        outs = []
        for lib, pre in ((ref_lib(), "ref"), (port_lib(), "orc")):
            buf = np.empty(q.size * 8 + (1 << 20), np.uint8)
            tl = C.c_size_t(0)
            n = getattr(lib, pre + "_huffman_encode")(q.ctypes.data_as(C.c_void_p), C.c_size_t(q.size), buf.ctypes.data_as(C.c_void_p), C.byref(tl))
            assert n > 0
            outs.append((bytes(buf[:n]), tl.value))
        assert outs[0] == outs[1]
        back = np.empty_like(q)
        enc = np.frombuffer(outs[1][0], np.uint8)
        m = port_lib().orc_huffman_decode(enc.ctypes.data_as(C.c_void_p), C.c_size_t(enc.size), C.c_size_t(q.size), back.ctypes.data_as(C.c_void_p))
        assert m == enc.size and np.array_equal(back, q)


@pytest.mark.parametrize("shape,dtype,kw", [
    ((40, 50, 70), np.float32, dict(cmprAlgo=ALGO_INTERP, absErrorBound=1e-3)),
    ((30, 36, 42), np.float64, dict(cmprAlgo=ALGO_LORENZO_REG, errorBoundMode=EB_REL, relErrorBound=1e-4)),
    ((30, 36, 42), np.float32, dict(cmprAlgo=ALGO_LORENZO_REG, lorenzo=0, absErrorBound=1e-3)),
    ((64, 64), np.float32, dict(cmprAlgo=ALGO_INTERP, absErrorBound=0.0)),
])
def test_port_streams_match_reference_and_decode(shape, dtype, kw):
    data = field_nd(shape, dtype)
    conf = make_config(shape, **kw)
    R, P = ref_lib(), port_lib()
    cap = R.ref_size_bound(dtype_code(data), C.byref(conf)) + 8192
    a, b = np.empty(cap, np.uint8), np.empty(cap, np.uint8)
    na = R.ref_compress(dtype_code(data), C.byref(conf), data.ctypes.data_as(C.c_void_p), a.ctypes.data_as(C.c_char_p), C.c_size_t(cap))
    nb = P.orc_compress(dtype_code(data), C.byref(conf), data.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_char_p), C.c_size_t(cap))
    assert na > 0 and na == nb and np.array_equal(a[:na], b[:nb])
    dec_r, dec_p = np.empty_like(data), np.empty_like(data)
    cr, cp = Config(), Config()
    assert R.ref_decompress(dtype_code(data), a.ctypes.data_as(C.c_char_p), C.c_size_t(na), dec_r.ctypes.data_as(C.c_void_p), C.byref(cr)) == 0
    assert P.orc_decompress(dtype_code(data), a.ctypes.data_as(C.c_char_p), C.c_size_t(na), dec_p.ctypes.data_as(C.c_void_p), C.byref(cp)) == 0
    assert np.array_equal(dec_r.view(np.uint8), dec_p.view(np.uint8))


def _tuner_inputs():
    rng = np.random.default_rng(5)
    n = np.arange(1 << 16)
    return [
        (field_nd((8, 8, 128), np.float32), dict(absErrorBound=1.0)),                         # too small to tune
        (field_nd((100, 70, 130), np.float32), dict(absErrorBound=1e-3)),                     # (alpha, beta) = (1, 1) wins
        (field_nd((48, 60, 72), np.float64), dict(absErrorBound=1e-5)),
        (field_nd((12, 40, 40, 40), np.float32), dict(absErrorBound=1e-2)),                   # 4-D, 16-cubes
        (field_nd((300, 500), np.float32), dict(absErrorBound=1e-3)),
        (field_g1(1 << 18), dict(absErrorBound=1e-4)),                                         # 1-D: the Lorenzo stack wins
        ((np.sin(n / 50.0) + 0.05 * rng.standard_normal(n.size)).astype(np.float32), dict(absErrorBound=1e-3)),   # linear wins
        (field_g1(100000), dict(errorBoundMode=EB_REL, relErrorBound=5e-7)),                   # tight relative bound
        (field_g3((64, 64, 64)), dict(errorBoundMode=EB_REL, relErrorBound=1e-7)),             # tuned stream overflows the buffer -> lossless
        (rng.standard_normal((40, 40, 40)).astype(np.float32), dict(absErrorBound=1e-3)),      # noise
    ]


@pytest.mark.parametrize("k", range(10))
def test_port_tuner_matches_reference(k):
    """ALGO_INTERP_LORENZO: profiling, sampling, the trial compressions and the decision sequence of
    SZ_compress_Interp_lorenzo, restated in oracle/sz3_oracle_t.inc -- the stream (with the tuned Config at its end) is
    the reference's, byte for byte, at the same buffer capacity (a length_error of the tuned run means lossless)."""
    from common import ALGO_INTERP_LORENZO
    data, kw = _tuner_inputs()[k]
    conf = make_config(data.shape, cmprAlgo=ALGO_INTERP_LORENZO, **kw)
    R, P = ref_lib(), port_lib()
    cap = R.ref_size_bound(dtype_code(data), C.byref(conf)) + 8192
    a, b = np.empty(cap, np.uint8), np.empty(cap, np.uint8)
    na = R.ref_compress(dtype_code(data), C.byref(conf), data.ctypes.data_as(C.c_void_p), a.ctypes.data_as(C.c_char_p), C.c_size_t(cap))
    nb = P.orc_compress(dtype_code(data), C.byref(conf), data.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_char_p), C.c_size_t(cap))
    assert na > 0 and na == nb and np.array_equal(a[:na], b[:nb])


@pytest.mark.parametrize("shape,dtype,kw", [
    ((100, 64, 64), np.float32, dict(cmprAlgo=1, absErrorBound=1e-3)),                                  # tuned slabs
    ((37, 40, 50), np.float64, dict(cmprAlgo=ALGO_LORENZO_REG, errorBoundMode=EB_REL, relErrorBound=1e-4)),   # global range
    ((64, 300), np.float32, dict(cmprAlgo=ALGO_INTERP, errorBoundMode=2, psnrErrorBound=70.0)),
    ((5, 40, 50), np.float32, dict(cmprAlgo=1, absErrorBound=1e-3)),      # fewer rows than threads: one-row slabs lose a dimension
    ((200000,), np.float32, dict(cmprAlgo=1, absErrorBound=1e-4)),        # 1-D slabs: some go to the Lorenzo stack
])
def test_port_openmp_container_matches_reference(shape, dtype, kw):
    """SZ_compress_OMP / SZ_decompress_OMP (api/impl/SZImplOMP.hpp): the restatement writes the reference's container
    byte for byte when told the reference's thread count, and decodes it to the same bits."""
    data = field_g1(shape[0]) if len(shape) == 1 else field_nd(shape, dtype)
    R, P = ref_lib(), port_lib()
    conf = make_config(shape, openmp=1, **kw)
    cap = R.ref_size_bound(dtype_code(data), C.byref(conf)) + 8192
    a, b = np.empty(cap, np.uint8), np.empty(cap, np.uint8)
    na = R.ref_compress(dtype_code(data), C.byref(conf), data.ctypes.data_as(C.c_void_p), a.ctypes.data_as(C.c_char_p), C.c_size_t(cap))
    # the slab count the reference just used (it lowers the OpenMP thread count for good when rows < threads)
    threads = C.CDLL("libgomp.so.1").omp_get_max_threads()
    pconf = make_config(shape, openmp=min(threads, shape[0]), **kw)
    nb = P.orc_compress(dtype_code(data), C.byref(pconf), data.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_char_p), C.c_size_t(cap))
    assert na > 0 and na == nb and np.array_equal(a[:na], b[:nb])
    dec_r, dec_p = np.empty_like(data), np.empty_like(data)
    cr, cp = Config(), Config()
    assert R.ref_decompress(dtype_code(data), a.ctypes.data_as(C.c_char_p), C.c_size_t(na), dec_r.ctypes.data_as(C.c_void_p), C.byref(cr)) == 0
    assert P.orc_decompress(dtype_code(data), a.ctypes.data_as(C.c_char_p), C.c_size_t(na), dec_p.ctypes.data_as(C.c_void_p), C.byref(cp)) == 0
    assert np.array_equal(dec_r.view(np.uint8), dec_p.view(np.uint8))
