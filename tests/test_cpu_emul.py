"""CPU-only check of the kernel bodies' traversal/indexing logic: sz3_b200/csrc/interp_body.cuh compiled with g++ and
driven by host threads (tests/emul) must reproduce the reference's quantization indices, order and unpredictables.
This is test infrastructure; the product only instantiates those bodies inside __global__ kernels."""
import ctypes as C

import numpy as np
import pytest

from common import (ALGO_INTERP, ALGO_LORENZO_REG, dtype_code, emul_lib, field_g3, field_nd, interp_blob_unpred, make_config,
                    ref_blockwise, ref_interp, ref_lib)

pytestmark = pytest.mark.skipif(ref_lib() is None, reason="oracle/_ref not built")


def emul(data, conf, eb, schedule, nthreads=4, hist=None):
    E = emul_lib()
    q = np.empty(data.size, np.int32)
    un = np.empty(data.size, data.dtype)
    nun = C.c_size_t(0)
    rc = E.emul_interp_decompose(dtype_code(data), C.byref(conf), C.c_double(eb), data.ctypes.data_as(C.c_void_p), schedule,
                                 nthreads, q.ctypes.data_as(C.c_void_p), un.ctypes.data_as(C.c_void_p), C.byref(nun),
                                 None if hist is None else hist.ctypes.data_as(C.c_void_p))
    assert rc == 0
    return q, un[:nun.value]


@pytest.mark.parametrize("schedule", [1, 4])
@pytest.mark.parametrize("shape,dtype,kw", [
    ((40, 50, 70), np.float32, dict(interpAlgo=1, interpDirection=0)),
    ((33, 65, 97), np.float32, dict(interpAlgo=0, interpDirection=5)),
    ((20, 37, 66), np.float64, dict(interpAlgo=1, interpDirection=3, interpAlpha=2.0, interpBeta=3.0)),
    ((8, 8, 128), np.float32, dict(interpAlgo=1, interpDirection=0)),
])
def test_emul_matches_reference_3d(shape, dtype, kw, schedule):
    data = field_nd(shape, dtype)
    conf = make_config(shape, cmprAlgo=ALGO_INTERP, interpAnchorStride=32, **kw)
    q_ref, blob_ref, _ = ref_interp(ref_lib(), data, conf, 1e-2)
    q, un = emul(data, conf, 1e-2, schedule, nthreads=32 if schedule == 4 else 4)
    assert np.array_equal(q, q_ref)
    _, un_ref = interp_blob_unpred(blob_ref, conf.N, dtype)
    assert np.array_equal(un, un_ref)


@pytest.mark.parametrize("shape", [(9, 12, 20, 18), (100, 333), (5000,)])
def test_emul_matches_reference_other_ranks(shape):
    data = field_nd(shape, np.float32)
    conf = make_config(shape, cmprAlgo=ALGO_INTERP, interpAnchorStride=[4096, 128, 32, 16][len(shape) - 1])
    q_ref, blob_ref, _ = ref_interp(ref_lib(), data, conf, 1e-3)
    q, un = emul(data, conf, 1e-3, 0)
    assert np.array_equal(q, q_ref)


@pytest.mark.parametrize("shape,dtype,kw", [
    ((64, 64, 64), np.float32, dict(interpAlgo=1, interpDirection=5)),
    ((64, 64, 64), np.float32, dict(interpAlgo=0, interpDirection=0)),
    ((100, 70, 130), np.float32, dict(interpAlgo=1, interpDirection=0)),
    ((66, 35, 34), np.float32, dict(interpAlgo=0, interpDirection=5)),
    ((2, 3, 200), np.float32, dict(interpAlgo=1, interpDirection=2)),
    ((37, 41, 130), np.float64, dict(interpAlgo=0, interpDirection=1)),
    ((45, 33, 36), np.float32, dict(interpAlgo=1, interpDirection=4)),
])
def test_emul_line_walker_shapes(shape, dtype, kw):
    """Edge tiles (n = 32 / odd extents / degenerate dims), both interpolators, the non-tuner directions."""
    data = field_nd(shape, dtype)
    conf = make_config(shape, cmprAlgo=ALGO_INTERP, interpAnchorStride=32, **kw)
    q_ref, blob_ref, _ = ref_interp(ref_lib(), data, conf, 1e-2)
    for nthreads in (16, 48):
        q, un = emul(data, conf, 1e-2, 4, nthreads=nthreads)
        assert np.array_equal(q, q_ref)
        _, un_ref = interp_blob_unpred(blob_ref, conf.N, dtype)
        assert np.array_equal(un, un_ref)


@pytest.mark.parametrize("shape,kw", [
    ((64, 64, 64), dict(interpAlgo=0, interpDirection=5)),
    ((64, 64, 64), dict(interpAlgo=1, interpDirection=0)),
    ((40, 70, 66), dict(interpAlgo=1, interpDirection=5)),
])
def test_emul_line_walker_full_cta(shape, kw):
    """512 emulated threads, as on the device: exercises the constant-increment row walk (needs 32 row slots)."""
    data = field_nd(shape, np.float32)
    conf = make_config(shape, cmprAlgo=ALGO_INTERP, interpAnchorStride=32, **kw)
    q_ref, blob_ref, _ = ref_interp(ref_lib(), data, conf, 1e-2)
    q, un = emul(data, conf, 1e-2, 4, nthreads=512)
    assert np.array_equal(q, q_ref)
    _, un_ref = interp_blob_unpred(blob_ref, conf.N, np.float32)
    assert np.array_equal(un, un_ref)


@pytest.mark.parametrize("shape,dtype,eb,bsz", [
    ((24, 30, 36), np.float64, 1e-3, 6),
    ((24, 30, 36), np.float32, 1e-3, 6),
    ((20, 33, 47), np.float32, 1e-2, 6),      # clipped blocks (extents 2, 3, 5)
    ((40, 45), np.float32, 1e-3, 16),
    ((8, 10, 12, 14), np.float64, 1e-3, 6),
    ((300,), np.float32, 1e-4, 128),
])
def test_emul_regression_matches_reference(shape, dtype, eb, bsz):
    """Regression-only BlockwiseDecomposition: fit, coefficient chain and predict+quantize bodies (blockwise.cuh)."""
    data = field_nd(shape, dtype)
    conf = make_config(shape, cmprAlgo=ALGO_LORENZO_REG, lorenzo=0, lorenzo2=0, regression=1, blockSize=bsz)
    q_ref, blob_ref = ref_blockwise(ref_lib(), data, conf, eb)
    E = emul_lib()
    q = np.empty(data.size, np.int32)
    cq = np.empty(data.size * 2 + 64, np.int32)
    un = np.empty(data.size, dtype)
    ncoef, nun = C.c_size_t(0), C.c_size_t(0)
    rc = E.emul_regression_decompose(dtype_code(data), C.byref(conf), C.c_double(eb), data.ctypes.data_as(C.c_void_p),
                                     q.ctypes.data_as(C.c_void_p), cq.ctypes.data_as(C.c_void_p), C.byref(ncoef),
                                     un.ctypes.data_as(C.c_void_p), C.byref(nun))
    assert rc == 0
    assert np.array_equal(q, q_ref), f"{int((q != q_ref).sum())} of {q.size} indices differ"
    assert int(np.frombuffer(blob_ref[:8], np.uint64)[0]) == ncoef.value
    if nun.value:
        tail = np.frombuffer(blob_ref[len(blob_ref) - nun.value * data.itemsize:], dtype)
        assert np.array_equal(tail, un[:nun.value])


@pytest.mark.parametrize("shape,dtype,kw", [
    ((40, 50, 70), np.float32, dict(interpAlgo=1, interpDirection=0)),
    ((33, 65, 97), np.float32, dict(interpAlgo=0, interpDirection=5)),
    ((20, 37, 66), np.float64, dict(interpAlgo=1, interpDirection=3, interpAlpha=2.0, interpBeta=3.0)),
    ((64, 64, 64), np.float32, dict(interpAlgo=0, interpDirection=2)),
    ((66, 35, 34), np.float32, dict(interpAlgo=0, interpDirection=5)),
    ((2, 3, 200), np.float32, dict(interpAlgo=1, interpDirection=4)),
    ((9, 12, 20, 18), np.float32, dict(interpAlgo=1, interpDirection=0, interpAnchorStride=16)),
    ((7, 18, 34, 21), np.float32, dict(interpAlgo=0, interpDirection=23, interpAnchorStride=16)),
    ((10, 17, 19, 36), np.float64, dict(interpAlgo=1, interpDirection=9, interpAnchorStride=8)),
])
def test_emul_lean_per_pass_matches_reference(shape, dtype, kw):
    """Row-mapped per-pass schedule (interp_lean.cuh): block tables, row descriptors, traversal positions, N = 3 and 4."""
    data = field_nd(shape, dtype)
    kw.setdefault("interpAnchorStride", 32)
    conf = make_config(shape, cmprAlgo=ALGO_INTERP, **kw)
    q_ref, blob_ref, _ = ref_interp(ref_lib(), data, conf, 1e-2)
    q, un = emul(data, conf, 1e-2, 5, nthreads=32)
    assert np.array_equal(q, q_ref), f"{int((q != q_ref).sum())} of {q.size} indices differ"
    _, un_ref = interp_blob_unpred(blob_ref, conf.N, dtype)
    assert np.array_equal(un, un_ref)


@pytest.mark.parametrize("shape,dtype,eb,kw", [
    ((24, 30, 36), np.float32, 1e-3, dict(lorenzo=1, regression=0)),
    ((20, 33, 47), np.float32, 1e-3, dict(lorenzo=1, regression=1)),
    ((30, 31, 32), np.float64, 1e-4, dict(lorenzo=1, lorenzo2=1, regression=1)),
    ((30, 31, 32), np.float32, 1e-2, dict(lorenzo=0, lorenzo2=1, regression=1)),
    ((40, 45), np.float32, 1e-3, dict(lorenzo=1, regression=1, blockSize=16)),
    ((3000,), np.float32, 1e-3, dict(lorenzo=1, lorenzo2=1, regression=0, blockSize=128)),
    ((9, 12, 13, 7), np.float64, 1e-3, dict(lorenzo=1, lorenzo2=1, regression=1, blockSize=4)),
    ((61, 67, 73), np.float32, 1e-4, dict(lorenzo=1, lorenzo2=1, regression=1)),
    ((25, 31, 37), np.float32, 1e-3, dict(lorenzo=1, regression=1, quantbinCnt=16)),
])
@pytest.mark.parametrize("minwin,walkbelow", [(4096, 24), (40, 0), (16, 1000000)])
def test_emul_lorenzo_stacks_match_reference(shape, dtype, eb, kw, minwin, walkbelow, monkeypatch):
    """Block wavefront + windowed selection iteration of lorenzo.cuh / pipeline.cu (default window; 40-block windows
    without the walk; the row-major walk for every stretch), against the reference's sequential walk."""
    monkeypatch.setenv("EMUL_MINWIN", str(minwin))
    monkeypatch.setenv("EMUL_WALKBELOW", str(walkbelow))
    monkeypatch.setenv("EMUL_WALKLEN", "64")
    data = field_nd(shape, dtype)
    conf = make_config(shape, cmprAlgo=ALGO_LORENZO_REG, **kw)
    q_ref, blob_ref = ref_blockwise(ref_lib(), data, conf, eb)
    E = emul_lib()
    q = np.empty(data.size, np.int32)
    sel = np.empty(data.size, np.uint8)
    cq = np.empty(data.size * 2 + 64, np.int32)
    un = np.empty(data.size, dtype)
    dec = np.empty(data.size, dtype)
    ncoef, nun = C.c_size_t(0), C.c_size_t(0)
    rc = E.emul_lorenzo_decompose(dtype_code(data), C.byref(conf), C.c_double(eb), data.ctypes.data_as(C.c_void_p),
                                  q.ctypes.data_as(C.c_void_p), sel.ctypes.data_as(C.c_void_p), cq.ctypes.data_as(C.c_void_p),
                                  C.byref(ncoef), un.ctypes.data_as(C.c_void_p), C.byref(nun), dec.ctypes.data_as(C.c_void_p))
    assert rc > 0
    assert np.array_equal(q, q_ref), f"{int((q != q_ref).sum())} of {q.size} indices differ"
    if conf.regression and (conf.lorenzo or conf.lorenzo2):
        assert int(np.frombuffer(blob_ref[:8], np.uint64)[0]) == ncoef.value
    if nun.value:
        tail = np.frombuffer(blob_ref[len(blob_ref) - nun.value * data.itemsize:], dtype)
        assert np.array_equal(tail, un[:nun.value])
    assert np.max(np.abs(dec.reshape(shape).astype(np.float64) - data)) <= eb


@pytest.mark.parametrize("minwin,walkbelow", [(4096, 24), (256, 100)])
def test_emul_lorenzo_noisy_field(minwin, walkbelow, monkeypatch):
    """Regression-heavy selection (G3 noise, eb 1e-2): many invalidated guesses, windows and walks interleaved."""
    monkeypatch.setenv("EMUL_MINWIN", str(minwin))
    monkeypatch.setenv("EMUL_WALKBELOW", str(walkbelow))
    monkeypatch.setenv("EMUL_WALKLEN", "64")
    data = field_g3((100, 100, 100))
    conf = make_config(data.shape, cmprAlgo=ALGO_LORENZO_REG)
    q_ref, blob_ref = ref_blockwise(ref_lib(), data, conf, 1e-2)
    E = emul_lib()
    q = np.empty(data.size, np.int32)
    sel = np.empty(data.size, np.uint8)
    cq = np.empty(data.size * 2 + 64, np.int32)
    un = np.empty(data.size, np.float32)
    ncoef, nun = C.c_size_t(0), C.c_size_t(0)
    rc = E.emul_lorenzo_decompose(0, C.byref(conf), C.c_double(1e-2), data.ctypes.data_as(C.c_void_p),
                                  q.ctypes.data_as(C.c_void_p), sel.ctypes.data_as(C.c_void_p), cq.ctypes.data_as(C.c_void_p),
                                  C.byref(ncoef), un.ctypes.data_as(C.c_void_p), C.byref(nun), None)
    assert rc > 1, rc
    assert np.array_equal(q, q_ref)
    assert int(np.frombuffer(blob_ref[:8], np.uint64)[0]) == ncoef.value


def test_host_huffman_tree_and_codes_match_reference():
    """The product builds the Huffman tree on the host (huffman_host.cpp); histogram and bit packing are CUDA kernels.
    With sequential stand-ins for those two, tree blob, code table and bit stream must be the reference's own
    (HuffmanEncoder::save + ::encode) -- ties in the heap included."""
    from common import huffman_host_emul_lib, port_lib
    E = huffman_host_emul_lib()
    checker, pre = (ref_lib(), "ref") if ref_lib() is not None else (port_lib(), "orc")
    rng = np.random.default_rng(11)
    cases = [
        rng.integers(32700, 32830, 20000).astype(np.int32),
        np.full(1000, 7, np.int32),                                              # one symbol: zero-length code
        np.array([5, 9] * 300, np.int32),                                        # two symbols, equal counts
        (32768 + np.round(rng.standard_normal(50000) * 3)).astype(np.int32),     # the usual shape of an index stream
        np.concatenate([np.zeros(40, np.int32), (32768 + np.round(rng.standard_normal(30000) * 40)).astype(np.int32)]),
        rng.integers(0, 65536, 30000).astype(np.int32),                          # almost every symbol once: many ties
        np.repeat(np.arange(100, 164, dtype=np.int32), 2 ** np.arange(64) % 7 + 1),
    ]
    for q in cases:
        q = np.ascontiguousarray(q)
        outs = []
        for fn in (lambda b, tl: E.emul_huffman_encode(q.ctypes.data, q.size, 65536, b.ctypes.data, C.addressof(tl)),
                   lambda b, tl: getattr(checker, pre + "_huffman_encode")(q.ctypes.data_as(C.c_void_p), C.c_size_t(q.size),
                                                                           b.ctypes.data_as(C.c_void_p), C.byref(tl))):
            buf = np.zeros(q.size * 8 + (1 << 20), np.uint8)
            tl = C.c_size_t(0)
            n = fn(buf, tl)
            assert n > 0, n
            outs.append((tl.value, bytes(buf[:n])))
        assert outs[0][0] == outs[1][0], "tree blob length"
        assert outs[0][1][:outs[0][0]] == outs[1][1][:outs[1][0]], "tree blob"
        assert outs[0][1] == outs[1][1], "bit stream"


@pytest.mark.parametrize("shape,kw", [
    ((64, 64, 64), dict(interpAlgo=1)),
    ((64, 96, 128), dict(interpAlgo=1)),
    ((64, 64, 64), dict(interpAlgo=0)),
    ((96, 64, 128), dict(interpAlgo=0)),
    ((32, 64, 96), dict(interpAlgo=1)),
    ((128, 128, 128), dict(interpAlgo=1)),                                    # stride 4 from a compact lattice (TMA path)
    ((128, 128, 128), dict(interpAlgo=0, interpAlpha=1.5, interpBeta=3.0)),
    ((96, 128, 160), dict(interpAlgo=1)),                                     # coarse levels mixed: box / line walker
])
def test_emul_box_schedule(shape, kw):
    """Box schedule (interp_box.cuh): per-lane phase functions run lane by lane with a host copy standing in for the
    TMA box; tiles of 32 and 33 points, owned and foreign low faces, both interpolators."""
    data = field_nd(shape, np.float32)
    data[5, 6, 7] = np.nan
    data[40 % shape[0], 33, 32] = np.inf
    conf = make_config(shape, cmprAlgo=ALGO_INTERP, interpAnchorStride=32, interpDirection=0, **kw)
    for eb in (1e-2, 1e-5) if data.size < 1000000 else (1e-3,):
        q_ref, blob_ref, _ = ref_interp(ref_lib(), data, conf, eb)
        hist = np.zeros(conf.quantbinCnt, np.uint64)
        q, un = emul(data, conf, eb, 6, nthreads=32, hist=hist)
        assert np.array_equal(q, q_ref)
        _, un_ref = interp_blob_unpred(blob_ref, conf.N, np.float32)
        assert np.array_equal(un, un_ref, equal_nan=True)
        assert np.array_equal(hist, np.bincount(q_ref, minlength=conf.quantbinCnt).astype(np.uint64))
