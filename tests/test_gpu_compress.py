"""GPU end-to-end parity through sz3b_compress (the SZ_compress boundary) against the reference.

* every stream produced on the GPU is decompressed by the UNMODIFIED reference decoder and must honour the bound;
* for inputs whose pre-zstd stream is below the multi-frame threshold the whole compressed file is byte-identical;
* compression ratio within 1 % of the reference at the same bound (north_star).
"""
import ctypes as C

import numpy as np
import pytest

from common import (field_g1, ALGO_INTERP, ALGO_INTERP_LORENZO, ALGO_LORENZO_REG, ALGO_LOSSLESS, EB_ABS, EB_PSNR, EB_REL, Config, dtype_code, field_g3,
                    field_g4, field_nd, make_config, product_lib, ref_lib)

pytestmark = pytest.mark.gpu


def gpu_compress(data, conf):
    L = product_lib()
    cap = L.sz3b_compress_bound(dtype_code(data), C.byref(conf))
    out = np.empty(cap, dtype=np.uint8)
    size = C.c_size_t(0)
    used = Config()
    rc = L.sz3b_compress(dtype_code(data), C.byref(conf), data.ctypes.data_as(C.c_void_p), 0, out.ctypes.data_as(C.c_char_p),
                         C.c_size_t(cap), C.byref(size), C.byref(used))
    assert rc == 0, L.sz3b_last_error()
    return out[:size.value].copy(), used


def ref_compress(data, conf):
    R = ref_lib()
    cap = R.ref_size_bound(dtype_code(data), C.byref(conf))
    out = np.empty(cap, dtype=np.uint8)
    n = R.ref_compress(dtype_code(data), C.byref(conf), data.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_char_p), C.c_size_t(cap))
    assert n > 0
    return out[:n].copy()


def ref_decompress(cmp, like):
    R = ref_lib()
    out = np.empty_like(like)
    conf = Config()
    rc = R.ref_decompress(dtype_code(like), cmp.ctypes.data_as(C.c_char_p), C.c_size_t(cmp.size), out.ctypes.data_as(C.c_void_p), C.byref(conf))
    assert rc == 0
    return out, conf


needs_ref = pytest.mark.skipif(ref_lib() is None, reason="oracle/_ref/libsz3ref.so not built")


@needs_ref
@pytest.mark.parametrize("algo", [ALGO_INTERP, ALGO_INTERP_LORENZO])
@pytest.mark.parametrize("shape,dtype,eb", [((100, 70, 130), np.float32, 1e-3), ((64, 80, 96), np.float64, 1e-5),
                                            ((128, 128, 128), np.float32, 1e-2), ((8, 8, 128), np.float32, 1.0),
                                            ((12, 40, 40, 40), np.float32, 1e-2), ((300, 500), np.float32, 1e-3)])
def test_stream_identical_small(shape, dtype, eb, algo):
    data = field_nd(shape, dtype)
    conf = make_config(shape, cmprAlgo=algo, absErrorBound=eb)
    ours, used = gpu_compress(data, conf)
    theirs = ref_compress(data, conf)
    assert ours.size == theirs.size and np.array_equal(ours, theirs), (ours.size, theirs.size)
    dec, _ = ref_decompress(ours, data)
    assert np.max(np.abs(dec.astype(np.float64) - data.astype(np.float64))) <= eb


@needs_ref
@pytest.mark.parametrize("mode,val", [(EB_REL, 1e-4), (EB_PSNR, 80.0)])
def test_error_bound_modes(mode, val):
    data = field_g3((96, 96, 96))
    conf = make_config(data.shape, cmprAlgo=ALGO_INTERP_LORENZO, errorBoundMode=mode, relErrorBound=val, psnrErrorBound=val)
    ours, used = gpu_compress(data, conf)
    theirs = ref_compress(data, conf)
    assert np.array_equal(ours, theirs)
    dec, dconf = ref_decompress(ours, data)
    assert np.max(np.abs(dec - data)) <= dconf.absErrorBound


@needs_ref
def test_noise_falls_back_to_lossless():
    rng = np.random.default_rng(3)
    data = rng.standard_normal((64, 64, 64)).astype(np.float32)
    conf = make_config(data.shape, absErrorBound=1e-7)
    ours, used = gpu_compress(data, conf)
    theirs = ref_compress(data, conf)
    assert used.cmprAlgo == ALGO_LOSSLESS
    assert np.array_equal(ours, theirs)
    dec, _ = ref_decompress(ours, data)
    assert np.array_equal(dec, data)


@needs_ref
def test_zero_bound_is_lossless():
    data = field_nd((40, 40, 40), np.float32)
    conf = make_config(data.shape, absErrorBound=0.0)
    ours, used = gpu_compress(data, conf)
    assert used.cmprAlgo == ALGO_LOSSLESS
    dec, _ = ref_decompress(ours, data)
    assert np.array_equal(dec, data)


@needs_ref
def test_g3_256_ratio_and_bound():
    data = field_g3((256, 256, 256))
    conf = make_config(data.shape, absErrorBound=1e-3)
    ours, used = gpu_compress(data, conf)
    theirs = ref_compress(data, conf)
    dec, dconf = ref_decompress(ours, data)
    assert np.max(np.abs(dec - data)) <= 1e-3
    r_ours, r_ref = data.nbytes / ours.size, data.nbytes / theirs.size
    assert abs(r_ours - r_ref) / r_ref < 0.01, (r_ours, r_ref)
    assert dconf.cmprAlgo == ALGO_INTERP


@needs_ref
def test_omp_container_decodes_with_reference():
    data = field_g3((100, 64, 64))
    conf = make_config(data.shape, absErrorBound=1e-3, openmp=4)
    ours, used = gpu_compress(data, conf)
    dec, dconf = ref_decompress(ours, data)
    assert np.max(np.abs(dec - data)) <= 1e-3


def test_capacity_check():
    L = product_lib()
    data = field_nd((20, 20, 20), np.float32)
    conf = make_config(data.shape)
    out = np.empty(64, dtype=np.uint8)
    size = C.c_size_t(0)
    rc = L.sz3b_compress(0, C.byref(conf), data.ctypes.data_as(C.c_void_p), 0, out.ctypes.data_as(C.c_char_p), C.c_size_t(64), C.byref(size), None)
    assert rc == -1  # std::invalid_argument in the reference (sz.hpp:47-49)


@needs_ref
def test_lossless_policy_adaptive_vs_full():
    """Lossless policies (include/sz3b.h: sz3b_set_lossless_policy; 2 = GPU stage, the default): all give streams
    the unmodified reference decoder reads, identical reconstructions, and ratios within 1 % of each other and of
    the reference."""
    L = product_lib()
    data = field_g3((384, 384, 384))
    conf = make_config(data.shape, absErrorBound=1e-3)
    theirs = ref_compress(data, conf)
    out = {}
    try:
        for pol in (0, 1, 2):
            L.sz3b_set_lossless_policy(pol)
            out[pol], _ = gpu_compress(data, conf)
    finally:
        L.sz3b_set_lossless_policy(2)
    dec0, _ = ref_decompress(out[0], data)
    dec1, _ = ref_decompress(out[1], data)
    assert np.array_equal(dec0, dec1)
    assert np.max(np.abs(dec1 - data)) <= 1e-3
    dec2, _ = ref_decompress(out[2], data)
    assert np.array_equal(dec0, dec2)
    r_ref, r0, r1, r2 = (data.nbytes / x.size for x in (theirs, out[0], out[1], out[2]))
    assert abs(r0 - r_ref) / r_ref < 0.01 and abs(r1 - r_ref) / r_ref < 0.01 and abs(r2 - r_ref) / r_ref < 0.01, (r_ref, r0, r1, r2)
    assert out[1].size >= out[0].size


@needs_ref
@pytest.mark.parametrize("kind,n,eb", [("g1", 1 << 20, 1e-4), ("g1", 300000, 1e-3), ("noisy", 1 << 18, 1e-3), ("walk", 200000, 1e-2),
                                       ("g1", 20000, 1e-4)])
def test_one_dimensional_default_algorithm(kind, n, eb):
    """1-D ALGO_INTERP_LORENZO (BASELINE.json config #1 is the 2^20 case): the tuner also tries the composed
    Lorenzo(1st + 2nd order) stack on the sampled blocks (SZAlgoInterp.hpp:226-282) and may hand the whole array to
    SZ_compress_LorenzoReg; either way the stream must be the reference's, byte for byte."""
    rng = np.random.default_rng(5)
    if kind == "g1":
        data = field_g1(n)
    elif kind == "noisy":
        data = (np.sin(np.arange(n) / 50.0) + 0.05 * rng.standard_normal(n)).astype(np.float32)
    else:
        data = np.cumsum(rng.standard_normal(n)).astype(np.float32)
    conf = make_config(data.shape, cmprAlgo=ALGO_INTERP_LORENZO, absErrorBound=eb)
    ours, used = gpu_compress(data, conf)
    theirs = ref_compress(data, conf)
    dec, dconf = ref_decompress(ours, data)
    assert used.cmprAlgo == dconf.cmprAlgo
    assert ours.size == theirs.size and np.array_equal(ours, theirs), (ours.size, theirs.size, used.cmprAlgo)
    assert np.max(np.abs(dec.astype(np.float64) - data.astype(np.float64))) <= eb


@needs_ref
@pytest.mark.parametrize("shape,dtype,algo", [
    ((96, 160, 128), np.float32, ALGO_INTERP_LORENZO),     # plane-ordered copy: 3 block-rows, even nz
    ((97, 160, 128), np.float32, ALGO_INTERP_LORENZO),     # odd nz: the last block-row is closed by an even plane
    ((65, 129, 130), np.float64, ALGO_INTERP_LORENZO),     # two block-rows, the second a single plane pair
    ((64, 256, 96), np.float32, ALGO_INTERP_LORENZO),
    ((130, 100, 70), np.float32, ALGO_INTERP_LORENZO),     # planes below 64 KiB: plain background copy
    ((40, 40, 40, 24), np.float32, ALGO_INTERP_LORENZO),   # 4-D: plain background copy
])
def test_pinned_host_input_stream_identical(shape, dtype, algo):
    """Pinned host input takes the overlapped path: the tuner samples host memory while the array goes up, and for 3-D
    arrays the copy runs in plane order with predict+quantize following it (level 1 block-row by block-row).  The
    stream must not depend on how the input arrived."""
    import torch
    data = field_nd(shape, dtype)
    conf = make_config(shape, cmprAlgo=algo, absErrorBound=1e-3)
    pinned = torch.from_numpy(data).pin_memory()
    L = product_lib()
    cap = L.sz3b_compress_bound(dtype_code(data), C.byref(conf))
    out = np.empty(cap, dtype=np.uint8)
    size = C.c_size_t(0)
    theirs = ref_compress(data, conf)
    for _ in range(2):   # the second call reuses warm buffers and events
        rc = L.sz3b_compress(dtype_code(data), C.byref(conf), C.c_void_p(pinned.data_ptr()), 0, out.ctypes.data_as(C.c_char_p),
                             C.c_size_t(cap), C.byref(size), None)
        assert rc == 0, L.sz3b_last_error()
        assert size.value == theirs.size and np.array_equal(out[:size.value], theirs)


def test_concurrent_callers_streams_identical():
    """The library is reentrant (every call borrows its own workspace and streams): three host threads compressing
    different arrays at the same time -- pinned host input on the plane-ordered path, pageable host input, a blockwise
    stack -- must each produce the stream a lone caller gets, call after call."""
    import threading

    import torch
    L = product_lib()
    L.sz3b_last_error.restype = C.c_char_p
    a = field_nd((96, 160, 128), np.float32)
    b = field_g3((80, 90, 100), np.float32)
    c = field_nd((60, 66, 72), np.float64)
    pinned = torch.from_numpy(a).pin_memory()
    jobs = [
        (a, make_config(a.shape, cmprAlgo=ALGO_INTERP_LORENZO, absErrorBound=1e-3), pinned.data_ptr()),
        (b, make_config(b.shape, cmprAlgo=ALGO_INTERP_LORENZO, absErrorBound=1e-3), b.ctypes.data),
        (c, make_config(c.shape, cmprAlgo=ALGO_LORENZO_REG, absErrorBound=1e-4), c.ctypes.data),
    ]
    alone = [gpu_compress(d, conf)[0] for d, conf, _ in jobs]
    failures = []

    def caller(k):
        d, conf, ptr = jobs[k]
        cap = L.sz3b_compress_bound(dtype_code(d), C.byref(conf))
        out = np.empty(cap, dtype=np.uint8)
        size = C.c_size_t(0)
        for it in range(4):
            rc = L.sz3b_compress(dtype_code(d), C.byref(conf), C.c_void_p(ptr), 0, out.ctypes.data_as(C.c_char_p), C.c_size_t(cap),
                                 C.byref(size), None)
            if rc != 0:
                failures.append((k, it, L.sz3b_last_error()))
            elif size.value != alone[k].size or not np.array_equal(out[:size.value], alone[k]):
                failures.append((k, it, "stream differs from the lone caller's", size.value, alone[k].size))

    threads = [threading.Thread(target=caller, args=(k,)) for k in range(len(jobs))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not failures, failures
