"""GPU end-to-end parity through sz3b_compress (the SZ_compress boundary) against the reference.

* every stream produced on the GPU is decompressed by the UNMODIFIED reference decoder and must honour the bound;
* for inputs whose pre-zstd stream is below the multi-frame threshold the whole compressed file is byte-identical;
* compression ratio within 1 % of the reference at the same bound (north_star).
"""
import ctypes as C

import numpy as np
import pytest

from common import (EB_ABS_AND_REL, EB_ABS_OR_REL, EB_L2NORM, field_g1, ALGO_INTERP, ALGO_INTERP_LORENZO, ALGO_LORENZO_REG, ALGO_LOSSLESS, EB_ABS, EB_PSNR, EB_REL, Config, dtype_code, field_g3,
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
@pytest.mark.parametrize("mode,kw", [
    (EB_L2NORM, dict(l2normErrorBound=2.0)),
    (EB_ABS_AND_REL, dict(absErrorBound=1e-3, relErrorBound=1e-3)),    # rel * range ~ 3.5e-3: the absolute bound wins
    (EB_ABS_AND_REL, dict(absErrorBound=1e-2, relErrorBound=1e-4)),    # the relative bound wins
    (EB_ABS_OR_REL, dict(absErrorBound=1e-3, relErrorBound=1e-3)),     # the relative bound wins
    (EB_ABS_OR_REL, dict(absErrorBound=1e-2, relErrorBound=1e-4)),     # the absolute bound wins
])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_error_bound_modes_l2_and_or(mode, kw, dtype):
    """calAbsErrorBound (Statistic.hpp:32-56) for the modes that combine or derive bounds: the stream (whose trailing
    Config carries the resolved absolute bound) must be the reference's byte for byte."""
    data = field_g3((96, 96, 96), dtype)
    conf = make_config(data.shape, cmprAlgo=ALGO_INTERP_LORENZO, errorBoundMode=mode, **kw)
    ours, used = gpu_compress(data, conf)
    theirs = ref_compress(data, conf)
    assert ours.size == theirs.size and np.array_equal(ours, theirs)
    dec, dconf = ref_decompress(ours, data)
    assert dconf.errorBoundMode == EB_ABS and dconf.absErrorBound == used.absErrorBound
    if mode == EB_L2NORM:
        err = dec.astype(np.float64) - data.astype(np.float64)
        assert np.sqrt((err ** 2).sum()) <= kw["l2normErrorBound"]
    else:
        assert np.max(np.abs(dec.astype(np.float64) - data.astype(np.float64))) <= dconf.absErrorBound


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


def _set_ref_threads(n):
    R = ref_lib()
    R.ref_set_threads(int(n))


@needs_ref
@pytest.mark.parametrize("shape,dtype,nslabs,kw", [
    ((100, 64, 64), np.float32, 4, dict(absErrorBound=1e-3)),
    ((128, 96, 96), np.float32, 2, dict(errorBoundMode=EB_REL, relErrorBound=1e-4)),
    ((64, 72, 80), np.float64, 8, dict(errorBoundMode=EB_PSNR, psnrErrorBound=80.0)),
    ((37, 50, 60), np.float32, 3, dict(absErrorBound=1e-2)),                       # uneven slabs
    ((16, 40, 40, 24), np.float32, 2, dict(errorBoundMode=EB_REL, relErrorBound=1e-3)),
])
def test_omp_container_identical_to_reference(shape, dtype, nslabs, kw):
    """conf.openmp = n on the GPU side against the reference's SZ_compress_OMP run with n OpenMP threads
    (SZImplOMP.hpp:43-108): same slab bounds, same shared bound, same per-slab Configs -- the whole container byte for
    byte, trailing outer Config included."""
    data = field_nd(shape, dtype) if len(shape) == 4 else field_g3(shape, dtype)
    conf = make_config(shape, cmprAlgo=ALGO_INTERP_LORENZO, openmp=nslabs, **kw)
    ours, used = gpu_compress(data, conf)
    _set_ref_threads(nslabs)
    try:
        rconf = make_config(shape, cmprAlgo=ALGO_INTERP_LORENZO, openmp=1, **kw)
        theirs = ref_compress(data, rconf)
    finally:
        _set_ref_threads(len(__import__("os").sched_getaffinity(0)))
    assert ours.size == theirs.size and np.array_equal(ours, theirs), (ours.size, theirs.size)
    dec, dconf = ref_decompress(ours, data)
    assert np.max(np.abs(dec.astype(np.float64) - data.astype(np.float64))) <= dconf.absErrorBound


@needs_ref
def test_omp_container_constant_field_is_lossless():
    """A constant array under a relative bound: range 0 -> absErrorBound 0 -> every slab stored losslessly, as the
    reference does (no error about a missing range)."""
    data = np.full((40, 30, 30), 2.5, dtype=np.float32)
    conf = make_config(data.shape, cmprAlgo=ALGO_INTERP_LORENZO, openmp=4, errorBoundMode=EB_REL, relErrorBound=1e-3)
    ours, used = gpu_compress(data, conf)
    dec, dconf = ref_decompress(ours, data)
    assert np.array_equal(dec, data)
    L = product_lib()
    # the slab entry point with range 0 (what sharded callers pass after their all-reduce)
    blob, blen, size = (C.c_ubyte * 256)(), C.c_size_t(0), C.c_size_t(0)
    out = np.empty(data.nbytes + 4096, np.uint8)
    rc = L.sz3b_compress_slab(0, C.byref(conf), 0, 4, data[:10].ctypes.data_as(C.c_void_p), 0, C.c_double(0.0),
                              out.ctypes.data_as(C.c_char_p), C.c_size_t(out.size), C.byref(size), blob, C.byref(blen))
    assert rc == 0, L.sz3b_last_error()
    rc = L.sz3b_compress_slab(0, C.byref(conf), 0, 4, data[:10].ctypes.data_as(C.c_void_p), 0, C.c_double(-1.0),
                              out.ctypes.data_as(C.c_char_p), C.c_size_t(out.size), C.byref(size), blob, C.byref(blen))
    assert rc == -1


@needs_ref
@pytest.mark.parametrize("policy", [0, 2])
def test_g3_512_headline_stream(policy):
    """The headline array (512^3 G3, abs 1e-3) through sz3b_compress with both lossless policies: the unmodified
    reference decodes it within the bound, the tuner picks what the reference picks, and the ratio is within 1 % of
    the reference's serial stream (equal to it for policy 0 up to the multi-frame zstd framing)."""
    L = product_lib()
    data = field_g3((512, 512, 512))
    conf = make_config(data.shape, absErrorBound=1e-3)
    theirs = ref_compress(data, conf)
    L.sz3b_set_lossless_policy(policy)
    try:
        ours, used = gpu_compress(data, conf)
    finally:
        L.sz3b_set_lossless_policy(2)
    dec, dconf = ref_decompress(ours, data)
    assert np.max(np.abs(dec - data)) <= 1e-3
    _, rconf = ref_decompress(theirs, data)
    assert (used.cmprAlgo, used.interpAlgo, used.interpDirection) == (ALGO_INTERP, 1, 0)
    assert dconf.cmprAlgo == rconf.cmprAlgo
    r_ours, r_ref = data.nbytes / ours.size, data.nbytes / theirs.size
    assert abs(r_ours - r_ref) / r_ref < 0.01, (r_ours, r_ref)


@needs_ref
def test_config3_full_size_stream_identical():
    """BASELINE.json config #3 at its full size: 384^3 float64, regression predictor only, REL 1e-4 -- the compressed
    file byte for byte (the coefficient chain, the side streams and the data indices all enter it)."""
    L = product_lib()
    data = field_g3((384, 384, 384), np.float64)
    conf = make_config(data.shape, cmprAlgo=ALGO_LORENZO_REG, errorBoundMode=EB_REL, relErrorBound=1e-4, lorenzo=0, lorenzo2=0,
                       regression=1)
    theirs = ref_compress(data, conf)
    L.sz3b_set_lossless_policy(0)
    L.sz3b_set_host_threads(16)   # (the split into zstd frames follows the pool size: fixed, so that the size check is)
    try:
        ours, used = gpu_compress(data, conf)
    finally:
        L.sz3b_set_lossless_policy(2)
        L.sz3b_set_host_threads(0)
    dec, dconf = ref_decompress(ours, data)
    assert np.max(np.abs(dec - data)) <= dconf.absErrorBound
    dec_ref, _ = ref_decompress(theirs, data)
    assert np.array_equal(dec, dec_ref)                      # same indices, same coefficients -> same reconstruction
    r_ours, r_ref = data.nbytes / ours.size, data.nbytes / theirs.size
    assert abs(r_ours - r_ref) / r_ref < 0.001, (r_ours, r_ref)
    if ours.size == theirs.size:                              # single-frame zstd: the file itself is identical
        assert np.array_equal(ours, theirs)


def test_multi_gpu_container_in_one_call():
    """conf.openmp with several visible GPUs: one sz3b_compress call spreads the slabs over the devices (one host thread
    per device inside the call) and must return the container a single device writes, byte for byte -- for host input,
    for device-resident input (peer copies), with an absolute and with a range-dependent bound."""
    import torch
    L = product_lib()
    ndev = L.sz3b_device_count()
    if ndev < 2:
        pytest.skip("needs at least two GPUs")
    data = field_g3((256, 128, 160))
    for kw in (dict(absErrorBound=1e-3), dict(errorBoundMode=EB_REL, relErrorBound=1e-4)):
        for nslabs in (ndev, 2 * ndev + 1):
            conf = make_config(data.shape, cmprAlgo=ALGO_INTERP_LORENZO, openmp=nslabs, **kw)
            L.sz3b_set_device_fanout(1)
            one, _ = gpu_compress(data, conf)
            L.sz3b_set_device_fanout(0)
            try:
                many, _ = gpu_compress(data, conf)
                dev = torch.from_numpy(data).cuda()
                cap = L.sz3b_compress_bound(0, C.byref(conf))
                out = np.empty(cap, dtype=np.uint8)
                size = C.c_size_t(0)
                rc = L.sz3b_compress(0, C.byref(conf), C.c_void_p(dev.data_ptr()), 1, out.ctypes.data_as(C.c_char_p),
                                     C.c_size_t(cap), C.byref(size), None)
                assert rc == 0, L.sz3b_last_error()
            finally:
                L.sz3b_set_device_fanout(-1)
            assert one.size == many.size and np.array_equal(one, many)
            assert size.value == one.size and np.array_equal(out[:size.value], one)
    if ref_lib() is not None:
        dec, dconf = ref_decompress(many, data)
        assert np.max(np.abs(dec - data)) <= dconf.absErrorBound


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
