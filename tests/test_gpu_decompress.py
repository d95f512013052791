"""GPU parity of the SZ_decompress half (sz3b_decompress) against the reference decoder.

Bar: the reconstructed array is BIT-IDENTICAL to what the unmodified reference decoder produces from the same stream
(recover() is deterministic arithmetic), for streams written by either implementation, and honours the bound."""
import ctypes as C

import numpy as np
import pytest

from common import (ALGO_INTERP, ALGO_INTERP_LORENZO, ALGO_LORENZO_REG, EB_REL, Config, dtype_code, field_g3, field_nd,
                    make_config, product_lib, ref_lib)
from test_gpu_compress import gpu_compress, ref_compress, ref_decompress

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(ref_lib() is None, reason="oracle/_ref/libsz3ref.so not built")]


def gpu_decompress(cmp, like, device=False):
    L = product_lib()
    conf = Config()
    cmp = np.ascontiguousarray(cmp)
    if device:
        import torch
        out_t = torch.empty(like.shape, dtype=torch.float32 if like.dtype == np.float32 else torch.float64, device="cuda")
        rc = L.sz3b_decompress(dtype_code(like), cmp.ctypes.data_as(C.c_char_p), C.c_size_t(cmp.size), C.c_void_p(out_t.data_ptr()), 1, C.byref(conf))
        assert rc == 0, L.sz3b_last_error()
        return out_t.cpu().numpy(), conf
    out = np.empty_like(like)
    rc = L.sz3b_decompress(dtype_code(like), cmp.ctypes.data_as(C.c_char_p), C.c_size_t(cmp.size), out.ctypes.data_as(C.c_void_p), 0, C.byref(conf))
    assert rc == 0, L.sz3b_last_error()
    return out, conf


def same_bits(a, b):
    return np.array_equal(a.view(np.uint8), b.view(np.uint8))


@pytest.mark.parametrize("writer", ["gpu", "ref"])
@pytest.mark.parametrize("shape,dtype,kw", [
    ((100, 70, 130), np.float32, dict(cmprAlgo=ALGO_INTERP_LORENZO, absErrorBound=1e-3)),
    ((64, 80, 96), np.float64, dict(cmprAlgo=ALGO_INTERP, absErrorBound=1e-5, interpAlgo=0, interpDirection=5)),
    ((33, 65, 97), np.float32, dict(cmprAlgo=ALGO_INTERP, absErrorBound=1e-2, interpAlgo=0, interpDirection=2)),
    ((64, 64, 64), np.float32, dict(cmprAlgo=ALGO_INTERP, absErrorBound=1e-2, interpAlgo=1, interpDirection=3)),
    ((12, 40, 40, 40), np.float32, dict(cmprAlgo=ALGO_INTERP_LORENZO, absErrorBound=1e-2)),
    ((300, 500), np.float32, dict(cmprAlgo=ALGO_INTERP_LORENZO, absErrorBound=1e-3)),
    ((5000,), np.float32, dict(cmprAlgo=ALGO_INTERP, absErrorBound=1e-3)),
    ((8, 8, 128), np.float32, dict(cmprAlgo=ALGO_INTERP, absErrorBound=1e-3, interpAnchorStride=0)),
    ((37, 41, 130), np.float32, dict(cmprAlgo=ALGO_INTERP, absErrorBound=1e-6, quantbinCnt=16)),
    ((60, 66, 72), np.float64, dict(cmprAlgo=ALGO_LORENZO_REG, lorenzo=0, regression=1, errorBoundMode=EB_REL, relErrorBound=1e-4)),
    ((40, 45), np.float32, dict(cmprAlgo=ALGO_LORENZO_REG, lorenzo=0, regression=1, absErrorBound=1e-3)),
    ((48, 48, 48), np.float32, dict(cmprAlgo=ALGO_LORENZO_REG, lorenzo=0, regression=1, absErrorBound=1e-7)),
    ((50, 43, 61), np.float32, dict(cmprAlgo=ALGO_LORENZO_REG, absErrorBound=1e-3)),                     # Lorenzo + regression
    ((30, 31, 32), np.float64, dict(cmprAlgo=ALGO_LORENZO_REG, lorenzo2=1, absErrorBound=1e-4)),         # all three
    ((24, 30, 36), np.float32, dict(cmprAlgo=ALGO_LORENZO_REG, regression=0, absErrorBound=1e-3)),       # Lorenzo alone
    ((31, 37, 25), np.float32, dict(cmprAlgo=ALGO_LORENZO_REG, lorenzo2=1, regression=0, absErrorBound=1e-3)),
    ((130, 77), np.float64, dict(cmprAlgo=ALGO_LORENZO_REG, lorenzo2=1, blockSize=16, absErrorBound=1e-3)),
    ((3000,), np.float32, dict(cmprAlgo=ALGO_LORENZO_REG, blockSize=128, absErrorBound=1e-3)),
    ((9, 12, 20, 18), np.float32, dict(cmprAlgo=ALGO_LORENZO_REG, absErrorBound=1e-3)),
    ((25, 31, 37), np.float32, dict(cmprAlgo=ALGO_LORENZO_REG, absErrorBound=1e-3, quantbinCnt=16)),
    # shapes of 32-multiples (+1): the last pass of the finest level through k_box_recover_x (TMA planes, lane = row)
    ((64, 96, 128), np.float32, dict(cmprAlgo=ALGO_INTERP, absErrorBound=1e-3, interpAlgo=1, interpDirection=0)),
    ((64, 96, 128), np.float32, dict(cmprAlgo=ALGO_INTERP, absErrorBound=1e-3, interpAlgo=0, interpDirection=0)),
    ((97, 65, 129), np.float32, dict(cmprAlgo=ALGO_INTERP, absErrorBound=1e-2, interpAlgo=1, interpDirection=0)),
    ((96, 64, 64), np.float32, dict(cmprAlgo=ALGO_INTERP, absErrorBound=1e-7, interpAlgo=1, interpDirection=0, quantbinCnt=64)),   # many stored values
    ((128, 128, 128), np.float32, dict(cmprAlgo=ALGO_INTERP_LORENZO, absErrorBound=1e-3)),
])
def test_decompress_bit_identical(shape, dtype, kw, writer):
    data = field_nd(shape, dtype)
    conf = make_config(shape, **kw)
    cmp = gpu_compress(data, conf)[0] if writer == "gpu" else ref_compress(data, conf)
    want, wconf = ref_decompress(cmp, data)
    got, gconf = gpu_decompress(cmp, data)
    assert same_bits(got, want), f"{int((got != want).sum())} elements differ"
    assert np.max(np.abs(got.astype(np.float64) - data.astype(np.float64))) <= wconf.absErrorBound
    assert gconf.cmprAlgo == wconf.cmprAlgo and gconf.absErrorBound == wconf.absErrorBound


def test_decompress_to_device_pointer():
    data = field_g3((96, 96, 96))
    conf = make_config(data.shape, absErrorBound=1e-3)
    cmp, _ = gpu_compress(data, conf)
    want, _ = ref_decompress(cmp, data)
    got, _ = gpu_decompress(cmp, data, device=True)
    assert same_bits(got, want)


def test_decompress_lossless_and_noise():
    rng = np.random.default_rng(3)
    data = rng.standard_normal((64, 64, 64)).astype(np.float32)
    conf = make_config(data.shape, absErrorBound=1e-7)     # falls back to ALGO_LOSSLESS
    for cmp in (gpu_compress(data, conf)[0], ref_compress(data, conf)):
        got, gconf = gpu_decompress(cmp, data)
        assert np.array_equal(got, data)


@pytest.mark.parametrize("writer", ["gpu", "ref"])
def test_decompress_omp_container(writer):
    data = field_g3((100, 64, 64))
    conf = make_config(data.shape, absErrorBound=1e-3, openmp=4 if writer == "gpu" else 1)
    cmp = gpu_compress(data, conf)[0] if writer == "gpu" else ref_compress(data, conf)
    want, _ = ref_decompress(cmp, data)
    got, gconf = gpu_decompress(cmp, data)
    assert same_bits(got, want)
    assert np.max(np.abs(got - data)) <= 1e-3


def test_decompress_special_values():
    data = field_nd((40, 50, 70), np.float32)
    data[3, 4, 5] = np.nan
    data[10, 11, 12] = np.inf
    data[20, 2, 7] = -np.inf
    conf = make_config(data.shape, cmprAlgo=ALGO_INTERP, absErrorBound=1e-3)
    cmp, _ = gpu_compress(data, conf)
    want, _ = ref_decompress(cmp, data)
    got, _ = gpu_decompress(cmp, data)
    assert same_bits(got, want)


def test_decompress_rejects_garbage():
    L = product_lib()
    junk = np.zeros(64, dtype=np.uint8)
    out = np.empty(8, dtype=np.float32)
    conf = Config()
    rc = L.sz3b_decompress(0, junk.ctypes.data_as(C.c_char_p), C.c_size_t(junk.size), out.ctypes.data_as(C.c_void_p), 0, C.byref(conf))
    assert rc == -1     # std::invalid_argument: magic number mismatch (sz.hpp:123-125)


def test_g3_256_roundtrip_through_python_binding():
    import sz3_b200
    from sz3_b200 import sz, szConfig
    data = field_g3((256, 256, 256))
    conf = szConfig(*data.shape)
    conf.absErrorBound = 1e-3
    cmp, ratio = sz.compress(data, conf)
    dec, dconf = sz.decompress(cmp, np.float32, data.shape)
    assert np.max(np.abs(dec - data)) <= 1e-3
    want, _ = ref_decompress(np.ascontiguousarray(cmp), data)
    assert same_bits(dec, want)


def _last_stages(L):
    names, ms, ln = (C.c_char_p * 64)(), (C.c_double * 64)(), (C.c_int * 64)()
    k = L.sz3b_last_profile(names, ms, ln, 64)
    return [names[i].decode() for i in range(k)]


@pytest.mark.parametrize("shape,dtype,eb,writer", [
    ((256, 256, 256), np.float32, 1e-3, "gpu"),       # 5.6 MB of frames from the GPU lossless stage
    ((256, 256, 256), np.float64, 1e-4, "gpu"),
    ((200, 300, 260), np.float32, 1e-4, "gpu"),       # ragged last block / frame
    ((256, 256, 256), np.float32, 1e-3, "ref"),       # libzstd's own frame: declined or decoded, same bits either way
    ((256, 256, 256), np.float32, 1e-3, "host"),      # host zstd policy: multi-frame libzstd output
])
def test_frames_decoded_on_gpu(shape, dtype, eb, writer):
    """Frame decoder 1 (sz3b_set_frame_decoder; sz3_b200/csrc/zhuf_dec.cuh): the frames of the GPU lossless stage are
    decoded on the GPU -- the stage list of the call shows it -- and the array is bit-identical to the reference
    decoder's; payloads of any other shape go through libzstd as before."""
    L = product_lib()
    data = field_g3(shape, dtype)
    conf = make_config(shape, cmprAlgo=ALGO_INTERP, absErrorBound=eb)
    if writer == "gpu":
        cmp = gpu_compress(data, conf)[0]
    elif writer == "ref":
        cmp = ref_compress(data, conf)
    else:
        L.sz3b_set_lossless_policy(0)
        try:
            cmp = gpu_compress(data, conf)[0]
        finally:
            L.sz3b_set_lossless_policy(2)
    want, _ = ref_decompress(cmp, data)
    before = L.sz3b_get_frame_decoder()
    try:
        L.sz3b_set_frame_decoder(1)
        got, _ = gpu_decompress(cmp, data)
        stages = _last_stages(L)
        got_dev, _ = gpu_decompress(cmp, data, device=True)
        L.sz3b_set_frame_decoder(0)
        plain, _ = gpu_decompress(cmp, data)
        assert "frames_gpu" not in _last_stages(L)
    finally:
        L.sz3b_set_frame_decoder(before)
    assert same_bits(got, want) and same_bits(got_dev, want) and same_bits(plain, want)
    if writer == "gpu":
        assert "frames_gpu" in stages and "zstd_host" not in stages, stages


def test_frames_on_gpu_damaged_payload():
    """Damage inside the frames: the call must end the way the host path ends (an error, or an array -- the frames carry
    no checksum), never crash."""
    L = product_lib()
    data = field_g3((256, 256, 256))
    conf = make_config(data.shape, cmprAlgo=ALGO_INTERP, absErrorBound=1e-3)
    cmp = gpu_compress(data, conf)[0]
    rng = np.random.default_rng(9)
    before = L.sz3b_get_frame_decoder()
    try:
        L.sz3b_set_frame_decoder(1)
        for _ in range(6):
            bad = cmp.copy()
            lo = 64
            i = int(rng.integers(lo, bad.size - 512))
            bad[i] ^= np.uint8(0x10)
            out = np.empty_like(data)
            c2 = Config()
            rc = L.sz3b_decompress(0, bad.ctypes.data_as(C.c_char_p), C.c_size_t(bad.size), out.ctypes.data_as(C.c_void_p), 0, C.byref(c2))
            assert rc in (0, -1, -2, -3, -4, -5), rc
    finally:
        L.sz3b_set_frame_decoder(before)


def test_container_slabs_frames_on_gpu():
    """An OpenMP container whose slabs are large enough for the GPU lossless stage: every slab's frames go through the
    GPU frame decoder, on the container's worker threads, and the array equals the reference decoder's."""
    L = product_lib()
    data = field_g3((512, 256, 256))
    conf = make_config(data.shape, cmprAlgo=ALGO_INTERP, absErrorBound=1e-3, openmp=2)
    cmp = gpu_compress(data, conf)[0]
    want, _ = ref_decompress(cmp, data)
    assert L.sz3b_get_frame_decoder() == 1
    for _ in range(2):
        got, _ = gpu_decompress(cmp, data)
        assert same_bits(got, want)
