"""Integer element types (int32 / int64, what tools/sz3/sz3.cpp:458-461 instantiates) through sz3b_compress /
sz3b_decompress against the UNMODIFIED reference: whole compressed files byte-identical (the streams here stay below
the multi-frame threshold, so the lossless stage is the reference's own zstd call), both decoders agree bit for bit,
the bound holds.  Covered: ALGO_INTERP and ALGO_INTERP_LORENZO (tuner, 1-D Lorenzo hand-over included) for N = 1..4,
both interpolators, every predictor stack of ALGO_LORENZO_REG (Lorenzo, regression, composed), REL bounds, the lossless
fallbacks."""
import ctypes as C

import numpy as np
import pytest

from common import (ALGO_INTERP, ALGO_INTERP_LORENZO, ALGO_LORENZO_REG, EB_ABS, EB_REL, Config, dtype_code, field_g1, field_nd,
                    make_config, product_lib, ref_lib)
from test_gpu_compress import gpu_compress, ref_compress, ref_decompress

pytestmark = pytest.mark.gpu
needs_ref = pytest.mark.skipif(ref_lib() is None, reason="oracle/_ref/libsz3ref.so not built")


def int_field(shape, dtype, scale, seed=7):
    g = field_nd(shape, np.float64, seed) if len(shape) > 1 else field_g1(shape[0], seed).astype(np.float64)
    return np.ascontiguousarray(np.rint(g * scale).astype(dtype))


def gpu_decompress(cmp, like):
    L = product_lib()
    out = np.empty_like(like)
    conf = Config()
    rc = L.sz3b_decompress(dtype_code(like), cmp.ctypes.data_as(C.c_char_p), C.c_size_t(cmp.size), out.ctypes.data_as(C.c_void_p), 0,
                           C.byref(conf))
    assert rc == 0, L.sz3b_last_error()
    return out


def check(data, conf, eb_abs=None):
    ours, used = gpu_compress(data, conf)
    theirs = ref_compress(data, conf)
    assert ours.size == theirs.size and np.array_equal(ours, theirs), (ours.size, theirs.size)
    dec_ref, _ = ref_decompress(ours, data)
    dec_gpu = gpu_decompress(theirs, data)
    assert np.array_equal(dec_ref, dec_gpu)
    if eb_abs is not None:
        err = np.abs(dec_gpu.astype(np.float64) - data.astype(np.float64)).max()
        assert err <= eb_abs


@needs_ref
@pytest.mark.parametrize("dtype", [np.int32, np.int64])
@pytest.mark.parametrize("algo", [ALGO_INTERP, ALGO_INTERP_LORENZO])
@pytest.mark.parametrize("shape,scale,eb", [((70, 50, 90), 1000.0, 2.0), ((64, 64, 64), 30.0, 1.0), ((200, 300), 5000.0, 10.0),
                                            ((20000,), 1.0e5, 3.0), ((10, 24, 24, 24), 800.0, 4.5)])
def test_int_interp_stream_identical(shape, scale, eb, algo, dtype):
    data = int_field(shape, dtype, scale)
    check(data, make_config(shape, cmprAlgo=algo, absErrorBound=eb), eb)


@needs_ref
@pytest.mark.parametrize("dtype", [np.int32, np.int64])
@pytest.mark.parametrize("interp", [0, 1])
@pytest.mark.parametrize("direction", [0, 5])
def test_int_interp_variants(dtype, interp, direction):
    data = int_field((48, 40, 56), dtype, 2000.0)
    check(data, make_config(data.shape, cmprAlgo=ALGO_INTERP, absErrorBound=3.0, interpAlgo=interp, interpDirection=direction), 3.0)


@needs_ref
@pytest.mark.parametrize("dtype", [np.int32, np.int64])
@pytest.mark.parametrize("shape", [(40, 36, 50), (150, 130), (5000,)])
@pytest.mark.parametrize("l1,l2", [(1, 0), (1, 1), (0, 1)])
def test_int_lorenzo_stacks(dtype, shape, l1, l2):
    data = int_field(shape, dtype, 3000.0)
    check(data, make_config(shape, cmprAlgo=ALGO_LORENZO_REG, absErrorBound=2.0, lorenzo=l1, lorenzo2=l2, regression=0, regression2=0), 2.0)


@needs_ref
@pytest.mark.parametrize("dtype", [np.int32, np.int64])
def test_int_rel_bound_and_large_values(dtype):
    data = int_field((60, 50, 40), dtype, 1.0e6 if dtype == np.int32 else 1.0e12)
    check(data, make_config(data.shape, cmprAlgo=ALGO_INTERP_LORENZO, errorBoundMode=EB_REL, relErrorBound=1e-4))


@needs_ref
@pytest.mark.parametrize("dtype", [np.int32, np.int64])
def test_int_lossless_fallbacks(dtype):
    rng = np.random.default_rng(5)
    noise = rng.integers(-2**30, 2**30, size=(40, 40, 40)).astype(dtype)
    check(noise, make_config(noise.shape, cmprAlgo=ALGO_INTERP_LORENZO, absErrorBound=1.0))     # ratio < 3 -> lossless
    const = np.full((33, 20, 17), 12345, dtype=dtype)
    check(const, make_config(const.shape, cmprAlgo=ALGO_INTERP_LORENZO, errorBoundMode=EB_REL, relErrorBound=1e-3))   # range 0


@needs_ref
@pytest.mark.parametrize("dtype", [np.int32, np.int64])
@pytest.mark.parametrize("shape,offset", [((40, 36, 50), 40000.0), ((40, 36, 50), 0.0), ((150, 130), 9000.0), ((6, 20, 22, 24), 50000.0)])
@pytest.mark.parametrize("l1,l2,reg", [(0, 0, 1), (1, 0, 1), (1, 1, 1)])
def test_int_regression_stacks(dtype, shape, offset, l1, l2, reg):
    # The reference's fit multiplies a size_t index by the value (RegressionPredictor.hpp:44): negative integers wrap
    # to ~2^64, the coefficients become x86's indefinite conversion results, and the quantizer's int64 cast overflows
    # in turn (LinearQuantizer.hpp:45) -- the reference's own round trip then breaks the bound on a few points.  The
    # GPU path reproduces all of it bit for bit (offset 0 = data with negative values); the bound is only asserted
    # where the reference itself keeps it.
    g = field_nd(shape, np.float64)
    data = np.ascontiguousarray(np.rint(g * 3000.0 + offset).astype(dtype))
    check(data, make_config(shape, cmprAlgo=ALGO_LORENZO_REG, absErrorBound=2.0, lorenzo=l1, lorenzo2=l2, regression=reg, regression2=0),
          2.0 if offset > 0 else None)
