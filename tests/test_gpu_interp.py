"""GPU parity: the fused interpolation predict+quantize kernels against the reference decomposition.

Bar (BASELINE.json north_star): bit-identical quantization indices and unpredictable values, same order.
Checker = oracle/_ref/libsz3ref.so (the unmodified reference; InterpolationDecomposition::compress + save) and,
where present, the C restatement oracle/libsz3oracle.so.  Every call goes through the C ABI (include/sz3b.h).
"""
import ctypes as C

import numpy as np
import pytest

from common import (ALGO_INTERP, dtype_code, field_g3, field_g4, field_nd, interp_blob_unpred, make_config, port_lib,
                    product_lib, ref_interp, ref_lib)

pytestmark = pytest.mark.gpu


def checker():
    lib = ref_lib()
    if lib is not None:
        return lib, "ref"
    lib = port_lib()
    assert lib is not None, "neither oracle/_ref/libsz3ref.so nor oracle/libsz3oracle.so is available"
    return lib, "orc"


def gpu_interp(data, conf, eb, schedule):
    L = product_lib()
    assert L is not None, "sz3_b200/lib/libsz3b200.so missing (run make)"
    n = data.size
    q = np.empty(n, dtype=np.int32)
    blob = np.empty(n * data.itemsize + 4096, dtype=np.uint8)
    blen = C.c_size_t(0)
    rc = L.sz3b_interp_decompose(dtype_code(data), C.byref(conf), C.c_double(eb), data.ctypes.data_as(C.c_void_p), 0, schedule,
                                 q.ctypes.data_as(C.c_void_p), blob.ctypes.data_as(C.c_void_p), C.c_size_t(blob.size), C.byref(blen))
    assert rc == 0, L.sz3b_last_error()
    return q, bytes(blob[:blen.value])


def check(shape, dtype, eb, schedule, data=None, **kw):
    lib, prefix = checker()
    data = field_nd(shape, dtype) if data is None else data
    conf = make_config(shape, cmprAlgo=ALGO_INTERP, **kw)
    if conf.interpAnchorStride < 0:
        conf.interpAnchorStride = [4096, 128, 32, 16][conf.N - 1]
    q_ref, blob_ref, _ = ref_interp(lib, data, conf, eb, prefix)
    q, blob = gpu_interp(data, conf, eb, schedule)
    assert np.array_equal(q, q_ref), f"{int((q != q_ref).sum())} of {q.size} indices differ"
    assert blob == blob_ref, "decomposition blob (header + unpredictable values) differs"


SHAPES3 = [(40, 50, 70), (33, 65, 97), (100, 70, 130), (64, 64, 64), (20, 20, 20), (8, 8, 128), (2, 3, 200), (33, 33, 33)]


@pytest.mark.parametrize("schedule", [1, 4, 5])
@pytest.mark.parametrize("shape", SHAPES3)
@pytest.mark.parametrize("algo", [0, 1])
@pytest.mark.parametrize("direction", [0, 5, 2])
def test_interp3d_f32(shape, algo, direction, schedule):
    check(shape, np.float32, 1e-2, schedule, interpAlgo=algo, interpDirection=direction)


@pytest.mark.parametrize("schedule", [1, 4, 5])
@pytest.mark.parametrize("kw", [
    dict(interpAlgo=1, interpDirection=0), dict(interpAlgo=0, interpDirection=3),
    dict(interpAlgo=1, interpDirection=1, interpAlpha=-1.0), dict(interpAlgo=1, interpDirection=4, interpAlpha=2.0, interpBeta=3.0),
    dict(interpAlgo=1, interpAnchorStride=8), dict(interpAlgo=0, interpAnchorStride=64), dict(interpAlgo=0, interpAnchorStride=128),
    dict(interpAlgo=1, quantbinCnt=64), dict(interpAlgo=1, quantbinCnt=262144),
])
def test_interp3d_f64_variants(kw, schedule):
    check((50, 60, 70), np.float64, 1e-4, schedule, **kw)


@pytest.mark.parametrize("schedule", [1, 4, 5])
def test_interp3d_many_unpredictable(schedule):
    # eb far below the data resolution and a tiny quantizer: most points overflow the radius
    check((37, 41, 130), np.float32, 1e-6, schedule, interpAlgo=1, quantbinCnt=16)


@pytest.mark.parametrize("schedule", [1, 4, 5])
def test_interp3d_special_values(schedule):
    data = field_nd((40, 40, 40), np.float32)
    data[3, 4, 5] = np.nan
    data[10, 11, 12] = np.inf
    data[20, 21, 22] = -np.inf
    data[30, 31, 32] = 1e30
    lib, prefix = checker()
    conf = make_config(data.shape, cmprAlgo=ALGO_INTERP, interpAnchorStride=32)
    q_ref, blob_ref, _ = ref_interp(lib, data, conf, 1e-3, prefix)
    q, blob = gpu_interp(data, conf, 1e-3, schedule)
    assert np.array_equal(q, q_ref)
    assert blob == blob_ref


@pytest.mark.parametrize("direction", [0, 23, 7])
@pytest.mark.parametrize("algo", [0, 1])
def test_interp4d(direction, algo):
    check((9, 20, 35, 40), np.float32, 1e-2, 0, interpAlgo=algo, interpDirection=direction)


def test_interp4d_f64():
    check((17, 17, 17, 17), np.float64, 1e-3, 0, interpAlgo=1)


@pytest.mark.parametrize("algo", [0, 1])
@pytest.mark.parametrize("direction", [0, 1])
@pytest.mark.parametrize("shape", [(100, 333), (129, 257), (300, 70)])
def test_interp2d(shape, direction, algo):
    check(shape, np.float32, 1e-3, 0, interpAlgo=algo, interpDirection=direction)


@pytest.mark.parametrize("algo", [0, 1])
@pytest.mark.parametrize("n", [100000, 4097, 37, 1 << 20])
def test_interp1d(n, algo):
    check((n,), np.float32, 1e-3, 0, interpAlgo=algo)


def test_interp_constant_and_tiny():
    check((50, 50, 50), np.float32, 1e-3, 0, data=np.full((50, 50, 50), 3.25, dtype=np.float32))
    check((1,), np.float32, 1e-3, 0, data=np.array([1.5], dtype=np.float32))
    check((2, 2, 2), np.float64, 1e-3, 0)


@pytest.mark.parametrize("schedule", [1, 4, 5])
def test_interp3d_g3_256(schedule):
    """The benchmark field at 256^3: 16.7 M indices, still seconds for the reference."""
    data = field_g3((256, 256, 256))
    check(data.shape, np.float32, 1e-3, schedule, data=data, interpAlgo=1)


# ---- box schedule (interp_box.cu): finest level by TMA-staged planes; dims must be multiples of 32 -------------------------
BOX_SHAPES = [(64, 64, 64), (32, 64, 96), (96, 64, 128), (128, 128, 160), (128, 128, 128), (256, 128, 256)]


@pytest.mark.parametrize("schedule", [0, 6])
@pytest.mark.parametrize("shape", BOX_SHAPES)
@pytest.mark.parametrize("algo", [0, 1])
def test_interp3d_box(shape, algo, schedule):
    check(shape, np.float32, 1e-2, schedule, interpAlgo=algo, interpDirection=0)


@pytest.mark.parametrize("kw", [dict(interpAlpha=-1.0), dict(interpAlpha=2.0, interpBeta=3.0), dict(interpAnchorStride=64),
                                dict(quantbinCnt=64), dict(quantbinCnt=1024)])
def test_interp3d_box_variants(kw):
    check((64, 96, 64), np.float32, 1e-3, 6, interpAlgo=1, interpDirection=0, **kw)


def test_interp3d_box_many_unpredictable():
    check((64, 64, 96), np.float32, 1e-6, 6, interpAlgo=1, quantbinCnt=16)
    check((64, 64, 96), np.float32, 1e-6, 6, interpAlgo=0, quantbinCnt=16)


@pytest.mark.parametrize("algo", [0, 1])
def test_interp3d_box_special_values(algo):
    data = field_nd((64, 64, 64), np.float32)
    data[3, 4, 5] = np.nan
    data[10, 11, 12] = np.inf
    data[20, 21, 22] = -np.inf
    data[30, 31, 32] = 1e30
    data[33, 32, 63] = np.nan
    data[63, 63, 63] = -np.inf
    check(data.shape, np.float32, 1e-3, 6, data=data, interpAlgo=algo, interpAnchorStride=32)


def test_interp3d_box_rejects_other_shapes():
    """Schedule 6 insists on the box kernel: a shape it does not cover is an error, not a silent fallback."""
    L = product_lib()
    data = field_nd((40, 50, 70), np.float32)
    conf = make_config(data.shape, cmprAlgo=ALGO_INTERP, interpAnchorStride=32)
    q = np.empty(data.size, dtype=np.int32)
    blob = np.empty(data.nbytes + 4096, dtype=np.uint8)
    blen = C.c_size_t(0)
    rc = L.sz3b_interp_decompose(0, C.byref(conf), C.c_double(1e-3), data.ctypes.data_as(C.c_void_p), 0, 6,
                                 q.ctypes.data_as(C.c_void_p), blob.ctypes.data_as(C.c_void_p), C.c_size_t(blob.size), C.byref(blen))
    assert rc != 0


@pytest.mark.parametrize("schedule", [0, 4, 6])
def test_interp3d_g3_512_against_reference(schedule):
    """The headline array (512^3 G3, abs 1e-3, what bench.py times) against the reference itself: every index and the
    decomposition blob, for the automatic schedule, the line walker and the box kernel."""
    data = field_g3((512, 512, 512))
    check(data.shape, np.float32, 1e-3, schedule, data=data, interpAlgo=1, interpDirection=0, interpAlpha=1.0, interpBeta=1.0)


def test_schedules_agree_512():
    """Full headline size: the two independent GPU schedules must produce identical streams (size-independent check)."""
    data = field_g3((512, 512, 512))
    conf = make_config(data.shape, cmprAlgo=ALGO_INTERP, interpAnchorStride=32)
    q1, b1 = gpu_interp(data, conf, 1e-3, 1)
    q2, b2 = gpu_interp(data, conf, 1e-3, 4)
    assert np.array_equal(q1, q2) and b1 == b2
    hist = np.bincount(q1, minlength=65536)
    assert hist.sum() == data.size and hist[0] >= 4096  # anchors are stored as unpredictables


def _huffman_both(q):
    """(ours, reference) = tree blob | size_t outSize | bits for the int32 array q, plus tree lengths."""
    L, R = product_lib(), ref_lib()
    L.sz3b_huffman_encode.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
    R.ref_huffman_encode.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
    cap = q.size * 8 + (1 << 22)
    ours, theirs = np.empty(cap, np.uint8), np.empty(cap, np.uint8)
    n1, t1, t2 = C.c_size_t(0), C.c_size_t(0), C.c_size_t(0)
    rc = L.sz3b_huffman_encode(q.ctypes.data, q.size, 0, ours.ctypes.data, cap, C.byref(n1), C.byref(t1))
    assert rc == 0, L.sz3b_last_error()
    n2 = R.ref_huffman_encode(q.ctypes.data, q.size, theirs.ctypes.data, C.byref(t2))
    assert n2 > 0
    return ours[:n1.value], t1.value, theirs[:n2], t2.value


@pytest.mark.skipif(ref_lib() is None, reason="oracle/_ref/libsz3ref.so not built")
@pytest.mark.parametrize("nsym", [20, 36, 40])
def test_huffman_long_codes(nsym):
    """Fibonacci symbol counts give a maximally skewed tree: code lengths up to nsym - 1 bits (beyond the 12-bit first
    level table of the decoder, and beyond 32 bits for nsym = 36 and 40 -- the case a 2^31-element array reaches on
    its rarest symbols).  Encoder against the reference byte for byte, GPU decoder on both streams."""
    fib = [1, 1]
    while len(fib) < nsym:
        fib.append(fib[-1] + fib[-2])
    scale = max(1, fib[-1] // 30_000_000)          # keep the array below ~80 M symbols
    counts = [max(1, f // scale) for f in fib]
    rng = np.random.default_rng(nsym)
    q = np.repeat(np.arange(nsym, dtype=np.int32) * 3 + 1000, counts)
    rng.shuffle(q)
    ours, t1, theirs, t2 = _huffman_both(q)
    assert t1 == t2 and ours.size == theirs.size and np.array_equal(ours, theirs)
    L = product_lib()
    L.sz3b_huffman_decode.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_void_p]
    back = np.empty(q.size, np.int32)
    rc = L.sz3b_huffman_decode(theirs.ctypes.data, theirs.size, t2, q.size, back.ctypes.data)
    assert rc == 0, L.sz3b_last_error()
    assert np.array_equal(back, q)
