"""Committed golden fixtures (tests/golden/, written by tests/golden/make_golden.py from the unmodified reference):
input array, the stream the reference's SZ_compress wrote for it, and the SHA-256 of what its SZ_decompress returns.

CPU suite: the fixtures are intact; the oracle restatement (oracle/sz3_oracle.c) writes and decodes them byte for byte;
when the prebuilt reference is present it still writes them (the fixtures are current).
GPU suite: the CUDA path writes the same bytes ("stream" cases, tuner decisions included) and decodes the reference's
streams to the same bits, through the C ABI.  None of this needs /root/reference or oracle/_ref at run time."""
import ctypes as C
import hashlib
import json
import os

import numpy as np
import pytest

from common import ROOT, Config, dtype_code, make_config, port_lib, product_lib, ref_lib

GOLDEN = os.path.join(ROOT, "tests", "golden")
with open(os.path.join(GOLDEN, "cases.json")) as _f:
    CASES = json.load(_f)
IDS = [c["name"] for c in CASES]


def load(case):
    z = np.load(os.path.join(GOLDEN, case["name"] + ".npz"))
    return np.ascontiguousarray(z["data"]), np.ascontiguousarray(z["stream"])


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("case", CASES, ids=IDS)
def test_fixture_intact(case):
    data, stream = load(case)
    assert list(data.shape) == case["shape"] and data.dtype.name == case["dtype"]
    assert stream.size == case["stream_bytes"] and sha(stream) == case["stream_sha256"]


@pytest.mark.parametrize("case", CASES, ids=IDS)
def test_oracle_port_writes_and_decodes_golden(case):
    P = port_lib()
    data, stream = load(case)
    conf = make_config(data.shape, **case["config"])
    if case["config"].get("openmp"):   # container: as many slabs as the reference had threads when it wrote the fixture
        conf.openmp = int.from_bytes(stream[16:20].tobytes(), "little")
    out = np.empty(stream.size + (1 << 20), np.uint8)   # tuned cases included: the restatement has the auto-tuner
    n = P.orc_compress(dtype_code(data), C.byref(conf), data.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_char_p),
                       C.c_size_t(out.size))
    assert n == stream.size and np.array_equal(out[:n], stream)
    dec, dconf = np.empty_like(data), Config()
    assert P.orc_decompress(dtype_code(data), stream.ctypes.data_as(C.c_char_p), C.c_size_t(stream.size),
                            dec.ctypes.data_as(C.c_void_p), C.byref(dconf)) == 0
    assert sha(dec) == case["decoded_sha256"]
    assert dconf.cmprAlgo == case["algo_in_stream"]
    if case["abs_error_bound"] > 0:
        assert np.max(np.abs(dec.astype(np.float64) - data.astype(np.float64))) <= case["abs_error_bound"]
    else:
        assert np.array_equal(dec, data)


@pytest.mark.skipif(ref_lib() is None, reason="oracle/_ref/libsz3ref.so not built")
@pytest.mark.parametrize("case", CASES, ids=IDS)
def test_reference_still_writes_golden(case):
    from test_gpu_compress import ref_compress
    if case["config"].get("openmp"):
        pytest.skip("the slab count of the container follows the thread count of the box")
    data, stream = load(case)
    theirs = ref_compress(data, make_config(data.shape, **case["config"]))
    assert theirs.size == stream.size and np.array_equal(theirs, stream), "regenerate: python tests/golden/make_golden.py"


@pytest.mark.gpu
@pytest.mark.parametrize("case", [c for c in CASES if c["gpu"] == "stream"], ids=[c["name"] for c in CASES if c["gpu"] == "stream"])
def test_gpu_writes_golden_stream(case):
    from test_gpu_compress import gpu_compress
    data, stream = load(case)
    ours, used = gpu_compress(data, make_config(data.shape, **case["config"]))
    assert ours.size == stream.size and np.array_equal(ours, stream), (ours.size, stream.size)


@pytest.mark.gpu
@pytest.mark.parametrize("case", [c for c in CASES if c["gpu"] == "decode"], ids=[c["name"] for c in CASES if c["gpu"] == "decode"])
def test_gpu_decodes_golden_stream(case):
    L = product_lib()
    L.sz3b_last_error.restype = C.c_char_p
    data, stream = load(case)
    out, conf = np.empty_like(data), Config()
    rc = L.sz3b_decompress(dtype_code(data), stream.ctypes.data_as(C.c_char_p), C.c_size_t(stream.size), out.ctypes.data_as(C.c_void_p),
                           0, C.byref(conf))
    assert rc == 0, L.sz3b_last_error()
    assert sha(out) == case["decoded_sha256"]
