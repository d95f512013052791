"""CPU-only checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/sz3b.h declares, its
host-side pieces (Config blob, size bound) agree with the reference, and compute calls fail loudly without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from common import ROOT, Config, EB_ABS, EB_PSNR, EB_REL, EB_ABS_AND_REL, make_config, product_lib, ref_lib


def declared_functions():
    src = open(os.path.join(ROOT, "include", "sz3b.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sz3b_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = product_lib()
    assert L is not None, "sz3_b200/lib/libsz3b200.so missing: run __graft_entry__.build()"
    names = declared_functions()
    assert len(names) >= 15
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/sz3b.h but not exported"


def test_version_and_device_count():
    L = product_lib()
    assert L.sz3b_version().decode() == "3.3.2"
    assert L.sz3b_device_count() >= 0


def test_config_init_drops_unit_dims():
    L = product_lib()
    c = Config()
    dims = (C.c_size_t * 4)(1, 20, 1, 30)
    assert L.sz3b_config_init(C.byref(c), 4, dims) == 0
    assert c.N == 2 and c.dims[0] == 20 and c.dims[1] == 30 and c.blockSize == 16
    assert c.cmprAlgo == 1 and c.quantbinCnt == 65536 and c.interpAlgo == 1
    dims5 = (C.c_size_t * 5)(2, 2, 2, 2, 2)
    assert L.sz3b_config_init(C.byref(c), 5, dims5) == -1   # std::invalid_argument in the reference


@pytest.mark.skipif(ref_lib() is None, reason="oracle/_ref not built")
@pytest.mark.parametrize("shape", [(512, 512, 512), (100,), (3, 70000), (7, 9, 11, 13), (1 << 20,)])
@pytest.mark.parametrize("mode", [EB_ABS, EB_REL, EB_PSNR, EB_ABS_AND_REL])
def test_config_blob_matches_reference(shape, mode):
    L, R = product_lib(), ref_lib()
    c = make_config(shape, errorBoundMode=mode, absErrorBound=1e-3, relErrorBound=1e-4, psnrErrorBound=80.0,
                    lorenzo2=1, regression=0, openmp=0, quantbinCnt=1024)
    a, b = np.zeros(256, np.uint8), np.zeros(256, np.uint8)
    na = L.sz3b_config_save(C.byref(c), a.ctypes.data_as(C.c_void_p))
    nb = R.ref_config_save(C.byref(c), b.ctypes.data_as(C.c_void_p))
    assert na == nb and bytes(a[:na]) == bytes(b[:nb])
    back = Config()
    assert L.sz3b_config_load(C.byref(back), a.ctypes.data_as(C.c_void_p), C.c_size_t(na)) == 0
    assert back.N == c.N and list(back.dims[:c.N]) == list(c.dims[:c.N])
    assert back.errorBoundMode == mode and back.quantbinCnt == 1024 and back.lorenzo2 == 1 and back.regression == 0


@pytest.mark.skipif(ref_lib() is None, reason="oracle/_ref not built")
@pytest.mark.parametrize("shape,dtype", [((512, 512, 512), 0), ((384, 384, 384), 1), ((1000,), 0), ((16, 64, 128, 128), 0)])
def test_size_bound_matches_reference(shape, dtype):
    L, R = product_lib(), ref_lib()
    c = make_config(shape)
    assert L.sz3b_compress_bound(dtype, C.byref(c)) == R.ref_size_bound(dtype, C.byref(c))


def test_compute_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    L = product_lib()
    data = np.zeros((16, 16, 16), np.float32)
    c = make_config(data.shape)
    cap = L.sz3b_compress_bound(0, C.byref(c))
    out = np.empty(cap, np.uint8)
    size = C.c_size_t(0)
    rc = L.sz3b_compress(0, C.byref(c), data.ctypes.data_as(C.c_void_p), 0, out.ctypes.data_as(C.c_char_p), C.c_size_t(cap),
                         C.byref(size), None)
    assert rc == -3, "no CPU fallback may exist: expected SZ3B_E_CUDA"
    assert b"CUDA" in L.sz3b_last_error() or b"cuda" in L.sz3b_last_error()


def test_host_threads_default_is_the_ranks_share_of_the_cores():
    """One rank per GPU on a shared host: the default pool is hardware concurrency / LOCAL_WORLD_SIZE (at least 2);
    an explicit sz3b_set_host_threads wins.  Read in a fresh process (the default is computed once)."""
    import subprocess
    import sys
    code = ("import ctypes as C, sys; sys.path.insert(0, 'tests'); from common import product_lib; L = product_lib(); "
            "a = L.sz3b_get_host_threads(); L.sz3b_set_host_threads(3); b = L.sz3b_get_host_threads(); "
            "L.sz3b_set_host_threads(0); print(a, b, L.sz3b_get_host_threads())")

    def run(env_extra):
        env = dict(os.environ)
        for k in ("LOCAL_WORLD_SIZE", "OMPI_COMM_WORLD_LOCAL_SIZE"):
            env.pop(k, None)
        env.update(env_extra)
        out = subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=env, capture_output=True, text=True, check=True).stdout
        return [int(x) for x in out.split()]

    cores = min(64, os.cpu_count() or 1)
    alone = run({})
    assert alone[1] == 3 and alone[0] == alone[2] and 1 <= alone[0] <= cores
    shared = run({"LOCAL_WORLD_SIZE": "4"})
    assert shared[0] == (max(2, alone[0] // 4) if alone[0] // 4 >= 1 else 2) and shared[1] == 3 and shared[2] == shared[0]
