import ctypes as C, sys, time
sys.path.insert(0, "tests")
import numpy as np, torch
from common import Config, field_g3, make_config, product_lib
L = product_lib()
data = field_g3((512, 512, 512))
conf = make_config(data.shape, absErrorBound=1e-3)
cap = L.sz3b_compress_bound(0, C.byref(conf))
out = np.empty(cap, dtype=np.uint8)   # pageable output too: what a plain C++ caller has
size = C.c_size_t(0)
ts = []
for r in range(6):
    t0 = time.perf_counter()
    rc = L.sz3b_compress(0, C.byref(conf), data.ctypes.data_as(C.c_void_p), 0, out.ctypes.data_as(C.c_char_p), C.c_size_t(cap), C.byref(size), None)
    ts.append((time.perf_counter() - t0) * 1e3)
    assert rc == 0
print("pageable in/out 512^3:", " ".join(f"{t:.1f}" for t in ts), "ms; ratio", data.nbytes / size.value)
dec = np.empty_like(data)
c2 = Config()
ts = []
for r in range(5):
    t0 = time.perf_counter()
    rc = L.sz3b_decompress(0, out.ctypes.data_as(C.c_char_p), C.c_size_t(size.value), dec.ctypes.data_as(C.c_void_p), 0, C.byref(c2))
    ts.append((time.perf_counter() - t0) * 1e3)
    assert rc == 0
print("decompress to pageable:", " ".join(f"{t:.1f}" for t in ts), "ms; max err", np.abs(dec - data).max())
