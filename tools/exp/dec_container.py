import ctypes as C, sys, time
sys.path.insert(0, "tests")
import numpy as np, torch
from common import Config, field_g3, make_config, product_lib
L = product_lib()
data = field_g3((512, 512, 512))
for nsl in (1, 4, 8):
    conf = make_config(data.shape, absErrorBound=1e-3, openmp=nsl if nsl > 1 else 0)
    cap = L.sz3b_compress_bound(0, C.byref(conf))
    cmp = torch.empty(cap, dtype=torch.uint8).pin_memory()
    pin = torch.from_numpy(data).pin_memory()
    size = C.c_size_t(0)
    assert L.sz3b_compress(0, C.byref(conf), C.c_void_p(pin.data_ptr()), 0, C.c_void_p(cmp.data_ptr()), C.c_size_t(cap), C.byref(size), None) == 0, L.sz3b_last_error()
    out = torch.empty(data.size, dtype=torch.float32).pin_memory()
    c2 = Config()
    ts = []
    for r in range(5):
        t0 = time.perf_counter()
        rc = L.sz3b_decompress(0, C.c_void_p(cmp.data_ptr()), C.c_size_t(size.value), C.c_void_p(out.data_ptr()), 0, C.byref(c2))
        ts.append((time.perf_counter() - t0) * 1e3)
        assert rc == 0, L.sz3b_last_error()
    err = np.abs(out.numpy().reshape(data.shape) - data).max()
    print(f"slabs {nsl}: ratio {data.nbytes/size.value:.3f} decompress to pinned host {min(ts):.2f} ms ({data.nbytes/min(ts)/1e6:.1f} GB/s) max err {err:.2e}")
