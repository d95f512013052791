import ctypes as C, sys
sys.path.insert(0, "tests")
import numpy as np, torch
from common import ALGO_LORENZO_REG, EB_REL, Config, field_g3, make_config, product_lib
L = product_lib()
d3 = field_g3((384, 384, 384), np.float64)
c3 = make_config(d3.shape, cmprAlgo=ALGO_LORENZO_REG, errorBoundMode=EB_REL, relErrorBound=1e-4, lorenzo=0, lorenzo2=0, regression=1)
dev = torch.from_numpy(d3).cuda()
cap = L.sz3b_compress_bound(1, C.byref(c3))
cmp = torch.empty(cap, dtype=torch.uint8).pin_memory()
size = C.c_size_t(0)
assert L.sz3b_compress(1, C.byref(c3), C.c_void_p(dev.data_ptr()), 1, C.c_void_p(cmp.data_ptr()), C.c_size_t(cap), C.byref(size), None) == 0
out = torch.empty(d3.size, dtype=torch.float64, device="cuda")
names, ms, ln = (C.c_char_p * 64)(), (C.c_double * 64)(), (C.c_int * 64)()
c2 = Config()
import time
for r in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    rc = L.sz3b_decompress(1, C.c_void_p(cmp.data_ptr()), C.c_size_t(size.value), C.c_void_p(out.data_ptr()), 1, C.byref(c2))
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) * 1e3
    assert rc == 0, L.sz3b_last_error()
    k = L.sz3b_last_profile(names, ms, ln, 64)
    print(r, f"{dt:.2f} ms", {names[i].decode(): round(ms[i], 3) for i in range(k)})
print("max err", float((out.cpu().numpy().reshape(d3.shape) - d3).__abs__().max()), "bound", c2.absErrorBound)
