#!/usr/bin/env python
"""Stage list + wall time of config #3 (384^3 float64, regression predictor, REL 1e-4) through sz3b_compress."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np, torch
from common import *
L = product_lib()
d = field_g3((384, 384, 384), np.float64)
conf = make_config(d.shape, cmprAlgo=ALGO_LORENZO_REG, errorBoundMode=EB_REL, relErrorBound=1e-4, lorenzo=0, lorenzo2=0, regression=1)
dev = torch.from_numpy(d).cuda()
cap = L.sz3b_compress_bound(1, C.byref(conf)); out = torch.empty(cap, dtype=torch.uint8).pin_memory().numpy(); size = C.c_size_t(0)
names = (C.c_char_p * 64)(); ms = (C.c_double * 64)(); ln = (C.c_int * 64)()
for r in range(5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    rc = L.sz3b_compress(1, C.byref(conf), C.c_void_p(dev.data_ptr()), 1, out.ctypes.data_as(C.c_char_p), C.c_size_t(cap), C.byref(size), None)
    dt = (time.perf_counter() - t0) * 1e3
    assert rc == 0, L.sz3b_last_error()
    k = L.sz3b_last_profile(names, ms, ln, 64)
    print(f"{dt:.2f} ms wall, {size.value} bytes:", " ".join(f"{names[i].decode()}={ms[i]:.2f}" for i in range(k)))
