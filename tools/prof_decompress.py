#!/usr/bin/env python
"""Profiling driver: sz3b_decompress of a stream of the bench workload (512^3 G3, abs 1e-3), output left in HBM.
usage: prof_decompress.py [reps] [edge]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np
import torch
from common import Config, field_g3, make_config, product_lib

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
n = int(sys.argv[2]) if len(sys.argv) > 2 else 512
L = product_lib()
data = field_g3((n, n, n))
conf = make_config(data.shape, absErrorBound=1e-3)
dev = torch.from_numpy(data).cuda()
cap = L.sz3b_compress_bound(0, C.byref(conf))
cmp = torch.empty(cap, dtype=torch.uint8).pin_memory()
size = C.c_size_t(0)
assert L.sz3b_compress(0, C.byref(conf), C.c_void_p(dev.data_ptr()), 1, C.c_void_p(cmp.data_ptr()), C.c_size_t(cap), C.byref(size), None) == 0
out = torch.empty(data.size, dtype=torch.float32, device="cuda")
names, ms, ln = (C.c_char_p * 64)(), (C.c_double * 64)(), (C.c_int * 64)()
c2 = Config()
import time
for r in range(reps):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    rc = L.sz3b_decompress(0, C.c_void_p(cmp.data_ptr()), C.c_size_t(size.value), C.c_void_p(out.data_ptr()), 1, C.byref(c2))
    assert rc == 0, L.sz3b_last_error()
    k = L.sz3b_last_profile(names, ms, ln, 64)
    print(r, f"{(time.perf_counter() - t0) * 1e3:.2f} ms wall", {names[i].decode(): (round(ms[i], 4), ln[i]) for i in range(k)})
