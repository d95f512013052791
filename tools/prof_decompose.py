#!/usr/bin/env python
"""Profiling driver: a few sz3b_interp_decompose calls on the benchmark field (512^3 G3, abs 1e-3, the tuned
parameters of C2) so that ncu sees only the predict+quantize kernels.  usage: prof_decompose.py [schedule] [reps] [n]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np
import torch
from common import ALGO_INTERP, field_g3, make_config, product_lib

schedule = int(sys.argv[1]) if len(sys.argv) > 1 else 0
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
n = int(sys.argv[3]) if len(sys.argv) > 3 else 512
L = product_lib()
data = field_g3((n, n, n))
conf = make_config(data.shape, cmprAlgo=ALGO_INTERP, interpAlgo=1, interpDirection=0, interpAlpha=1.0, interpBeta=1.0,
                   interpAnchorStride=32)
dev = torch.from_numpy(data).cuda()
q = np.empty(data.size, dtype=np.int32)
blob = np.empty(data.nbytes + 4096, dtype=np.uint8)
blen = C.c_size_t(0)
names = (C.c_char_p * 64)()
ms = (C.c_double * 64)()
ln = (C.c_int * 64)()
for r in range(reps):
    rc = L.sz3b_interp_decompose(0, C.byref(conf), C.c_double(1e-3), C.c_void_p(dev.data_ptr()), 1, schedule,
                                 q.ctypes.data_as(C.c_void_p), blob.ctypes.data_as(C.c_void_p), C.c_size_t(blob.size),
                                 C.byref(blen))
    assert rc == 0, L.sz3b_last_error()
    k = L.sz3b_last_profile(names, ms, ln, 64)
    print(r, {names[i].decode(): round(ms[i], 4) for i in range(k)})
