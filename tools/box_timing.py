import ctypes as C, os, sys
sys.path.insert(0, "tests")
import numpy as np, torch
from common import ALGO_INTERP, field_g3, make_config, product_lib
L = product_lib()
for n in (32, 512):
    data = field_g3((n, n, n))
    conf = make_config(data.shape, cmprAlgo=ALGO_INTERP, interpAlgo=1, interpDirection=0, interpAlpha=1.0, interpBeta=1.0, interpAnchorStride=32)
    dev = torch.from_numpy(data).cuda()
    q = np.empty(data.size, dtype=np.int32); blob = np.empty(data.nbytes + 4096, dtype=np.uint8); blen = C.c_size_t(0)
    for r in range(3):
        rc = L.sz3b_interp_decompose(0, C.byref(conf), C.c_double(1e-3), C.c_void_p(dev.data_ptr()), 1, 6, q.ctypes.data_as(C.c_void_p), blob.ctypes.data_as(C.c_void_p), C.c_size_t(blob.size), C.byref(blen))
        assert rc == 0, L.sz3b_last_error()
        t = (C.c_ulonglong * 64)()
        L.sz3b_debug_box_timing(t)
        t = list(t)[:10]
        names = ["start", "fill issued", "fill done+sync", "pass0 done", "sync", "plane: wait done", "merge+pass1", "pass2", "copy/plane_out", "end"]
        print(n, r, " ".join(f"{names[i]}={t[i]-t[0]}" for i in range(10)))
