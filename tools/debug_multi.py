import ctypes as C, os, sys
sys.path.insert(0, "tests")
import numpy as np
from common import *
L = product_lib()
data = field_g3((256, 128, 160))
for nslabs in (2, 5):
    conf = make_config(data.shape, cmprAlgo=ALGO_INTERP_LORENZO, openmp=nslabs, absErrorBound=1e-3)
    cap = L.sz3b_compress_bound(0, C.byref(conf)); out = np.empty(cap, np.uint8); size = C.c_size_t(0)
    for fan in (1, 0, 0):
        L.sz3b_set_device_fanout(fan)
        rc = L.sz3b_compress(0, C.byref(conf), data.ctypes.data_as(C.c_void_p), 0, out.ctypes.data_as(C.c_char_p), C.c_size_t(cap), C.byref(size), None)
        print(nslabs, fan, rc, L.sz3b_last_error(), size.value, flush=True)
