"""One Lorenzo-only 256^3 compression (diagnostics: per-front kernel times under ncu).  usage: python tests/lz_one.py [n] [regression]"""
import ctypes as C
import sys

import numpy as np

sys.path.insert(0, "tests")
sys.path.insert(0, ".")
import torch  # noqa: E402
from common import ALGO_LORENZO_REG, field_g3, make_config, product_lib  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reg = int(sys.argv[2]) if len(sys.argv) > 2 else 0
L = product_lib()
data = field_g3((n, n, n))
conf = make_config(data.shape, cmprAlgo=ALGO_LORENZO_REG, regression=reg, absErrorBound=1e-3)
cap = L.sz3b_compress_bound(0, C.byref(conf))
out = np.empty(cap, np.uint8)
size = C.c_size_t(0)
dev = torch.from_numpy(data).cuda()
for _ in range(2):
    rc = L.sz3b_compress(0, C.byref(conf), C.c_void_p(dev.data_ptr()), 1, out.ctypes.data_as(C.c_char_p), C.c_size_t(cap), C.byref(size), None)
    assert rc == 0, L.sz3b_last_error()
print("ratio", data.nbytes / size.value)
