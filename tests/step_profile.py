"""Per-step stage profile of sz3b_compress on the bench workload (diagnostics).  usage: python tests/step_profile.py [edge] [steps]"""
import ctypes as C
import sys
import time

import numpy as np

sys.path.insert(0, "tests")
sys.path.insert(0, ".")
import torch  # noqa: E402
from common import Config, field_g3, make_config, product_lib  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
L = product_lib()
data = field_g3((n, n, n))
conf = make_config(data.shape, absErrorBound=1e-3)
cap = L.sz3b_compress_bound(0, C.byref(conf))
out = torch.empty(cap, dtype=torch.uint8).pin_memory().numpy()
dev = torch.from_numpy(data).cuda()
pinned = torch.from_numpy(data).pin_memory()
size = C.c_size_t(0)
names, ms, nl = (C.c_char_p * 64)(), (C.c_double * 64)(), (C.c_int * 64)()
pols = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [2, 0]
for pol in pols:
    L.sz3b_set_lossless_policy(pol)
    for label, ptr, loc in (("device", dev.data_ptr(), 1), ("pinned", pinned.data_ptr(), 0)):
        for it in range(steps):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            rc = L.sz3b_compress(0, C.byref(conf), C.c_void_p(ptr), loc, out.ctypes.data_as(C.c_char_p), C.c_size_t(cap), C.byref(size), None)
            dt = (time.perf_counter() - t0) * 1e3
            assert rc == 0, L.sz3b_last_error()
            k = L.sz3b_last_profile(names, ms, nl, 64)
            prof = " ".join(f"{names[i].decode()}={ms[i]:.2f}" for i in range(min(k, 64)) if not names[i].decode().startswith("tune_") or names[i].decode() == "tune_overlapped_with_h2d")
            print(f"policy {pol} {label:7s} step {it}: {dt:7.2f} ms  {prof}", flush=True)
