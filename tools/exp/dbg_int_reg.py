import sys, os
sys.path.insert(0, "tests")
os.environ["SZ3B_INT_REGRESSION"] = "1"
import numpy as np, ctypes as C
from common import *
from test_gpu_compress import gpu_compress, ref_compress, ref_decompress
from test_gpu_int_types import gpu_decompress
for shape in [(40, 36, 50), (12, 12, 12), (6, 6, 6), (12, 6, 6)]:
    g = field_nd(shape, np.float64)
    data = np.ascontiguousarray(np.rint(g * 3000.0).astype(np.int64))
    conf = make_config(shape, cmprAlgo=ALGO_LORENZO_REG, absErrorBound=2.0, lorenzo=0, lorenzo2=0, regression=1, regression2=0)
    ours, _ = gpu_compress(data, conf); theirs = ref_compress(data, conf)
    d1, _ = ref_decompress(theirs, data); d2 = gpu_decompress(theirs, data); d3, _ = ref_decompress(ours, data); d4 = gpu_decompress(ours, data)
    print(shape, ours.size, theirs.size, "gpu-dec(ref)==ref-dec(ref):", np.array_equal(d1, d2), "ref-dec(ours)==gpu-dec(ours):", np.array_equal(d3, d4),
          "dec(ours)==dec(ref):", np.array_equal(d3, d1), "ndiff", int((d3 != d1).sum()), "min", data.min())
    if not np.array_equal(d3, d1):
        idx = np.argwhere(d3 != d1)[:5]; print(idx.tolist(), [ (int(d3[tuple(i)]), int(d1[tuple(i)]), int(data[tuple(i)])) for i in idx])
