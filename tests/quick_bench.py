import sys, time, ctypes as C, numpy as np
sys.path.insert(0, 'tests'); sys.path.insert(0, '.')
import torch
from common import *
import sz3_b200
from sz3_b200 import sz, szConfig
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
t=time.time(); data = field_g3((n,n,n)); print('gen', time.time()-t, flush=True)
for algo in (2, 1):
    conf = szConfig(n,n,n); conf.cmprAlgo = algo; conf.absErrorBound = 1e-3
    d_dev = torch.from_numpy(data).cuda()
    pinned = torch.from_numpy(data).pin_memory()
    for name, src in (('device', d_dev), ('pinned', pinned), ('pageable', data)):
        for it in range(3):
            torch.cuda.synchronize(); t=time.time()
            cmp, ratio = sz.compress(src, conf)
            torch.cuda.synchronize(); dt=time.time()-t
        print('algo', algo, name, 'time %.2f ms' % (dt*1e3), 'GB/s %.2f' % (data.nbytes/dt/1e9), 'ratio %.3f' % ratio)
        for st in sz.last_profile(): print('    %-24s %8.3f ms  launches %d' % st)
R = ref_lib()
if R is not None:
    c = make_config(data.shape, absErrorBound=1e-3)
    cap = R.ref_size_bound(0, C.byref(c)); out = np.empty(cap, np.uint8)
    t=time.time(); m = R.ref_compress(0, C.byref(c), data.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_char_p), C.c_size_t(cap)); dt=time.time()-t
    print('ref serial: %.2f s, ratio %.3f' % (dt, data.nbytes/m))
    c.openmp = 1
    t=time.time(); m = R.ref_compress(0, C.byref(c), data.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_char_p), C.c_size_t(cap)); dt=time.time()-t
    print('ref omp: %.2f s, ratio %.3f' % (dt, data.nbytes/m))
