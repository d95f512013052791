"""Times the other BASELINE.json configurations through sz3b_compress (device-resident and pinned-host input) next to
the reference's CPU path; prints one JSON line per configuration.  usage: python tests/bench_configs.py [c3] [c4s] ..."""
import ctypes as C
import json
import sys
import time

import numpy as np

sys.path.insert(0, "tests")
sys.path.insert(0, ".")
import torch  # noqa: E402
from common import (ALGO_INTERP_LORENZO, ALGO_LORENZO_REG, EB_PSNR, EB_REL, Config, dtype_code, field_g3, field_g4,  # noqa: E402
                    make_config, product_lib, ref_lib)


def profile(L):
    names = (C.c_char_p * 64)()
    ms = (C.c_double * 64)()
    launches = (C.c_int * 64)()
    n = L.sz3b_last_profile(names, ms, launches, 64)
    acc = {}
    for i in range(min(n, 64)):
        acc[names[i].decode()] = acc.get(names[i].decode(), 0.0) + ms[i]
    return {k: round(v, 3) for k, v in acc.items()}


def run(name, data, conf, reps=3, ref=True):
    L = product_lib()
    L.sz3b_last_error.restype = C.c_char_p
    cap = L.sz3b_compress_bound(dtype_code(data), C.byref(conf))
    out = torch.empty(cap, dtype=torch.uint8).pin_memory().numpy()
    size = C.c_size_t(0)
    used = Config()
    dev = torch.from_numpy(data).cuda()
    pinned = torch.from_numpy(data).pin_memory()
    res = {"config": name, "shape": list(data.shape), "dtype": str(data.dtype), "bytes": data.nbytes}
    for label, ptr, loc in (("device", dev.data_ptr(), 1), ("pinned_host", pinned.data_ptr(), 0)):
        best = None
        for _ in range(reps + 1):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            rc = L.sz3b_compress(dtype_code(data), C.byref(conf), C.c_void_p(ptr), loc, out.ctypes.data_as(C.c_char_p),
                                 C.c_size_t(cap), C.byref(size), C.byref(used))
            dt = time.perf_counter() - t0
            if rc != 0:
                res["error"] = L.sz3b_last_error().decode()
                print(json.dumps(res), flush=True)
                return
            best = dt if best is None else min(best, dt)
        res[label + "_ms"] = round(best * 1e3, 3)
        res[label + "_GBps"] = round(data.nbytes / best / 1e9, 3)
        if label == "device":
            res["stages_ms"] = profile(L)
    res["ratio"] = round(data.nbytes / size.value, 4)
    res["algo_used"] = used.cmprAlgo
    R = ref_lib()
    if ref and R is not None:
        rc_conf = Config.from_buffer_copy(bytes(conf))
        rc_conf.openmp = 1
        rcap = R.ref_size_bound(dtype_code(data), C.byref(rc_conf))
        rout = np.empty(rcap, dtype=np.uint8)
        t0 = time.perf_counter()
        n = R.ref_compress(dtype_code(data), C.byref(rc_conf), data.ctypes.data_as(C.c_void_p), rout.ctypes.data_as(C.c_char_p), C.c_size_t(rcap))
        dt = time.perf_counter() - t0
        res["ref_omp_ms"] = round(dt * 1e3, 1)
        res["ref_omp_GBps"] = round(data.nbytes / dt / 1e9, 3)
        res["ref_omp_ratio"] = round(data.nbytes / n, 4)
    print(json.dumps(res), flush=True)


which = sys.argv[1:] or ["c3", "c4s"]
if "c3" in which:   # config #3: 3-D float64 384^3, regression predictor only, REL 1e-4
    d = field_g3((384, 384, 384), np.float64)
    run("C3 384^3 f64 regression REL 1e-4", d, make_config(d.shape, cmprAlgo=ALGO_LORENZO_REG, lorenzo=0, lorenzo2=0, regression=1,
                                                            errorBoundMode=EB_REL, relErrorBound=1e-4))
if "c4s" in which:  # config #4 at 1/32 size: 4-D float32 (CESM-like), PSNR 80
    d = field_g4((16, 128, 256, 256))
    run("C4/32 16x128x256x256 f32 PSNR 80", d, make_config(d.shape, cmprAlgo=ALGO_INTERP_LORENZO, errorBoundMode=EB_PSNR, psnrErrorBound=80.0))
if "c2d" in which:  # config #2 with the other tuner direction forced
    from common import ALGO_INTERP
    d = field_g3((512, 512, 512))
    run("C2 512^3 f32 ALGO_INTERP dir 5", d, make_config(d.shape, cmprAlgo=ALGO_INTERP, absErrorBound=1e-3, interpDirection=5), ref=False)
if "lz" in which:   # ALGO_LORENZO_REG stacks with a Lorenzo predictor (block wavefront, lorenzo.cu)
    d = field_g3((256, 256, 256))
    run("LZ 256^3 f32 lorenzo ABS 1e-3", d, make_config(d.shape, cmprAlgo=ALGO_LORENZO_REG, regression=0, absErrorBound=1e-3))
    run("LZ 256^3 f32 lorenzo+regression ABS 1e-3", d, make_config(d.shape, cmprAlgo=ALGO_LORENZO_REG, absErrorBound=1e-3))
    run("LZ 256^3 f32 lorenzo+regression ABS 1e-2", d, make_config(d.shape, cmprAlgo=ALGO_LORENZO_REG, absErrorBound=1e-2))
if "lz512" in which:
    d = field_g3((512, 512, 512))
    run("LZ 512^3 f32 lorenzo ABS 1e-3", d, make_config(d.shape, cmprAlgo=ALGO_LORENZO_REG, regression=0, absErrorBound=1e-3))
    run("LZ 512^3 f32 lorenzo+regression ABS 1e-3", d, make_config(d.shape, cmprAlgo=ALGO_LORENZO_REG, absErrorBound=1e-3), reps=1)
