"""Manual large-shape check (not collected by pytest): one slab of BASELINE config #5 (256 x 2048 x 2048 float32, 4 GiB,
generated on the device), compressed with a device pointer, decompressed by this library on the GPU, bound checked on
the GPU; optionally (--ref) the stream is also decoded by the unmodified reference and compared bit for bit.
With --c4: one slab of BASELINE config #4 instead (32 x 256 x 512 x 512 float32 CESM-like field, 8 GiB, PSNR 80).
usage: python tests/large_check.py [--ref] [--c4] [d0 d1 d2 [d3]]"""
import ctypes as C
import sys
import time

import numpy as np
import torch

sys.path.insert(0, "tests")
sys.path.insert(0, ".")
from common import Config, make_config, product_lib, ref_lib  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
c4 = "--c4" in sys.argv   # 4-D CESM-like field, PSNR mode (otherwise G3, absolute bound)
shape = tuple(int(a) for a in args) if args else ((32, 256, 512, 512) if c4 else (256, 2048, 2048))
if c4 and len(shape) != 4:
    raise SystemExit("--c4 needs four dimensions")
use_ref = "--ref" in sys.argv
dev = torch.device("cuda")
tp = 2 * np.pi
gen = torch.Generator(device=dev)
gen.manual_seed(1234)
if c4:   # G4 of SURVEY.md 8(d), built slice by slice to keep temporaries small
    T_, Z_, Y_, X_ = shape
    data = torch.empty(shape, device=dev, dtype=torch.float32)
    z = torch.arange(Z_, device=dev, dtype=torch.float32)[:, None, None]
    y = torch.arange(Y_, device=dev, dtype=torch.float32)[None, :, None]
    x = torch.arange(X_, device=dev, dtype=torch.float32)[None, None, :]
    for t in range(T_):
        data[t] = (280 + 30 * torch.cos(np.pi * y / Y_) + 5 * torch.sin(tp * x / X_ + 0.2 * t) + 0.1 * z * torch.sin(tp * y / 32)
                   + 0.05 * torch.randn((Z_, Y_, X_), device=dev, generator=gen))
else:
    z = torch.arange(shape[0], device=dev, dtype=torch.float32)[:, None, None]
    y = torch.arange(shape[1], device=dev, dtype=torch.float32)[None, :, None]
    x = torch.arange(shape[2], device=dev, dtype=torch.float32)[None, None, :]
    data = torch.sin(tp * x / 64) * torch.cos(tp * y / 96) + 0.5 * torch.sin(tp * z / 128 + 0.3) + 0.25 * torch.sin(tp * (x + y + z) / 37)
    data += 0.002 * torch.randn(shape, device=dev, generator=gen)
    data = data.contiguous()
torch.cuda.synchronize()
L = product_lib()
L.sz3b_last_error.restype = C.c_char_p
import os  # noqa: E402
if os.environ.get("SZ3B_POLICY"):
    L.sz3b_set_lossless_policy(int(os.environ["SZ3B_POLICY"]))
conf = make_config(shape, errorBoundMode=2, psnrErrorBound=80.0) if c4 else make_config(shape, absErrorBound=1e-3)
bound = None if c4 else 1e-3
cap = L.sz3b_compress_bound(0, C.byref(conf))
out = torch.empty(cap, dtype=torch.uint8).pin_memory().numpy()
size = C.c_size_t(0)
used = Config()
for it in range(2):
    t0 = time.perf_counter()
    rc = L.sz3b_compress(0, C.byref(conf), C.c_void_p(data.data_ptr()), 1, out.ctypes.data_as(C.c_char_p), C.c_size_t(cap), C.byref(size), C.byref(used))
    dt = time.perf_counter() - t0
    assert rc == 0, L.sz3b_last_error()
nbytes = data.numel() * 4
print(f"compress {shape}: {dt*1e3:.1f} ms, {nbytes/dt/1e9:.1f} GB/s device-resident, ratio {nbytes/size.value:.3f}, algo {used.cmprAlgo}, "
      f"interp {used.interpAlgo} dir {used.interpDirection} alpha {used.interpAlpha}", flush=True)
names, ms, launches = (C.c_char_p * 64)(), (C.c_double * 64)(), (C.c_int * 64)()
n = L.sz3b_last_profile(names, ms, launches, 64)
acc = {}
for i in range(n):
    acc[names[i].decode()] = acc.get(names[i].decode(), 0) + ms[i]
print({k: round(v, 2) for k, v in acc.items()}, flush=True)
dec = torch.empty_like(data)
dconf = Config()
t0 = time.perf_counter()
rc = L.sz3b_decompress(0, out.ctypes.data_as(C.c_char_p), C.c_size_t(size.value), C.c_void_p(dec.data_ptr()), 1, C.byref(dconf))
dt = time.perf_counter() - t0
ours_ok = rc == 0
if not ours_ok:
    print("sz3b_decompress FAILED:", L.sz3b_last_error().decode(), flush=True)
    L.sz3b_peek_config(out.ctypes.data_as(C.c_char_p), C.c_size_t(size.value), C.byref(dconf))
if bound is None:
    bound = dconf.absErrorBound   # the resolved absolute bound travels in the stream's Config
if ours_ok:
    err = max(float((dec[i].double() - data[i].double()).abs().max()) for i in range(shape[0]))
    print(f"decompress: {dt*1e3:.1f} ms, max abs error {err:.3e} (bound {bound:.3e})", flush=True)
    for it in range(2):
        t0 = time.perf_counter()
        rc = L.sz3b_decompress(0, out.ctypes.data_as(C.c_char_p), C.c_size_t(size.value), C.c_void_p(dec.data_ptr()), 1, C.byref(dconf))
        dt = time.perf_counter() - t0
    n = L.sz3b_last_profile(names, ms, launches, 64)
    print(f"decompress (warm): {dt*1e3:.1f} ms, {nbytes/dt/1e9:.1f} GB/s", {names[i].decode(): round(ms[i], 2) for i in range(n)}, flush=True)
    assert err <= bound
if use_ref:
    R = ref_lib()
    host = np.empty(shape, np.float32)   # needs the array again in host memory
    rconf = Config()
    t0 = time.perf_counter()
    rc = R.ref_decompress(0, out.ctypes.data_as(C.c_char_p), C.c_size_t(size.value), host.ctypes.data_as(C.c_void_p), C.byref(rconf))
    print(f"reference decoder: rc {rc}, {time.perf_counter()-t0:.1f} s", flush=True)
    assert rc == 0
    rerr = max(float((torch.from_numpy(host[i]).to(dev).double() - data[i].double()).abs().max()) for i in range(shape[0]))
    print(f"reference decoder: max abs error {rerr:.3e} (bound {bound:.3e})", flush=True)
    if ours_ok:
        same = np.array_equal(host.view(np.uint32), dec.cpu().numpy().view(np.uint32))
        print("bit-identical to the reference decoder:", same)
        assert same
    assert rerr <= bound
assert ours_ok
print("ok")
