"""H2D probe (diagnostics): 512 MiB pinned -> device as one copy, as 8 MiB pieces, and in the plane order of
to_device_planes (even planes as strided 2-D copies, then odd planes in groups of 16)."""
import ctypes as C
import time

import torch

rt = C.CDLL("libcudart.so.12")
rt.cudaMemcpyAsync.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
rt.cudaMemcpy2DAsync.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p]
nz, plane = 512, 512 * 512 * 4
host = torch.empty(nz * plane, dtype=torch.uint8).pin_memory()
dev = torch.empty(nz * plane, dtype=torch.uint8, device="cuda")
st = torch.cuda.Stream()
H2D = 1


def timed(fn, reps=5):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        st.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best * 1e3


def whole():
    rt.cudaMemcpyAsync(dev.data_ptr(), host.data_ptr(), nz * plane, H2D, st.cuda_stream)


def pieces():
    p = 8 << 20
    for off in range(0, nz * plane, p):
        rt.cudaMemcpyAsync(dev.data_ptr() + off, host.data_ptr() + off, p, H2D, st.cuda_stream)


def planes(group_even=8):
    d, h = dev.data_ptr(), host.data_ptr()
    n_even = (nz + 1) // 2
    for k in range(0, n_even, group_even):
        cnt = min(group_even, n_even - k)
        rt.cudaMemcpy2DAsync(d + 2 * k * plane, 2 * plane, h + 2 * k * plane, 2 * plane, plane, cnt, H2D, st.cuda_stream)
    for b in range((nz - 1 + 31) // 32):
        z0, z1 = b * 32 + 1, min((b + 1) * 32, nz - 1)
        cnt = (z1 - z0) // 2 + 1
        rt.cudaMemcpy2DAsync(d + z0 * plane, 2 * plane, h + z0 * plane, 2 * plane, plane, cnt, H2D, st.cuda_stream)


def planes_1d():
    d, h = dev.data_ptr(), host.data_ptr()
    for z in list(range(0, nz, 2)) + list(range(1, nz, 2)):
        rt.cudaMemcpyAsync(d + z * plane, h + z * plane, plane, H2D, st.cuda_stream)


for name, fn in (("one copy", whole), ("8 MiB pieces", pieces), ("plane order, 2-D copies", planes),
                 ("plane order, 2-D copies of 32", lambda: planes(32)), ("plane order, 1-D copies per plane", planes_1d)):
    ms = timed(fn)
    print(f"{name:36s} {ms:7.3f} ms  {nz * plane / ms / 1e6:6.1f} GB/s")
