"""sz3_b200.sharded -- one process per GPU, outermost-dimension slabs (the reference's OpenMP decomposition,
include/SZ3/api/impl/SZImplOMP.hpp:16-117, with ranks in the role of OpenMP threads).

Rank r of G owns rows [r*d0/G, (r+1)*d0/G) of dims[0] and compresses them as an independent stream.  The path has two
exchanges and both are scalar-sized: the min/max all-reduce that turns a REL/PSNR bound into one shared absolute bound
(:57-68) and the gather of per-slab byte counts / payloads (:93-107).  Rank `dst` assembles the reference's OpenMP
container, which the unmodified reference decoder (SZ_decompress_OMP) reads.

`torch.distributed` is plumbing only (NCCL between GPUs, gloo in the CPU tests); `slab_compress` is the C-ABI call
sz3b_compress_slab unless a test injects a checker-backed stand-in.
"""
import ctypes as C

import numpy as np

from . import _Config, _buffer_of, _check, lib, szConfig


def slab_range(rank, world, d0):
    """Rows of dims[0] owned by `rank` (SZImplOMP.hpp:48-50: lo = tid*d0/nThreads, hi = (tid+1)*d0/nThreads)."""
    return (rank * d0) // world, ((rank + 1) * d0) // world


def core_share(local_rank, local_world, allowed=None):
    """The cores rank `local_rank` of `local_world` ranks on this host should keep to: a contiguous, disjoint share of
    the cores the process may run on (what `mpirun --bind-to core` hands an MPI rank).  Empty when the host has fewer
    cores than ranks."""
    import os
    cores = sorted(os.sched_getaffinity(0)) if allowed is None else sorted(allowed)
    k = len(cores) // max(1, local_world)
    return cores[local_rank * k:(local_rank + 1) * k]


def bind_rank(local_rank, local_world):
    """Binds the calling process to its `core_share` and sizes the library's host pool to it (at most six threads: every
    concurrent tuner trial keeps a thread waiting on its stream, and ranks whose pools fill their whole share stall each
    other's main threads and NCCL proxies -- DESIGN.md section 8).  Call it before the first compression of a rank
    launched one-per-GPU; returns the cores it bound to ([] = left alone)."""
    import os
    mine = core_share(local_rank, local_world)
    if not mine or local_world < 2:
        return []
    os.sched_setaffinity(0, set(mine))
    lib().sz3b_set_host_threads(max(2, min(len(mine), 6)))
    return mine


def _gpu_slab_compress(slab, conf_c, rank, world, value_range):
    """(payload bytes, Config blob bytes) of this rank's slab via sz3b_compress_slab (GPU)."""
    ptr, loc, code, shape, keep = _buffer_of(slab)
    L = lib()
    one = _Config.from_buffer_copy(bytes(conf_c))
    lo, hi = slab_range(rank, world, conf_c.dims[0])
    one.dims[0] = hi - lo
    cap = L.sz3b_compress_bound(code, C.byref(one))
    out = np.empty(cap, dtype=np.uint8)
    size, blob, blob_len = C.c_size_t(0), (C.c_ubyte * 256)(), C.c_size_t(0)
    _check(L.sz3b_compress_slab(code, C.byref(conf_c), rank, world, C.c_void_p(ptr), loc, C.c_double(value_range),
                                out.ctypes.data_as(C.c_char_p), C.c_size_t(cap), C.byref(size), blob, C.byref(blob_len)))
    return out[:size.value].tobytes(), bytes(blob[:blob_len.value])


def _minmax(slab):
    ptr, loc, code, shape, keep = _buffer_of(slab)
    mn, mx = C.c_double(0), C.c_double(0)
    _check(lib().sz3b_minmax(code, C.c_void_p(ptr), loc, C.c_size_t(int(np.prod(shape))), C.byref(mn), C.byref(mx)))
    return mn.value, mx.value


def compress_sharded(slab, config, global_dims, group=None, dst=0, slab_compress=None, minmax=None):
    """Collective over `group`: every rank passes ITS slab (numpy array or CUDA tensor) and the same `config` with
    `global_dims` = dims of the whole array.  Returns the assembled stream (uint8 ndarray) on rank `dst`, None elsewhere.
    """
    import torch
    import torch.distributed as dist
    if not isinstance(config, szConfig):
        raise TypeError("config must be szConfig")
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    conf = config.copy()
    conf.setDims(*global_dims)
    for k in ("cmprAlgo", "errorBoundMode", "absErrorBound", "relErrorBound", "psnrErrorBound", "l2normErrorBound",
              "quantbinCnt", "lorenzo", "lorenzo2", "regression", "regression2", "interpAlgo", "interpDirection",
              "interpAnchorStride", "interpAlpha", "interpBeta"):
        setattr(conf, k, getattr(config, k))
    if world > conf.dims[0]:
        raise ValueError("more ranks than rows in dims[0] (the reference clamps nThreads to dims[0], SZImplOMP.hpp:32-35)")
    lo, hi = slab_range(rank, world, conf.dims[0])
    shape = tuple(slab.shape)
    if shape[0] != hi - lo and not (hi - lo == 1 and len(shape) == len(global_dims) - 1):
        raise ValueError(f"rank {rank} must pass rows [{lo}, {hi}) of dims[0]; got a slab of shape {shape}")
    backend = dist.get_backend(group) if dist.is_initialized() else "none"
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")

    # (1) shared absolute bound: all-reduce of min and max (only when the bound is not absolute already)
    value_range = 0.0
    if conf.errorBoundMode not in (0, 3):   # not ABS / L2NORM
        mn, mx = (minmax or _minmax)(slab)
        t = torch.tensor([-mn, mx], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        value_range = float(t[1].item() + t[0].item())

    # (2) independent slab streams
    payload, blob = (slab_compress or _gpu_slab_compress)(slab, conf._c, rank, world, value_range)

    # (3) sizes -> every rank (the offsets of SZImplOMP.hpp:93-99), payloads -> dst
    sizes = torch.zeros(world, 2, dtype=torch.int64, device=dev)
    mine = torch.tensor([[len(payload), len(blob)]], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_gather_into_tensor(sizes, mine, group=group)
    else:
        sizes = mine
    sizes = sizes.cpu().numpy()
    width = int(sizes.sum(axis=1).max())
    buf = torch.zeros(width, dtype=torch.uint8)
    buf[:len(payload) + len(blob)] = torch.frombuffer(bytearray(payload + blob), dtype=torch.uint8)
    buf = buf.to(dev)
    # `dst` is a rank of `group`; torch.distributed.gather wants the global rank
    gathered = [torch.empty_like(buf) for _ in range(world)] if rank == dst else None
    if world > 1:
        dist.gather(buf, gathered, dst=dist.get_global_rank(group, dst) if group is not None else dst, group=group)
    else:
        gathered = [buf]
    if rank != dst:
        return None
    parts = [g.cpu().numpy() for g in gathered]
    payloads = [parts[r][:sizes[r, 0]].tobytes() for r in range(world)]
    blobs = [parts[r][sizes[r, 0]:sizes[r, 0] + sizes[r, 1]].tobytes() for r in range(world)]
    return assemble_container(conf, payloads, blobs, dtype_code=_buffer_of(slab)[2])


def assemble_container(conf, payloads, blobs, dtype_code=0):
    """int nThreads | Config blob x n | size_t x n | payloads (SZImplOMP.hpp:100-107) inside the outer framing."""
    L = lib()
    n = len(payloads)
    blob_arr = (C.POINTER(C.c_ubyte) * n)(*[C.cast(C.create_string_buffer(b, len(b)), C.POINTER(C.c_ubyte)) for b in blobs])
    pay_keep = [C.create_string_buffer(p, len(p)) for p in payloads]
    pay_arr = (C.c_char_p * n)(*[C.cast(k, C.c_char_p) for k in pay_keep])
    blob_sizes = (C.c_size_t * n)(*[len(b) for b in blobs])
    pay_sizes = (C.c_size_t * n)(*[len(p) for p in payloads])
    L.sz3b_omp_header_size.restype = C.c_size_t
    cap = L.sz3b_omp_header_size(n, blob_sizes) + sum(len(p) for p in payloads) + 1024
    out = np.empty(cap, dtype=np.uint8)
    size = C.c_size_t(0)
    _check(L.sz3b_omp_assemble(dtype_code, C.byref(conf._c), n, blob_arr, blob_sizes, pay_sizes, pay_arr,
                               out.ctypes.data_as(C.c_char_p), C.c_size_t(cap), C.byref(size)))
    return out[:size.value].copy()
