"""sz3_b200 -- Python host binding of libsz3b200.so (CUDA, sm_100a).

Mirrors the reference's Python interface for this path (tools/pysz/src/pysz/sz.pyx: ``szConfig``, ``sz.compress``,
``sz.decompress``) on top of the C ABI in include/sz3b.h.  Inputs may be NumPy arrays (host) or CUDA torch tensors
(zero-copy: the device pointer is handed to the library).  There is no CPU implementation behind this module: loading
fails loudly when the CUDA library has not been built, and every compute call fails when no GPU is visible.
"""
import ctypes as C
import enum
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "lib", "libsz3b200.so")


class SZ3BError(RuntimeError):
    pass


class _Config(C.Structure):
    _fields_ = [
        ("N", C.c_int32), ("dims", C.c_uint64 * 4), ("cmprAlgo", C.c_int32), ("errorBoundMode", C.c_int32),
        ("absErrorBound", C.c_double), ("relErrorBound", C.c_double), ("psnrErrorBound", C.c_double),
        ("l2normErrorBound", C.c_double), ("openmp", C.c_int32), ("quantbinCnt", C.c_int32),
        ("blockSize", C.c_int32), ("lorenzo", C.c_int32), ("lorenzo2", C.c_int32), ("regression", C.c_int32),
        ("regression2", C.c_int32), ("interpAlgo", C.c_int32), ("interpDirection", C.c_int32),
        ("interpAnchorStride", C.c_int32), ("interpAlpha", C.c_double), ("interpBeta", C.c_double),
        ("dataType", C.c_int32), ("predDim", C.c_int32),
    ]


class szErrorBoundMode(enum.IntEnum):  # enum EB, Config.hpp:54
    ABS = 0
    REL = 1
    PSNR = 2
    L2NORM = 3
    ABS_AND_REL = 4
    ABS_OR_REL = 5


class szAlgorithm(enum.IntEnum):  # enum ALGO, Config.hpp:68
    LORENZO_REG = 0
    INTERP_LORENZO = 1
    INTERP = 2
    NOPRED = 3
    LOSSLESS = 4


_lib = None


def lib():
    """The loaded CUDA library (loads on first use; raises if it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise SZ3BError(f"{_LIB_PATH} is missing: run `make` (or __graft_entry__.build()) first; "
                            "sz3_b200 has no CPU fallback")
        L = C.CDLL(_LIB_PATH)
        L.sz3b_last_error.restype = C.c_char_p
        L.sz3b_version.restype = C.c_char_p
        L.sz3b_compress_bound.restype = C.c_size_t
        L.sz3b_config_save.restype = C.c_size_t
        L.sz3b_omp_header_size.restype = C.c_size_t
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        msg = lib().sz3b_last_error().decode(errors="replace")
        if rc == -1:
            raise ValueError(msg)  # std::invalid_argument in the reference
        raise SZ3BError(f"[{rc}] {msg}")


class szConfig:
    """Python face of SZ3::Config (same attribute names as pysz's szConfig)."""

    def __init__(self, *dims):
        self._c = _Config()
        self.setDims(*(dims or (1,)))

    def setDims(self, *dims):
        # pysz accepts both setDims(a, b, c) and setDims((a, b, c)) (tools/pysz/src/pysz/sz.pyx)
        if len(dims) == 1 and isinstance(dims[0], (tuple, list)):
            dims = tuple(dims[0])
        keep = {k: getattr(self._c, k) for k, _ in _Config._fields_} if self._c.quantbinCnt else None
        arr = (C.c_size_t * len(dims))(*[int(d) for d in dims])
        fresh = _Config()
        _check(lib().sz3b_config_init(C.byref(fresh), len(dims), arr))
        if keep:  # setDims only touches N/dims/num/predDim/blockSize (Config.hpp:161-177)
            for k, v in keep.items():
                if k not in ("N", "dims", "predDim", "blockSize"):
                    setattr(fresh, k, v)
        self._c = fresh
        return self.num_elements

    @property
    def dims(self):
        return tuple(int(self._c.dims[i]) for i in range(self._c.N))

    @property
    def num_elements(self):
        n = 1
        for d in self.dims:
            n *= d
        return n

    def copy(self):
        o = szConfig.__new__(szConfig)
        o._c = _Config.from_buffer_copy(bytes(self._c))
        return o

    def __repr__(self):
        return f"szConfig(dims={self.dims}, num_elements={self.num_elements})"


def _forward(name):
    def g(self):
        return getattr(self._c, name)

    def s(self, v):
        setattr(self._c, name, int(v) if isinstance(getattr(self._c, name), int) else float(v))

    return property(g, s)


for _n, _t in _Config._fields_:
    if _n not in ("N", "dims"):
        setattr(szConfig, _n, _forward(_n))


def _buffer_of(data):
    """(pointer, loc, dtype code, shape, keepalive) of a numpy array or CUDA torch tensor."""
    try:
        import torch
        if isinstance(data, torch.Tensor):
            if not data.is_contiguous():
                data = data.contiguous()
            code = {torch.float32: 0, torch.float64: 1, torch.int32: 7, torch.int64: 9}.get(data.dtype)
            if code is None:
                raise TypeError(f"Unsupported dtype: {data.dtype}. Supported on the GPU path: float32, float64, int32, int64")
            if data.is_cuda:
                # the library reads the buffer on its own streams: order it after what torch has queued on the current
                # stream (a .contiguous() copy above included) -- include/sz3b.h, stream contract
                lib().sz3b_set_caller_stream(C.c_void_p(torch.cuda.current_stream(data.device).cuda_stream), 1)
                return data.data_ptr(), 1, code, tuple(data.shape), data
            data = data.numpy()
    except ImportError:
        pass
    if not isinstance(data, np.ndarray):
        raise TypeError("data must be a numpy.ndarray or a torch.Tensor")
    code = {np.dtype(np.float32): 0, np.dtype(np.float64): 1, np.dtype(np.int32): 7, np.dtype(np.int64): 9}.get(data.dtype)
    if code is None:
        raise TypeError(f"Unsupported dtype: {data.dtype}. Supported on the GPU path: float32, float64, int32, int64")
    if not data.flags["C_CONTIGUOUS"]:
        data = np.ascontiguousarray(data)
    return data.ctypes.data, 0, code, tuple(data.shape), data


class sz:
    """SZ3 compression/decompression on the B200 path (same call shapes as pysz.sz)."""

    @staticmethod
    def compress(data, config, out=None, return_config=False):
        """Returns (compressed uint8 ndarray, ratio).  `out`: optional preallocated uint8 array (e.g. pinned)."""
        if not isinstance(config, szConfig):
            raise TypeError(f"config must be szConfig, got {type(config)}")
        ptr, loc, code, shape, keep = _buffer_of(data)
        config.setDims(*shape)
        L = lib()
        bound = L.sz3b_compress_bound(code, C.byref(config._c))
        if out is None:
            out = np.empty(bound, dtype=np.uint8)
        elif out.nbytes < bound:
            raise ValueError("compressed buffer not large enough")
        size = C.c_size_t(0)
        used = _Config()
        _check(L.sz3b_compress(code, C.byref(config._c), C.c_void_p(ptr), loc, out.ctypes.data_as(C.c_char_p),
                               C.c_size_t(out.nbytes), C.byref(size), C.byref(used)))
        original = int(np.prod(shape)) * (4 if code == 0 else 8)
        res = out[:size.value]
        if return_config:
            uc = szConfig.__new__(szConfig)
            uc._c = used
            return res, original / float(size.value), uc
        return res, original / float(size.value)

    @staticmethod
    def decompress(compressed, dtype, shape, device=None):
        """Returns (array, szConfig).  device=None -> numpy array; device='cuda' -> torch CUDA tensor."""
        compressed = np.ascontiguousarray(np.frombuffer(compressed, dtype=np.uint8))
        dt = np.dtype(dtype)
        code = {np.dtype(np.float32): 0, np.dtype(np.float64): 1, np.dtype(np.int32): 7, np.dtype(np.int64): 9}.get(dt)
        if code is None:
            raise TypeError(f"Unsupported dtype: {dtype}")
        L = lib()
        conf = _Config()
        _check(L.sz3b_peek_config(compressed.ctypes.data_as(C.c_char_p), C.c_size_t(compressed.size), C.byref(conf)))
        n = 1
        for i in range(conf.N):
            n *= conf.dims[i]
        if int(np.prod(shape)) != n:
            raise ValueError(f"shape {shape} does not match the stream ({n} elements)")
        if device is None:
            out = np.empty(shape, dtype=dt)
            ptr, loc = out.ctypes.data, 0
        else:
            import torch
            out = torch.empty(shape, dtype={0: torch.float32, 1: torch.float64, 7: torch.int32, 9: torch.int64}[code], device=device)
            lib().sz3b_set_caller_stream(C.c_void_p(torch.cuda.current_stream(out.device).cuda_stream), 1)
            ptr, loc = out.data_ptr(), 1
        _check(L.sz3b_decompress(code, compressed.ctypes.data_as(C.c_char_p), C.c_size_t(compressed.size),
                                 C.c_void_p(ptr), loc, C.byref(conf)))
        uc = szConfig.__new__(szConfig)
        uc._c = conf
        return out, uc

    @staticmethod
    def last_profile():
        """[(stage, milliseconds, kernel launches)] of the calling thread's last compress/decompress call."""
        L = lib()
        cap = 64
        names = (C.c_char_p * cap)()
        ms = (C.c_double * cap)()
        launches = (C.c_int * cap)()
        n = L.sz3b_last_profile(names, ms, launches, cap)
        return [(names[i].decode(), ms[i], launches[i]) for i in range(min(n, cap))]

    @staticmethod
    def set_lossless_policy(policy):
        """Where the lossless stage over the packed stream runs (include/sz3b.h): 2 = on the GPU as standard zstd frames
        of Huffman-only literal blocks (default), 0 = zstd level 3 on the host like the reference, 1 = adaptive host."""
        lib().sz3b_set_lossless_policy(int(policy))

    @staticmethod
    def get_lossless_policy():
        return int(lib().sz3b_get_lossless_policy())

    @staticmethod
    def set_host_threads(n):
        """Host threads the library may use (zstd workers, concurrent tuner trials); 0 = hardware concurrency."""
        lib().sz3b_set_host_threads(int(n))

    @staticmethod
    def get_host_threads():
        return int(lib().sz3b_get_host_threads())

    @staticmethod
    def set_device_fanout(n):
        """GPUs one compress call with config.openmp > 0 spreads its slabs over (0 = all visible; include/sz3b.h)."""
        lib().sz3b_set_device_fanout(int(n))

    @staticmethod
    def set_host_wait(mode):
        """0 = host threads spin while they wait for the device (default), 1 = they poll and yield the core."""
        lib().sz3b_set_host_wait(int(mode))


__all__ = ["sz", "szConfig", "szErrorBoundMode", "szAlgorithm", "SZ3BError", "lib"]
