"""The N > 1 path on CPU: two gloo ranks run sz3_b200.sharded.compress_sharded (slab bounds, min/max all-reduce,
size all-gather, payload gather, container assembly through sz3b_omp_assemble).  The per-slab compression itself needs
a GPU, so the checker (oracle/_ref, the unmodified reference) stands in for it here -- test infrastructure only.
The assembled container must decode with the reference's own OpenMP decoder and honour the bound."""
import ctypes as C
import os
import socket
import struct
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from common import EB_REL, Config, field_g3, make_config, ref_lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(ref_lib() is None, reason="oracle/_ref not built")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _ref_slab_compress(slab, conf_c, rank, world, value_range):
    """SZ_compress_dispatcher on one slab, through the reference: returns (payload, Config blob)."""
    R = ref_lib()
    c = Config.from_buffer_copy(bytes(conf_c))
    c.dims[0] = slab.shape[0]
    c.openmp = 0
    if c.errorBoundMode == EB_REL:       # the shared bound of SZImplOMP.hpp:57-68, already resolved
        c.absErrorBound = c.relErrorBound * np.float32(value_range)
        c.errorBoundMode = 0
    cap = R.ref_size_bound(0, C.byref(c))
    out = np.empty(cap, dtype=np.uint8)
    n = R.ref_compress(0, C.byref(c), slab.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_char_p), C.c_size_t(cap))
    assert n > 0
    payload_len = struct.unpack("<Q", out[8:16].tobytes())[0]
    payload = out[16:16 + payload_len].tobytes()
    blob = bytearray(out[16 + payload_len:n].tobytes())
    return payload, bytes(blob)


def _worker(rank, world, port, mode, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sz3_b200 import szConfig
    from sz3_b200.sharded import compress_sharded, slab_range
    data = field_g3((50, 48, 40))
    lo, hi = slab_range(rank, world, data.shape[0])
    conf = szConfig(*data.shape)
    if mode == "rel":
        conf.errorBoundMode = EB_REL
        conf.relErrorBound = 1e-3
    else:
        conf.absErrorBound = 1e-3
    out = compress_sharded(np.ascontiguousarray(data[lo:hi]), conf, data.shape, slab_compress=_ref_slab_compress,
                           minmax=lambda s: (float(s.min()), float(s.max())))
    if rank == 0:
        ret["stream"] = out.tobytes()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["abs", "rel"])
def test_two_rank_container_decodes_with_reference(mode):
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), mode, ret), nprocs=world, join=True)
    stream = np.frombuffer(ret["stream"], dtype=np.uint8).copy()
    data = field_g3((50, 48, 40))
    R = ref_lib()
    dec = np.empty_like(data)
    conf = Config()
    rc = R.ref_decompress(0, stream.ctypes.data_as(C.c_char_p), C.c_size_t(stream.size), dec.ctypes.data_as(C.c_void_p), C.byref(conf))
    assert rc == 0
    eb = 1e-3 if mode == "abs" else 1e-3 * float(np.float32(data.max()) - np.float32(data.min()))
    assert conf.openmp == 1
    assert np.max(np.abs(dec.astype(np.float64) - data.astype(np.float64))) <= eb * (1 + 1e-6)
    # container header: nThreads == world
    payload = struct.unpack("<Q", stream[8:16].tobytes())[0]
    assert struct.unpack("<i", stream[16:20].tobytes())[0] == world
    assert 16 + payload < stream.size


def test_slab_ranges_match_reference_split():
    from sz3_b200.sharded import slab_range
    for d0, world in [(50, 2), (7, 3), (2048, 8), (5, 5), (100, 7)]:
        edges = [slab_range(r, world, d0) for r in range(world)]
        assert edges[0][0] == 0 and edges[-1][1] == d0
        for (a, b), (c, d) in zip(edges, edges[1:]):
            assert b == c and b > a


def test_core_shares_are_disjoint_and_cover():
    """sharded.core_share: the ranks of a host get contiguous, disjoint, equally sized shares of the allowed cores."""
    from sz3_b200.sharded import core_share
    allowed = list(range(3, 35))           # 32 cores, not starting at 0
    for world in (1, 2, 4, 8):
        shares = [core_share(r, world, allowed) for r in range(world)]
        assert all(len(s) == 32 // world for s in shares)
        flat = [c for s in shares for c in s]
        assert flat == allowed[:len(flat)] and len(set(flat)) == len(flat)
    assert core_share(5, 64, allowed) == []     # fewer cores than ranks: nothing to bind to
