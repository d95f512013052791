#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric: compress GB/s + ratio, 3-D float32 512^3, abs error bound 1e-3, on N B200s.

    python bench.py --gpus N --steps K --warmup W            # the CUDA library (libsz3b200.so) through its C ABI
    python bench.py --impl reference --gpus N ...            # the reference's own OpenMP CPU path (oracle/_ref)

One "step" = one pass of the whole hot path (tuner -> predict+quantize -> histogram/Huffman -> bit pack -> lossless
stage) over one 512x512x512 float32 array per GPU.  For N > 1 the workload is an (N*512)x512x512 array cut into N
outermost-dimension slabs (the reference's OpenMP decomposition, api/impl/SZImplOMP.hpp:43-108): rank r compresses
slab r, the ranks all-gather their byte counts, and every rank's frames cross PCIe straight to their final offset in
ONE shared container (a pinned shared-memory segment; rank 0 writes the header and the trailing Config) -- all inside
the timed region.  After timing, the unmodified reference decodes that container and every rank checks its slab
against the bound.  Weak scaling: work per GPU is fixed.

  value     whole-job GB/s with the input already resident in HBM (compressed stream / container delivered to host)
  e2e       same, input in pinned HOST memory: H2D of the array and D2H of the stream inside the timed region
  roofline  the fused predict+quantize launches (compact lattices + anchors + k_interp_box per level): algorithmic
            bytes N*(sizeof(T)+4) / their device time (CUDA events recorded by the library on its own stream),
            against MEASURED_PEAKS.json's HBM copy peak
  cpu_baseline / --impl reference: oracle/_ref (the unmodified reference, conf.openmp = true, OpenMP team = all host
            cores, set explicitly: launchers export OMP_NUM_THREADS=1) on the same array(s)
  extras    host_zstd_policy (the same two measurements with the reference's own host zstd call), e2e_two_callers,
            other_configs (BASELINE.json configs #3 and one slab of #5: time, ratio, parity against the reference),
            stages_ms (device / host stage times of the last timed `value` step); --diag prints per-rank details
"""
import argparse
import ctypes as C
import json
import os
import platform
import struct
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

EDGE = 512
EB = 1e-3
METRIC = "compress throughput, 3D f32 512^3 abs-eb 1e-3"


def workload_name(edge, world):
    s = f"3D float32 {edge}x{edge}x{edge} per GPU, ALGO_INTERP_LORENZO abs-eb 1e-3"
    if world > 1:
        s += f", {world * edge}x{edge}x{edge} slab-sharded over {world} GPUs (OpenMP container)"
    return s


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--edge", type=int, default=EDGE, help="cube edge (default 512 = the metric's configuration)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip other_configs / two-callers / host-zstd extras")
    ap.add_argument("--diag", action="store_true", help="per-rank step times and host-link rate on stderr")
    ap.add_argument("--no-pin", action="store_true", help="N > 1: do not bind this rank to its share of the host cores")
    ap.add_argument("--no-numa", action="store_true", help="N > 1: leave this rank's host buffers wherever the kernel puts them")
    ap.add_argument("--host-threads", type=int, default=None, help="host threads per rank (default: this rank's share of the cores)")
    ap.add_argument("--host-wait", type=int, default=None, help="0 = spin while waiting for the device, 1 = poll and yield")
    ap.add_argument("--lossless-policy", type=int, default=None, help="0 = host zstd on every chunk, 1 = adaptive host zstd, 2 = GPU lossless stage (library default)")
    return ap.parse_args()


def slab_field(rank, edge):
    """Slab `rank` of the synthetic field G3 extended along z; float32, seeded."""
    from common import field_g3
    if rank == 0:
        return field_g3((edge, edge, edge))
    z = (np.arange(edge, dtype=np.float32) + np.float32(rank * edge))[:, None, None]
    y = np.arange(edge, dtype=np.float32)[None, :, None]
    x = np.arange(edge, dtype=np.float32)[None, None, :]
    tp = np.float32(2 * np.pi)
    a = (np.sin(tp * x / np.float32(64)) * np.cos(tp * y / np.float32(96)) + np.float32(0.5) * np.sin(tp * z / np.float32(128) + np.float32(0.3))
         + np.float32(0.25) * np.sin(tp * (x + y + z) / np.float32(37)))
    noise = np.random.default_rng(1234 + rank).standard_normal((edge, edge, edge), dtype=np.float32)
    return np.ascontiguousarray((a + np.float32(0.002) * noise).astype(np.float32))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md, clocks line)."""

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for k, nm in enumerate(names):
                    if r[4 + k].lower().startswith("active"):
                        reasons.add(nm)
            except (ValueError, IndexError):
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic():
    """dram bytes per predict+quantize step from a committed ncu --set full capture: (bytes, where it came from).
    A capture of an earlier build, not of this run -- the label says which."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            d = json.load(f)
            return d.get("predict_quantize_dram_bytes"), d.get("source", "profiles/traffic.json")
    except Exception:
        return None, None


def make_conf(edge, **kw):
    from common import ALGO_INTERP_LORENZO, make_config
    return make_config((edge, edge, edge), cmprAlgo=ALGO_INTERP_LORENZO, absErrorBound=EB, **kw)


# ----------------------------------------------------------------------------------------------------------------------
# CPU side: the unmodified reference (oracle/_ref) or, if that library is absent, the C restatement (oracle/)
# ----------------------------------------------------------------------------------------------------------------------
def cpu_checker():
    from common import port_lib, ref_lib
    lib = ref_lib()
    if lib is not None:
        return lib, "ref", "reference"
    lib = port_lib()
    if lib is not None:
        return lib, "orc", "port"
    return None, None, None


def host_cores():
    return len(os.sched_getaffinity(0))


def prefer_gpu_numa_node(torch, local):
    """Pinned host buffers of this rank on the NUMA node its GPU hangs off: set_mempolicy(MPOL_PREFERRED) before
    anything is allocated.  A hint only (the kernel falls back to other nodes), and a no-op on single-node hosts or
    where the container forbids the call.  With eight ranks copying at once the host side of the links is what bounds
    e2e; buffers on the far socket make every DMA cross the inter-socket link as well."""
    try:
        p = torch.cuda.get_device_properties(local)
        bus = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()]
        if node < 0 or len(nodes) < 2:
            return {"gpu_node": node, "host_nodes": len(nodes), "policy": "default"}
        libc = C.CDLL(None, use_errno=True)
        mask = (C.c_ulong * 4)()
        mask[node // 64] = 1 << (node % 64)
        rc = libc.syscall(238, 1, mask, 257)   # x86-64 SYS_set_mempolicy, MPOL_PREFERRED
        return {"gpu_node": node, "host_nodes": len(nodes), "policy": "preferred" if rc == 0 else f"default (errno {C.get_errno()})"}
    except Exception as ex:   # placement is an optimisation, never a reason to fail
        return {"policy": "default", "why": str(ex)[:80]}


def set_ref_threads(lib, prefix, n):
    """OpenMP team of the reference's conf.openmp path, set explicitly (torchrun exports OMP_NUM_THREADS=1).
    Returns the count in force."""
    if prefix == "ref" and hasattr(lib, "ref_set_threads"):
        lib.ref_set_threads(int(n))
        return int(lib.ref_get_max_threads())
    return 1


class _StdoutToStderr:
    """The reference printf()s a line per OpenMP call; keep stdout for the one JSON line."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *a):
        os.dup2(self.saved, 1)
        os.close(self.saved)


def cpu_compress_time(lib, prefix, data, conf, reps):
    with _StdoutToStderr():
        cap = getattr(lib, prefix + "_size_bound")(0 if data.dtype == np.float32 else 1, C.byref(conf))
        out = np.empty(cap, dtype=np.uint8)
        times, size = [], 0
        for _ in range(reps):
            t0 = time.perf_counter()
            size = getattr(lib, prefix + "_compress")(0 if data.dtype == np.float32 else 1, C.byref(conf),
                                                      data.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_char_p), C.c_size_t(cap))
            times.append(time.perf_counter() - t0)
            assert size > 0, "reference compression failed"
        return times, size, out


def cpu_decompress(lib, prefix, cmp_ptr, cmp_size, out_arr):
    from common import Config
    conf = Config()
    with _StdoutToStderr():
        rc = getattr(lib, prefix + "_decompress")(0 if out_arr.dtype == np.float32 else 1, C.c_char_p(cmp_ptr) if isinstance(cmp_ptr, bytes) else C.c_void_p(cmp_ptr),
                                                  C.c_size_t(cmp_size), out_arr.ctypes.data_as(C.c_void_p), C.byref(conf))
    return rc, conf


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    lib, prefix, kind = cpu_checker()
    if lib is None:
        print(json.dumps({"impl": "reference", "unavailable": "neither oracle/_ref/libsz3ref.so nor oracle/libsz3oracle.so is built"}))
        return
    cores = host_cores()
    threads = set_ref_threads(lib, prefix, cores)
    edge, world = args.edge, max(1, args.gpus)
    # the same array the GPU arm compresses: N slabs of edge^3 stacked along z
    data = slab_field(0, edge) if world == 1 else np.concatenate([slab_field(r, edge) for r in range(world)], axis=0)
    from common import ALGO_INTERP_LORENZO, make_config
    conf = make_config(data.shape, cmprAlgo=ALGO_INTERP_LORENZO, absErrorBound=EB, openmp=1)
    cpu_compress_time(lib, prefix, data, conf, max(args.warmup, 1) if world == 1 else 1)
    times, size, _ = cpu_compress_time(lib, prefix, data, conf, args.steps)
    total = sum(times)
    gbs = data.nbytes * args.steps / total / 1e9
    sample = (f"the whole workload per step: {data.shape[0]}x{edge}x{edge} float32 ({world} slab(s) of the GPU arm), "
              f"SZ_compress with conf.openmp=true, OpenMP team = {threads} threads on {cores} cores")
    line = {
        "impl": "reference", "metric": METRIC, "value": gbs, "unit": "GB/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "ratio": data.nbytes / size,
        "config": {"workload": workload_name(edge, world), "field": "G3 (SURVEY.md 8d), seeded"},
        "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------------------------
# GPU side
# ----------------------------------------------------------------------------------------------------------------------
PLACE_FN = C.CFUNCTYPE(C.c_void_p, C.c_void_p, C.c_size_t)


class SharedContainer:
    """One pinned shared-memory segment that all ranks of the node map: the OpenMP container of a step is assembled in
    it (SZImplOMP.hpp:93-107: int nThreads | Config x n | size_t x n | payloads, inside the outer framing of sz.hpp)."""

    def __init__(self, torch, dist, rank, world, nbytes, tag):
        self.path = f"/dev/shm/sz3b_bench_{os.environ.get('MASTER_PORT', '0')}_{tag}"
        self.rank, self.nbytes = rank, nbytes
        if rank == 0:
            with open(self.path, "wb") as f:
                f.truncate(nbytes)
        dist.barrier()
        self.map = np.memmap(self.path, dtype=np.uint8, mode="r+", shape=(nbytes,))
        self.ptr = self.map.ctypes.data
        rc = torch.cuda.cudart().cudaHostRegister(self.ptr, nbytes, 0)
        self.registered = int(rc) == 0 if not isinstance(rc, tuple) else int(rc[0]) == 0
        self.torch = torch
        dist.barrier()

    def close(self, dist):
        if self.registered:
            self.torch.cuda.cudart().cudaHostUnregister(self.ptr)
        del self.map
        dist.barrier()
        if self.rank == 0:
            try:
                os.unlink(self.path)
            except OSError:
                pass


def run_ours(args):
    import torch
    import torch.distributed as dist
    from common import Config, product_lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device visible; the product has no CPU path (use --impl reference for the CPU arm)")
    numa = prefer_gpu_numa_node(torch, local) if (world > 1 and platform.machine() == "x86_64" and not args.no_numa) else None
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = product_lib()
    if L is None:
        raise SystemExit("bench.py: sz3_b200/lib/libsz3b200.so missing; run `python -c 'import __graft_entry__ as g; g.build()'`")
    L.sz3b_last_error.restype = C.c_char_p
    L.sz3b_slab_conf_blob_size.restype = C.c_size_t
    if args.lossless_policy is not None:
        L.sz3b_set_lossless_policy(args.lossless_policy)
    # one rank per GPU on one host: every rank gets its share of the cores (16 pool threads per rank on 2 cores per
    # rank cost half of e2e; tests/gpu_cores.sh).  Waiting threads keep spinning: yielding measured slower.
    cores = host_cores()
    share = max(1, cores // world)
    pinned_cores = None
    k = int(os.environ.get("SZ3B_BENCH_PIN_CORES") or 0) or share   # (experiments: a smaller share than the host offers)
    if world > 1 and not args.no_pin:
        # One rank per GPU: every rank binds itself to its own share of the cores (what `mpirun --bind-to` / numactl do
        # for an MPI job) and sizes its pool to that share.  The pool threads spin while they wait for the device, so
        # ranks that roam over each other's cores stall one another: unbound, 12 threads per rank on 24 cores measured
        # 6.1-6.9 / 14.4 ms per step (value / e2e) with stalls of up to 90 ms; bound to disjoint shares 5.5 / 13.3 ms
        # (tests/gpu_2gpu_probe.sh).
        allowed = sorted(os.sched_getaffinity(0))
        mine_cores = allowed[local * k:(local + 1) * k]
        if len(mine_cores) == k:
            os.sched_setaffinity(0, set(mine_cores))
            pinned_cores = [mine_cores[0], mine_cores[-1]]
            share = k
    # (pool size: every trial compression of the tuner has a thread spinning on its stream; 12 of them per rank measured
    #  6.1 / 14.4-15.3 ms per step with stalls of 25-145 ms even when bound, 6: 5.1 / 13.0 ms, 4 on a 4-core share
    #  5.0-5.8 / 12.7-13.3 ms)
    host_threads = args.host_threads if args.host_threads is not None else (max(2, min(share, 6)) if world > 1 else 0)
    host_wait = args.host_wait if args.host_wait is not None else 0
    L.sz3b_set_host_threads(host_threads)
    L.sz3b_set_host_wait(host_wait)
    L.sz3b_set_device_fanout(1)   # one process per GPU here; the in-process fan-out is measured by extras
    edge = args.edge
    nbytes = edge ** 3 * 4
    host = slab_field(rank, edge)
    pinned = torch.from_numpy(host).pin_memory()
    dev = pinned.cuda(non_blocking=False)

    gconf = make_conf(edge)
    if world > 1:
        gconf.dims[0] = edge * world
    cap = L.sz3b_compress_bound(0, C.byref(make_conf(edge)))
    out = torch.empty(cap, dtype=torch.uint8).pin_memory()
    out_np = out.numpy()
    used = Config()
    blob = (C.c_ubyte * 256)()
    blob_len = C.c_size_t(0)
    size = C.c_size_t(0)

    # ---- N > 1: the shared container ---------------------------------------------------------------------------------
    shared = None
    if world > 1:
        blob_sizes = [int(L.sz3b_slab_conf_blob_size(C.byref(gconf), r, world)) for r in range(world)]
        header = 16 + 4 + sum(blob_sizes) + 8 * world          # outer header | nThreads | Configs | sizes
        slab_room = 160 << 20                                   # far above the ~45 MB a slab takes
        shared = SharedContainer(torch, dist, rank, world, header + world * slab_room + 4096, "c")
        sizes_dev = torch.zeros(world, dtype=torch.int64, device="cuda")
        blobs_dev = torch.zeros(world, 256, dtype=torch.uint8, device="cuda")
        state = {"sizes": None}

        trace = os.environ.get("SZ3B_STEP_TRACE") and rank == 0

        def place(_user, payload_size):
            # the one exchange of the path (SZImplOMP.hpp:93-99): every slab's byte count -> this slab's offset
            t0 = time.perf_counter()
            mine = torch.tensor([payload_size], dtype=torch.int64, device="cuda")
            dist.all_gather_into_tensor(sizes_dev, mine)
            s = sizes_dev.cpu().numpy()
            state["sizes"] = s
            state["t_place"] = time.perf_counter() - t0
            return shared.ptr + header + int(s[:rank].sum())

        place_c = PLACE_FN(place)

    def step(ptr, loc):
        if world == 1:
            rc = L.sz3b_compress(0, C.byref(gconf), C.c_void_p(ptr), loc, out_np.ctypes.data_as(C.c_char_p), C.c_size_t(cap),
                                 C.byref(size), C.byref(used))
            if rc != 0:
                raise RuntimeError(L.sz3b_last_error().decode())
            return size.value
        t0 = time.perf_counter()
        rc = L.sz3b_compress_slab_placed(0, C.byref(gconf), rank, world, C.c_void_p(ptr), loc, C.c_double(0.0), place_c, None,
                                         C.byref(size), blob, C.byref(blob_len))
        if rc != 0:
            raise RuntimeError(L.sz3b_last_error().decode())
        t1 = time.perf_counter()
        # the slab Configs (fixed size, known before the payloads) -> rank 0, which writes the container's header
        mine = torch.zeros(256, dtype=torch.uint8)
        mine[:blob_len.value] = torch.frombuffer(bytearray(bytes(blob[:blob_len.value])), dtype=torch.uint8)
        dist.all_gather_into_tensor(blobs_dev, mine.cuda())
        if rank == 0:
            s = state["sizes"]
            blobs = blobs_dev.cpu().numpy()
            body = struct.pack("<i", world) + b"".join(bytes(blobs[r, :blob_sizes[r]]) for r in range(world)) + \
                struct.pack(f"<{world}Q", *[int(v) for v in s])
            payload = len(body) + int(s.sum())
            head = struct.pack("<IIQ", 0xF342F310, (3 << 24) | (3 << 16) | (2 << 8), payload) + body
            shared.map[:len(head)] = np.frombuffer(head, dtype=np.uint8)
            oc = Config.from_buffer_copy(bytes(gconf))
            oc.openmp = 1
            ob = (C.c_ubyte * 256)()
            n = L.sz3b_config_save(C.byref(oc), ob)
            end = 16 + payload
            shared.map[end:end + n] = np.frombuffer(bytes(ob[:n]), dtype=np.uint8)
            state["total"] = end + n
        if trace:
            print(f"[trace] loc {loc}: slab call {1e3 * (t1 - t0):.2f} ms (of which size exchange {1e3 * state['t_place']:.2f}), "
                  f"header {1e3 * (time.perf_counter() - t1):.2f} ms", file=sys.stderr, flush=True)
        return size.value

    def profile():
        names = (C.c_char_p * 64)()
        ms = (C.c_double * 64)()
        launches = (C.c_int * 64)()
        n = L.sz3b_last_profile(names, ms, launches, 64)
        return [(names[i].decode(), ms[i], launches[i]) for i in range(min(n, 64))]

    local_ms = []
    last_profile = {}
    finest_ms = [0.0]   # the finest interpolation level's launch (recorded inside predict_quantize), timed `value` steps

    def timed(ptr, loc, steps, collect):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        pq_ms, launches, csize = 0.0, 0, 0
        finest_ms[0] = 0.0
        for _ in range(steps):
            csize = step(ptr, loc)
            if collect:
                prof = profile()
                last_profile["stages"] = prof
                for name, ms, nl in prof:
                    launches += nl
                    if name == "predict_quantize":
                        pq_ms += ms
                    if name == "predict_quantize_finest_level":
                        finest_ms[0] += ms
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms_total = e0.elapsed_time(e1)
        local_ms.append(ms_total / steps)
        if world > 1:
            t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_total = float(t.item())
        return ms_total, pq_ms, launches, csize

    warm = max(args.warmup, 3)
    for _ in range(warm):
        step(dev.data_ptr(), 1)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_dev, pq_ms, launches, csize = timed(dev.data_ptr(), 1, args.steps, True)
    finest_total_ms = finest_ms[0]
    for _ in range(warm):
        step(pinned.data_ptr(), 0)
    ms_e2e, _, _, _ = timed(pinned.data_ptr(), 0, args.steps, False)
    h2d, d2h = C.c_size_t(0), C.c_size_t(0)
    L.sz3b_last_transfer(C.byref(h2d), C.byref(d2h))
    clocks = sampler.stop() if rank == 0 else None

    # ---- N > 1: the reference decodes the container of the last step, every rank checks its slab ---------------------
    container_check = None
    if world > 1:
        dist.barrier()
        dec = SharedContainer(torch, dist, rank, world, world * nbytes, "d")
        ok = 1
        if rank == 0:
            lib, prefix, kind = cpu_checker()
            if lib is not None:
                set_ref_threads(lib, prefix, cores)
                arr = np.frombuffer(dec.map, dtype=np.float32)
                rc, dconf = cpu_decompress(lib, prefix, shared.ptr, state["total"], arr)
                ok = 1 if rc == 0 else 0
            else:
                ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        worst = -1.0
        if int(flag.item()) == 1:
            mine = np.frombuffer(dec.map, dtype=np.float32)[rank * edge ** 3:(rank + 1) * edge ** 3].reshape(edge, edge, edge)
            worst = float(np.max(np.abs(mine.astype(np.float64) - host.astype(np.float64))))
        t = torch.tensor([worst], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        container_check = {"decoded_by": "unmodified reference SZ_decompress (oracle/_ref), container of the last e2e step",
                           "ok": bool(int(flag.item()) == 1 and float(t.item()) <= EB), "max_abs_error": float(t.item()),
                           "bytes": int(state["total"]) if rank == 0 else None}
        dec.close(dist)

    if args.diag:   # per-rank view: this rank's own step times and its share of the host link with all ranks copying
        if world > 1:
            dist.barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(5):
            dev.copy_(pinned, non_blocking=True)
        c1.record()
        torch.cuda.synchronize()
        h2d_rate = 5 * nbytes / (c0.elapsed_time(c1) * 1e-3) / 1e9
        print(f"[diag] rank {rank} gpu {local}: value path {local_ms[0]:.2f} ms/step, e2e path {local_ms[1]:.2f} ms/step, "
              f"plain H2D with all ranks copying {h2d_rate:.1f} GB/s, last e2e step: "
              + " ".join(f"{n}={m:.2f}" for n, m, _ in profile()), file=sys.stderr, flush=True)
        if world > 1:
            dist.barrier()

    extras = not args.no_extras
    # the same two measurements with the lossless stage on the host (zstd level 3 on every chunk, the reference's own
    # call), reported next to the headline so that the effect of the GPU lossless stage is visible
    policy = L.sz3b_get_lossless_policy()
    host_zstd = None
    if policy == 2 and extras and world == 1:
        L.sz3b_set_lossless_policy(0)
        for _ in range(3):
            step(dev.data_ptr(), 1)
        ms_dev0, _, _, csize0 = timed(dev.data_ptr(), 1, args.steps, False)
        for _ in range(2):
            step(pinned.data_ptr(), 0)
        ms_e2e0, _, _, _ = timed(pinned.data_ptr(), 0, args.steps, False)
        L.sz3b_set_lossless_policy(2)
        host_zstd = (ms_dev0, ms_e2e0, csize0)

    # Extra (not the headline): the same e2e call issued by two host threads, each with its own output buffer, so that
    # the H2D of one array overlaps the encode / lossless / D2H tail of the other (the library is reentrant: every call
    # borrows its own workspace and streams).  Throughput of a caller that has a queue of arrays to compress.
    two_callers = None
    if world == 1 and args.steps >= 2 and extras:
        try:
            out2 = torch.empty(cap, dtype=torch.uint8).pin_memory().numpy()
            outs = [out_np, out2]
            errs = []

            def caller(k, n):
                sz_k = C.c_size_t(0)
                for _ in range(n):
                    rc = L.sz3b_compress(0, C.byref(make_conf(edge)), C.c_void_p(pinned.data_ptr()), 0,
                                         outs[k].ctypes.data_as(C.c_char_p), C.c_size_t(cap), C.byref(sz_k), None)
                    if rc != 0 or sz_k.value != csize:
                        errs.append((rc, sz_k.value))

            def run(n):
                ths = [threading.Thread(target=caller, args=(k, n)) for k in range(2)]
                for t in ths:
                    t.start()
                for t in ths:
                    t.join()

            run(2)
            per = (args.steps + 1) // 2
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            run(per)
            e1.record()
            torch.cuda.synchronize()
            if not errs:
                two_callers = {"value": nbytes * 2 * per / (e0.elapsed_time(e1) * 1e-3) / 1e9, "unit": "GB/s", "calls": 2 * per,
                               "note": "sz3b_compress from two host threads at once, pinned host input and output; "
                                       "streams byte-identical in size to the single caller's"}
        except Exception as ex:   # never let the extra measurement take the bench line down
            two_callers = {"error": str(ex)[:200]}

    other = None
    decomp = None
    if world == 1 and extras:
        try:
            decomp = decompress_extra(L, torch, dev, make_conf(edge), host, args.steps)
        except Exception as ex:
            decomp = {"error": str(ex)[:300]}
        try:
            other = other_configs(L, torch)
        except Exception as ex:
            other = {"error": str(ex)[:300]}

    total_csize = csize
    if world > 1:
        t = torch.tensor([csize], dtype=torch.int64, device="cuda")
        dist.all_reduce(t)
        total_csize = int(t.item())

    if rank == 0:
        peak, peak_src = measured_peak()
        total_bytes = nbytes * world
        value = total_bytes * args.steps / (ms_dev * 1e-3) / 1e9
        e2e = total_bytes * args.steps / (ms_e2e * 1e-3) / 1e9
        pq_avg_ms = pq_ms / args.steps
        alg_bytes = edge ** 3 * (4 + 4)
        achieved = alg_bytes / (pq_avg_ms * 1e-3) / 1e9 if pq_avg_ms > 0 else 0.0
        traffic, traffic_src = ncu_traffic()
        # the dominant launch alone (the contract's per-launch reading): the finest level = every point with an odd
        # coordinate, N - ceil(edge/2)^3 of them, 8 algorithmic bytes each
        fin_avg_ms = finest_total_ms / args.steps
        fin_bytes = (edge ** 3 - ((edge + 1) // 2) ** 3) * (4 + 4)
        fin_achieved = fin_bytes / (fin_avg_ms * 1e-3) / 1e9 if fin_avg_ms > 0 else None
        line = {
            "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world,
            "steps": args.steps, "warmup": warm, "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "ratio": total_bytes / total_csize,
            "config": {"workload": workload_name(edge, world),
                       "field": "G3 (SURVEY.md 8d), seeded", "l2": "input 512 MiB per step > 126 MB L2 (no explicit flush)",
                       "lossless_policy": {0: "host zstd-3 on every chunk", 1: "adaptive host zstd: probes, raw zstd frames where zstd gains < 1 % (include/sz3b.h)",
                                           2: "GPU lossless stage: zstd frames of Huffman-only literal blocks, one table per 128 KiB (sz3_b200/csrc/zhuf.cuh); decodes with the unmodified reference"}[policy],
                       "host": {"cores": cores, "threads_per_rank": host_threads or cores, "device_wait": ["spin", "yield"][host_wait],
                                **({"rank0_bound_to_cores": pinned_cores} if pinned_cores else {}),
                                **({"numa_rank0": numa} if numa else {})},
                       "value_path": ("sz3b_compress, device-resident input, stream delivered to host" if world == 1 else
                                      "sz3b_compress_slab_placed per rank, device-resident slab; sizes all-gathered, frames delivered "
                                      "to their offsets in one shared pinned container, header written by rank 0 (all timed)"),
                       "e2e_path": ("sz3b_compress, pinned host input (H2D + D2H inside the timed region)" if world == 1 else
                                    "same with pinned host slabs (H2D + D2H + container assembly inside the timed region)")},
            "e2e": {"value": e2e, "unit": "GB/s", "h2d_bytes_per_step": h2d.value * world, "d2h_bytes_per_step": d2h.value * world,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": ("predict_quantize (k_box_compact + k_interp_anchor + k_interp_box x 5 levels; the histogram of "
                                                    "HuffmanEncoder::init is a pass of its own, stage huffman_histogram)"),
                         "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                         "algorithmic_bytes": alg_bytes, "ms_per_step": pq_avg_ms,
                         "timing": "CUDA events recorded by the library on its own stream around the launches, timed `value` steps",
                         "dominant_launch": {"kernel": "k_interp_box<cubic>, finest level (s = 1): one launch, 7/8 of the points",
                                             "algorithmic_bytes": fin_bytes, "ms": fin_avg_ms, "achieved": fin_achieved,
                                             "frac": (fin_achieved / peak) if fin_achieved else None,
                                             "note": "same events, around that launch only; `frac` above is the whole stage "
                                                     "(all levels, anchors, lattice compaction) as in round 1"}},
            "clocks": clocks,
            "stages_ms": {n: round(m, 4) for n, m, _ in _merge(last_profile.get("stages", []))},
        }
        if two_callers is not None:
            line["e2e_two_callers"] = two_callers
        if host_zstd is not None:
            line["host_zstd_policy"] = {
                "value": total_bytes * args.steps / (host_zstd[0] * 1e-3) / 1e9, "e2e": total_bytes * args.steps / (host_zstd[1] * 1e-3) / 1e9,
                "unit": "GB/s", "ratio": nbytes / host_zstd[2], "note": "same run with sz3b_set_lossless_policy(0): zstd level 3 on the host"}
        if container_check is not None:
            line["container_check"] = container_check
        if other is not None:
            line["other_configs"] = other
        if decomp is not None:
            line["decompress"] = decomp
        if world == 1 and not args.no_cpu_baseline:
            lib, prefix, kind = cpu_checker()
            if lib is not None:
                threads = set_ref_threads(lib, prefix, cores)
                times, rsize, _ = cpu_compress_time(lib, prefix, host, make_conf(edge, openmp=1), 3)
                line["cpu_baseline"] = {"value": nbytes / min(times) / 1e9, "unit": "GB/s", "cores": threads, "kind": kind,
                                        "sample": f"the full {edge}^3 array, SZ_compress conf.openmp=true (OpenMP team = {threads} threads on {cores} cores), best of 3",
                                        "ratio": nbytes / rsize}
            else:
                line["cpu_baseline"] = {"value": None, "unit": "GB/s", "cores": 0, "kind": "unavailable", "sample": "no oracle library built"}
        print(json.dumps(line))
    if world > 1:
        shared.close(dist)
        dist.destroy_process_group()


def decompress_extra(L, torch, dev_in, cconf, original, steps):
    """sz3b_decompress of the stream the timed leg wrote (512^3 float32, GPU lossless stage): output left in HBM
    (`value`-like) and delivered to pinned host memory (`e2e`-like), per-stage times, the reference decoder's time on
    the host cores for the same stream, and parity (both decoders return the same bits)."""
    from common import Config
    n = original.size
    cap = L.sz3b_compress_bound(0, C.byref(cconf))
    cmp_buf = torch.empty(cap, dtype=torch.uint8).pin_memory()
    cmp_ptr = cmp_buf.data_ptr()
    size = C.c_size_t(0)
    if L.sz3b_compress(0, C.byref(cconf), C.c_void_p(dev_in.data_ptr()), 1, C.c_void_p(cmp_ptr), C.c_size_t(cap), C.byref(size), None) != 0:
        raise RuntimeError(L.sz3b_last_error().decode())
    csize = size.value
    dev_out = torch.empty(n, dtype=torch.float32, device="cuda")
    host_out = torch.empty(n, dtype=torch.float32).pin_memory()
    conf = Config()

    def one(ptr, loc):
        rc = L.sz3b_decompress(0, C.c_void_p(cmp_ptr), C.c_size_t(csize), C.c_void_p(ptr), loc, C.byref(conf))
        if rc != 0:
            raise RuntimeError(L.sz3b_last_error().decode())

    def timed(ptr, loc):
        for _ in range(2):
            one(ptr, loc)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            one(ptr, loc)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    ms_dev = timed(dev_out.data_ptr(), 1)
    names, ms, ln = (C.c_char_p * 64)(), (C.c_double * 64)(), (C.c_int * 64)()
    k = L.sz3b_last_profile(names, ms, ln, 64)
    stages = {}
    for i in range(k):
        stages[names[i].decode()] = round(stages.get(names[i].decode(), 0.0) + ms[i], 4)
    ms_host = timed(host_out.data_ptr(), 0)
    r = {"workload": "a stream of the bench workload (512^3 float32, abs 1e-3, default lossless policy)", "compressed_bytes": csize,
         "ms_device_resident": ms_dev, "ms_pinned_host": ms_host,
         "GBps_device_resident": n * 4 / ms_dev / 1e6, "GBps_pinned_host": n * 4 / ms_host / 1e6, "stages_ms": stages}
    lib, prefix, kind = cpu_checker()
    if lib is not None:
        dec = np.empty(n, dtype=np.float32)
        t0 = time.perf_counter()
        rc, _ = cpu_decompress(lib, prefix, cmp_ptr, csize, dec)
        t_ref = time.perf_counter() - t0
        same = bool(rc == 0 and np.array_equal(dec.view(np.uint32), host_out.numpy().view(np.uint32)))
        r["reference"] = {"ms": t_ref * 1e3, "GBps": n * 4 / t_ref / 1e9, "kind": kind, "threads": 1,
                          "note": "SZ_decompress of the same stream (the reference decodes a single stream on one thread)"}
        r["parity"] = {"same_bits_as_reference_decoder": same,
                       "max_abs_error": float(np.max(np.abs(dec.astype(np.float64) - original.reshape(-1).astype(np.float64))))}
    return r


def other_configs(L, torch):
    """BASELINE.json configs #3 (384^3 float64, regression predictor, REL 1e-4) and one slab of #5 (256x2048x2048 float32,
    abs 1e-3): device-resident and pinned-host time through sz3b_compress, ratio, the reference's time on the host cores,
    and parity (the reference decodes our stream within the bound; for #3 additionally the same reconstruction as the
    reference's own stream, i.e. the same indices and coefficients)."""
    from common import ALGO_INTERP_LORENZO, ALGO_LORENZO_REG, EB_REL, Config, field_g3, make_config
    lib, prefix, kind = cpu_checker()
    cores = host_cores()
    res = {}

    def run(name, data, conf, alg_bytes_name, reps):
        code = 0 if data.dtype == np.float32 else 1
        pinned = torch.from_numpy(data).pin_memory()
        dev = pinned.cuda()
        cap = L.sz3b_compress_bound(code, C.byref(conf))
        out = torch.empty(cap, dtype=torch.uint8).pin_memory().numpy()
        size = C.c_size_t(0)
        used = Config()

        def one(ptr, loc):
            rc = L.sz3b_compress(code, C.byref(conf), C.c_void_p(ptr), loc, out.ctypes.data_as(C.c_char_p), C.c_size_t(cap), C.byref(size),
                                 C.byref(used))
            if rc != 0:
                raise RuntimeError(L.sz3b_last_error().decode())

        def timed(ptr, loc):
            for _ in range(2):
                one(ptr, loc)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                one(ptr, loc)
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / reps

        ms_dev = timed(dev.data_ptr(), 1)
        names, ms, ln = (C.c_char_p * 64)(), (C.c_double * 64)(), (C.c_int * 64)()
        k = L.sz3b_last_profile(names, ms, ln, 64)
        stages = {}
        for i in range(k):
            stages[names[i].decode()] = stages.get(names[i].decode(), 0.0) + ms[i]
        ms_e2e = timed(pinned.data_ptr(), 0)
        r = {"shape": list(data.shape), "dtype": str(data.dtype), "ms_device_resident": ms_dev, "ms_pinned_host": ms_e2e,
             "GBps_device_resident": data.nbytes / ms_dev / 1e6, "GBps_pinned_host": data.nbytes / ms_e2e / 1e6,
             "ratio": data.nbytes / size.value, "stages_ms": {a: round(b, 4) for a, b in stages.items()}}
        pq = stages.get("predict_quantize")
        if pq:
            peak, _ = measured_peak()
            alg = data.size * (data.itemsize + 4)
            r["roofline"] = {"kernel": alg_bytes_name, "algorithmic_bytes": alg, "ms": pq, "achieved": alg / pq / 1e6,
                             "frac": alg / pq / 1e6 / peak, "unit": "GB/s"}
        # decompression of the stream just written, output left in HBM
        try:
            dev_out = torch.empty(data.size, dtype=torch.float32 if data.dtype == np.float32 else torch.float64, device="cuda")
            dconf = Config()

            def dec_one():
                if L.sz3b_decompress(code, out.ctypes.data_as(C.c_char_p), C.c_size_t(size.value), C.c_void_p(dev_out.data_ptr()), 1,
                                     C.byref(dconf)) != 0:
                    raise RuntimeError(L.sz3b_last_error().decode())

            dec_one()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                dec_one()
            e1.record()
            torch.cuda.synchronize()
            ms_dec = e0.elapsed_time(e1) / reps
            k = L.sz3b_last_profile(names, ms, ln, 64)
            dstages = {}
            for i in range(k):
                dstages[names[i].decode()] = round(dstages.get(names[i].decode(), 0.0) + ms[i], 4)
            r["decompress"] = {"ms_device_resident": ms_dec, "GBps_device_resident": data.nbytes / ms_dec / 1e6, "stages_ms": dstages,
                               "max_abs_error": float(np.max(np.abs(dev_out.cpu().numpy().astype(np.float64) - data.reshape(-1).astype(np.float64))))}
            del dev_out
        except Exception as ex:
            r["decompress"] = {"error": str(ex)[:200]}
        if lib is not None:
            threads = set_ref_threads(lib, prefix, cores)
            rconf = Config.from_buffer_copy(bytes(conf))
            rconf.openmp = 1
            times, rsize, rout = cpu_compress_time(lib, prefix, data, rconf, 2)
            r["reference"] = {"ms": min(times) * 1e3, "GBps": data.nbytes / min(times) / 1e9, "ratio": data.nbytes / rsize,
                              "threads": threads, "kind": kind}
            dec = np.empty_like(data)
            rc, dconf = cpu_decompress(lib, prefix, out.ctypes.data, size.value, dec)
            err = float(np.max(np.abs(dec.astype(np.float64) - data.astype(np.float64)))) if rc == 0 else None
            r["parity"] = {"reference_decodes_ours": rc == 0, "max_abs_error": err, "bound": float(dconf.absErrorBound),
                           "within_bound": bool(rc == 0 and err <= dconf.absErrorBound)}
        return r

    d3 = field_g3((384, 384, 384), np.float64)
    c3 = make_config(d3.shape, cmprAlgo=ALGO_LORENZO_REG, errorBoundMode=EB_REL, relErrorBound=1e-4, lorenzo=0, lorenzo2=0, regression=1)
    res["config3_384c_f64_regression_rel1e-4"] = run("c3", d3, c3, "regression fit + chain + k_reg_predict", 5)
    del d3
    z = np.arange(256, dtype=np.float32)[:, None, None]
    y = np.arange(2048, dtype=np.float32)[None, :, None]
    x = np.arange(2048, dtype=np.float32)[None, None, :]
    tp = np.float32(2 * np.pi)
    d5 = (np.sin(tp * x / np.float32(64)) * np.cos(tp * y / np.float32(96)) + np.float32(0.5) * np.sin(tp * z / np.float32(128) + np.float32(0.3))
          + np.float32(0.25) * np.sin(tp * (x + y + z) / np.float32(37)))
    d5 += np.float32(0.002) * np.random.default_rng(99).standard_normal(d5.shape, dtype=np.float32)
    c5 = make_config(d5.shape, cmprAlgo=ALGO_INTERP_LORENZO, absErrorBound=EB)
    res["config5_one_slab_256x2048x2048_f32_abs1e-3"] = run("c5", np.ascontiguousarray(d5), c5, "predict_quantize (box schedule)", 3)
    return res


def _merge(stages):
    acc, order = {}, []
    for n, m, l in stages:
        if n not in acc:
            acc[n] = [0.0, 0]
            order.append(n)
        acc[n][0] += m
        acc[n][1] += l
    return [(n, acc[n][0], acc[n][1]) for n in order]


if __name__ == "__main__":
    a = parse()
    # The contract is ONE JSON line on stdout.  Libraries underneath also write there at the C level (NCCL prints its
    # version line, the reference printf()s "OpenMP enabled for compression ..."): send file descriptor 1 to stderr for
    # the duration of the run and give Python's stdout the original descriptor back, so that only print() reaches it.
    sys.stdout.flush()
    _real = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(_real, "w", buffering=1)
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
    sys.stdout.flush()
