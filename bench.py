#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric: compress GB/s + ratio, 3-D float32 512^3, abs error bound 1e-3, on N B200s.

    python bench.py --gpus N --steps K --warmup W            # the CUDA library (libsz3b200.so) through its C ABI
    python bench.py --impl reference --gpus N ...            # the reference's own OpenMP CPU path (oracle/_ref)

One "step" = one pass of the whole hot path (tuner -> predict+quantize -> histogram/Huffman -> bit pack -> zstd) over
one 512x512x512 float32 array per GPU.  For N > 1 rank r owns slab r of an (N*512)x512x512 array (outermost-dimension
slabs, the reference's OpenMP decomposition, api/impl/SZImplOMP.hpp:43-86); the only exchange is an all-gather of the
per-slab byte counts (weak scaling: work per GPU is fixed).

  value     whole-job GB/s with the input already resident in HBM (compressed stream delivered to host memory)
  e2e       same, input in pinned HOST memory: H2D of the array and D2H of the packed stream inside the timed region
  roofline  the fused predict+quantize launches (k_interp_*): algorithmic bytes N*(sizeof(T)+4) / their device time
            (CUDA events recorded by the library on its own stream), against MEASURED_PEAKS.json's HBM copy peak
  cpu_baseline  oracle/_ref (the unmodified reference, conf.openmp = true) on the box's host cores, same array
  extras    host_zstd_policy (the same two measurements with the reference's own host zstd call), e2e_two_callers
            (N = 1: the e2e call issued by two host threads at once), config.host (cores, threads per rank);
            --diag prints per-rank step times, the host-link rate and an A/B of the host thread settings on stderr
"""
import argparse
import ctypes as C
import json
import os
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


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--edge", type=int, default=EDGE, help="cube edge (default 512 = the metric's configuration)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--diag", action="store_true", help="per-rank step times and host-link rate on stderr")
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
    """dram bytes per predict+quantize step from the committed ncu --set full capture (profiles/traffic.json)."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get("predict_quantize_dram_bytes")
    except Exception:
        return None


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


class _StdoutToStderr:
    """The reference printf()s a line per OpenMP call; keep stdout for the one JSON line."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *a):
        os.dup2(self.saved, 1)
        os.close(self.saved)


def cpu_compress_time(lib, prefix, data, edge, reps, openmp=True):
    with _StdoutToStderr():
        return _cpu_compress_time(lib, prefix, data, edge, reps, openmp)


def _cpu_compress_time(lib, prefix, data, edge, reps, openmp=True):
    conf = make_conf(edge, openmp=1 if openmp else 0)
    cap = getattr(lib, prefix + "_size_bound")(0, C.byref(conf))
    out = np.empty(cap, dtype=np.uint8)
    best, size = None, 0
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        size = getattr(lib, prefix + "_compress")(0, C.byref(conf), data.ctypes.data_as(C.c_void_p),
                                                  out.ctypes.data_as(C.c_char_p), C.c_size_t(cap))
        dt = time.perf_counter() - t0
        assert size > 0, "reference compression failed"
        times.append(dt)
        best = dt if best is None else min(best, dt)
    return best, times, size


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    lib, prefix, kind = cpu_checker()
    if lib is None:
        print(json.dumps({"impl": "reference", "unavailable": "neither oracle/_ref/libsz3ref.so nor oracle/libsz3oracle.so is built"}))
        return
    cores = os.cpu_count() or 1
    edge = args.edge
    data = slab_field(0, edge)
    _, _, _ = cpu_compress_time(lib, prefix, data, edge, max(args.warmup, 1))
    _, times, size = cpu_compress_time(lib, prefix, data, edge, args.steps)
    total = sum(times)
    gbs = data.nbytes * args.steps / total / 1e9
    sample = (f"one {edge}^3 float32 array per step (the N=1 workload), SZ_compress with conf.openmp=true, "
              f"OMP threads = {cores}; CPU throughput does not grow with --gpus")
    line = {
        "impl": "reference", "metric": "compress throughput, 3D f32 512^3 abs-eb 1e-3", "value": gbs, "unit": "GB/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "ratio": data.nbytes / size,
        "config": {"workload": f"3D float32 {edge}x{edge}x{edge} ALGO_INTERP_LORENZO abs-eb 1e-3", "field": "G3 (SURVEY.md 8d), seeded"},
        "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------------------------
# GPU side
# ----------------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from common import Config, product_lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device visible; the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = product_lib()
    if L is None:
        raise SystemExit("bench.py: sz3_b200/lib/libsz3b200.so missing; run `python -c 'import __graft_entry__ as g; g.build()'`")
    L.sz3b_last_error.restype = C.c_char_p
    if args.lossless_policy is not None:
        L.sz3b_set_lossless_policy(args.lossless_policy)
    # one rank per GPU on one host: every rank gets its share of the cores (16 pool threads per rank on 2 cores per
    # rank cost half of e2e; tests/gpu_cores.sh).  Waiting threads keep spinning: yielding measured slower.
    cores = len(os.sched_getaffinity(0))
    host_threads = args.host_threads if args.host_threads is not None else (max(2, cores // world) if world > 1 else 0)
    host_wait = args.host_wait if args.host_wait is not None else 0
    L.sz3b_set_host_threads(host_threads)
    L.sz3b_set_host_wait(host_wait)
    edge = args.edge
    nbytes = edge ** 3 * 4
    host = slab_field(rank, edge)
    pinned = torch.from_numpy(host).pin_memory()
    dev = pinned.cuda(non_blocking=False)

    # single GPU: SZ_compress.  N GPUs: rank r = slab r of the OpenMP container (SZImplOMP.hpp), sizes all-gathered.
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
    sizes_dev = torch.zeros(world, dtype=torch.int64, device="cuda") if world > 1 else None

    def step(ptr, loc):
        if world == 1:
            rc = L.sz3b_compress(0, C.byref(gconf), C.c_void_p(ptr), loc, out_np.ctypes.data_as(C.c_char_p), C.c_size_t(cap),
                                 C.byref(size), C.byref(used))
        else:
            rc = L.sz3b_compress_slab(0, C.byref(gconf), rank, world, C.c_void_p(ptr), loc, C.c_double(0.0),
                                      out_np.ctypes.data_as(C.c_char_p), C.c_size_t(cap), C.byref(size), blob, C.byref(blob_len))
        if rc != 0:
            raise RuntimeError(L.sz3b_last_error().decode())
        if world > 1:   # the one exchange of the path: per-slab byte counts -> offsets (SZImplOMP.hpp:93-105)
            mine = torch.tensor([size.value], dtype=torch.int64, device="cuda")
            dist.all_gather_into_tensor(sizes_dev, mine)
        return size.value

    def profile():
        names = (C.c_char_p * 64)()
        ms = (C.c_double * 64)()
        launches = (C.c_int * 64)()
        n = L.sz3b_last_profile(names, ms, launches, 64)
        return [(names[i].decode(), ms[i], launches[i]) for i in range(min(n, 64))]

    def timed(ptr, loc, steps, collect):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        pq_ms, launches, csize = 0.0, 0, 0
        for _ in range(steps):
            csize = step(ptr, loc)
            if collect:
                for name, ms, nl in profile():
                    launches += nl
                    if name == "predict_quantize":
                        pq_ms += ms
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

    local_ms = []
    for _ in range(max(args.warmup, 3)):
        step(dev.data_ptr(), 1)
    # untimed settling beyond W: the first calls still grow per-workspace buffers and the SM clock is still ramping
    # (the first three 512^3 steps run 30-50 % slower than the steady state); wait until two steps in a row agree
    prev = None
    for _ in range(40):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        step(dev.data_ptr(), 1)
        dt = time.perf_counter() - t0
        settled = prev is not None and abs(dt - prev) < 0.03 * prev
        if world > 1:   # every rank must leave the loop in the same iteration (step() holds a collective)
            flag = torch.tensor([1 if settled else 0], dtype=torch.int32, device="cuda")
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            settled = bool(flag.item())
        if settled:
            break
        prev = dt
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_dev, pq_ms, launches, csize = timed(dev.data_ptr(), 1, args.steps, True)
    for _ in range(max(args.warmup, 3)):
        step(pinned.data_ptr(), 0)
    ms_e2e, _, _, _ = timed(pinned.data_ptr(), 0, args.steps, False)
    h2d, d2h = C.c_size_t(0), C.c_size_t(0)
    L.sz3b_last_transfer(C.byref(h2d), C.byref(d2h))
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
        for th, wt in [(cores, 0), (max(2, cores // world), 0), (2, 0), (cores, 1), (max(2, cores // world), 1)]:
            L.sz3b_set_host_threads(th)
            L.sz3b_set_host_wait(wt)
            for _ in range(3):
                step(dev.data_ptr(), 1)
            a, _, _, _ = timed(dev.data_ptr(), 1, args.steps, False)
            for _ in range(2):
                step(pinned.data_ptr(), 0)
            b, _, _, _ = timed(pinned.data_ptr(), 0, args.steps, False)
            if rank == 0:
                print(f"[diag] host threads {th}, wait {['spin', 'yield'][wt]}: value path {a / args.steps:.2f} ms/step, "
                      f"e2e path {b / args.steps:.2f} ms/step (max over ranks)", file=sys.stderr, flush=True)
        L.sz3b_set_host_threads(host_threads)
        L.sz3b_set_host_wait(host_wait)
    clocks = sampler.stop() if rank == 0 else None
    # the same two measurements with the lossless stage on the host (zstd level 3 on every chunk, the reference's own
    # call), reported next to the headline so that the effect of the GPU lossless stage is visible
    policy = L.sz3b_get_lossless_policy()
    host_zstd = None
    if policy == 2:
        L.sz3b_set_lossless_policy(0)
        for _ in range(4):
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
    if world == 1 and args.steps >= 2:
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
        line = {
            "metric": "compress throughput, 3D f32 512^3 abs-eb 1e-3", "value": value, "unit": "GB/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "ratio": total_bytes / total_csize,
            "config": {"workload": f"3D float32 {edge}x{edge}x{edge} per GPU, ALGO_INTERP_LORENZO abs-eb 1e-3"
                                   + (f", slab-sharded over {world} GPUs (OpenMP container)" if world > 1 else ""),
                       "field": "G3 (SURVEY.md 8d), seeded", "l2": "input 512 MiB per step > 126 MB L2 (no explicit flush)",
                       "lossless_policy": {0: "host zstd-3 on every chunk", 1: "adaptive host zstd: probes, raw zstd frames where zstd gains < 1 % (include/sz3b.h)",
                                           2: "GPU lossless stage: zstd frames of Huffman-only literal blocks, one table per 128 KiB (sz3_b200/csrc/zhuf.cuh); decodes with the unmodified reference"}[policy],
                       "host": {"cores": cores, "threads_per_rank": host_threads or cores, "device_wait": ["spin", "yield"][host_wait]},
                       "value_path": "sz3b_compress, device-resident input, stream delivered to host",
                       "e2e_path": "sz3b_compress, pinned host input (H2D + D2H inside the timed region)"},
            "e2e": {"value": e2e, "unit": "GB/s", "h2d_bytes_per_step": h2d.value, "d2h_bytes_per_step": d2h.value,
                    "ms_per_step": ms_e2e / args.steps},
            "e2e_two_callers": two_callers,
            "host_zstd_policy": None if host_zstd is None or world > 1 else {
                "value": total_bytes * args.steps / (host_zstd[0] * 1e-3) / 1e9, "e2e": total_bytes * args.steps / (host_zstd[1] * 1e-3) / 1e9,
                "unit": "GB/s", "ratio": nbytes / host_zstd[2], "note": "same run with sz3b_set_lossless_policy(0): zstd level 3 on the host"},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": "predict_quantize (k_interp_anchor + k_interp_ltile x 5 levels)",
                         "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": ncu_traffic(), "algorithmic_bytes": alg_bytes,
                         "ms_per_step": pq_avg_ms},
            "clocks": clocks,
            "stages_ms": {n: round(m, 4) for n, m, _ in _merge(profile())},
        }
        if world == 1 and not args.no_cpu_baseline:
            lib, prefix, kind = cpu_checker()
            if lib is not None:
                cores = os.cpu_count() or 1
                best, times, rsize = cpu_compress_time(lib, prefix, host, edge, 3)
                line["cpu_baseline"] = {"value": nbytes / best / 1e9, "unit": "GB/s", "cores": cores, "kind": kind,
                                        "sample": f"the full {edge}^3 array, SZ_compress conf.openmp=true ({cores} OMP threads), best of 3",
                                        "ratio": nbytes / rsize}
            else:
                line["cpu_baseline"] = {"value": None, "unit": "GB/s", "cores": 0, "kind": "unavailable", "sample": "no oracle library built"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


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
