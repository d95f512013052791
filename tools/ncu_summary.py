#!/usr/bin/env python
"""Condense ncu outputs into the tracked summaries under profiles/.

  python tools/ncu_summary.py launches gpurun_out/launches_TAG.csv  > profiles/TAG_launches.md
  python tools/ncu_summary.py full     gpurun_out/prof_X_TAG.ncu-rep > profiles/TAG_full.md
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_registers", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum",
]


def launches(path):
    rows = list(csv.reader(open(path)))
    hdr = None
    agg = collections.OrderedDict()
    seq = []
    for r in rows:
        if len(r) > 5 and r[0] == "ID":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            try:
                v = float(d["Metric Value"].replace(",", ""))
            except ValueError:
                continue
            k = d["Kernel Name"].split("(")[0]
            a = agg.setdefault(k, [0, 0.0])
            a[0] += 1
            a[1] += v
            seq.append((d["ID"], k, d["Grid Size"], d["Block Size"], v))
    tot = sum(a[1] for a in agg.values())
    print("# ncu launch list (gpu__time_duration.sum, --clock-control none): %s\n" % path)
    print("| kernel | launches | total us | share |\n|---|---:|---:|---:|")
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("| `%s` | %d | %.1f | %.1f%% |" % (k, n, t / 1e3, 100 * t / tot))
    print("\nLast 40 launches (one full bench step):\n\n| id | kernel | grid | block | us |\n|---|---|---|---|---:|")
    for s in seq[-40:]:
        print("| %s | `%s` | %s | %s | %.1f |" % (s[0], s[1], s[2], s[3], s[4] / 1e3))


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print("# ncu --set full capture: %s\n" % path)
    for r in rows[2:]:
        print("## %s grid %s block %s\n" % (r[idx["Kernel Name"]].split("(")[0], r[idx["Grid Size"]], r[idx["Block Size"]]))
        print("| metric | value | unit |\n|---|---:|---|")
        for k in KEYS:
            if k in idx:
                print("| %s | %s | %s |" % (k, r[idx[k]], units[idx[k]]))
        stall = [h for h in hdr if "issue_stalled" in h and h.endswith("per_warp_active.pct")]
        vals = sorted([(float(r[idx[h]] or 0), h) for h in stall], reverse=True)[:5]
        for v, h in vals:
            print("| %s | %.1f | %% |" % (h.replace("smsp__average_warps_issue_stalled_", "stall_").replace("_per_warp_active.pct", ""), v))
        print()


def traffic(path):
    """profiles/traffic.json for bench.py: DRAM bytes of one predict+quantize step = sum over the level launches of one
    step in an `ncu --set full` capture (dram__bytes_read.sum + dram__bytes_write.sum)."""
    import json
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}

    def to_bytes(v, u):
        m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        return float(v) * m.get(u, 1)

    per = []
    for r in rows[2:]:
        rd = to_bytes(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]])
        wr = to_bytes(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
        per.append({"kernel": r[idx["Kernel Name"]].split("(")[0], "grid": r[idx["Grid Size"]], "dram_read": rd, "dram_write": wr,
                    "time_us": float(r[idx["gpu__time_duration.sum"]]) * {"us": 1, "ms": 1e3, "ns": 1e-3}.get(units[idx["gpu__time_duration.sum"]], 1)})
    print(json.dumps({"source": path, "predict_quantize_dram_bytes": sum(p["dram_read"] + p["dram_write"] for p in per), "launches": per}, indent=1))


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic}[sys.argv[1]](sys.argv[2])
