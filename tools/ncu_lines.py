#!/usr/bin/env python
"""Per-source-line view of one kernel of an ncu report: executed warp instructions and stall samples, joined with the
line table of the cubin (nvdisasm --print-line-info).

  python tools/ncu_lines.py gpurun_out/prof.ncu-rep build/obj/interp_box.o k_interp_boxILb1 [top]
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep, obj, kpat = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
lines = {}   # offset -> (file, line)
cur = None
infunc = False
for ln in dis.splitlines():
    if ln.startswith("//-") and ".text." in ln:
        infunc = kpat in ln
        continue
    if not infunc:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*)", ln)
    if m:
        lines[int(m.group(1), 16)] = cur
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None
agg = collections.defaultdict(lambda: [0, 0, 0, 0])   # warp inst, thread inst, samples, sass count
base = None
tot = [0, 0, 0]
opagg = collections.defaultdict(int)
for r in rows:
    if r and r[0] == "Address":
        hdr = {h: i for i, h in enumerate(r)}
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    addr = int(r[0], 16)
    if base is None:
        base = addr
    key = lines.get(addr - base, ("?", 0))
    wi, ti, sm = int(r[hdr["Instructions Executed"]]), int(r[hdr["Thread Instructions Executed"]]), int(r[hdr["# Samples"]])
    a = agg[key]
    a[0] += wi
    a[1] += ti
    a[2] += sm
    a[3] += 1
    tot[0] += wi
    tot[1] += ti
    tot[2] += sm
    opagg[r[hdr["Source"]].split()[0].split(".")[0] if not r[hdr["Source"]].strip().startswith("@") else r[hdr["Source"]].split()[1].split(".")[0]] += wi
print("total warp inst %d, thread inst %d (%.1f active lanes), samples %d" % (tot[0], tot[1], tot[1] / max(1, tot[0]), tot[2]))
print("| file:line | warp inst | share | lanes | samples | share | sass |\n|---|---:|---:|---:|---:|---:|---:|")
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("| %s:%d | %d | %.1f%% | %.1f | %d | %.1f%% | %d |" % (key[0], key[1], a[0], 100.0 * a[0] / tot[0], a[1] / max(1, a[0]), a[2],
                                                            100.0 * a[2] / max(1, tot[2]), a[3]))
print("\nby opcode (warp inst):")
for k, v in sorted(opagg.items(), key=lambda kv: -kv[1])[:25]:
    print("  %-10s %10d  %.1f%%" % (k, v, 100.0 * v / tot[0]))
