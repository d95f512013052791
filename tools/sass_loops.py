#!/usr/bin/env python
"""List the loops of one kernel's SASS with an instruction mix (static check before spending GPU time).
usage: python tools/sass_loops.py build/obj/interp_kernels.o <mangled kernel name> [dump_lo dump_hi]"""
import re, subprocess, sys
obj, fun = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", "-fun", fun, obj], capture_output=True, text=True).stdout
ins = []
for l in out.splitlines():
    m = re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);', l)
    if m: ins.append((int(m.group(1), 16), m.group(2).strip()))
a2i = {a: i for i, (a, _) in enumerate(ins)}
print("instructions:", len(ins))
if len(sys.argv) > 4:
    for a, t in ins[int(sys.argv[3]):int(sys.argv[4]) + 1]: print(t)
    sys.exit(0)
for i, (a, t) in enumerate(ins):
    if 'BRA' in t:
        m = re.search(r'0x([0-9a-f]+)', t)
        if m and int(m.group(1), 16) < a and int(m.group(1), 16) in a2i:
            j = a2i[int(m.group(1), 16)]
            body = [x[1] for x in ins[j:i + 1]]
            c = lambda k: sum(1 for b in body if k in b)
            print("loop %5d-%5d len %4d  DSETP %d LDG %d LDS %d STS %d STG %d F2F %d CALL %d local %d" %
                  (j, i, len(body), c('DSETP'), c('LDG'), c('LDS'), c('STS'), c('STG'), c('F2F'), c('CALL'), c('STL') + c('LDL')))
