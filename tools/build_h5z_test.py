#!/usr/bin/env python
"""Builds build/h5z/h5z_roundtrip: the reference's HDF5 filter plugin (tools/H5Z-SZ3/src/H5Z_SZ3.cpp + its header,
UNMODIFIED, read where they lie under /root/reference) compiled against this repo's drop-in SZ3 headers and the stub of
the HDF5 plugin API under tests/h5stub (HDF5 itself is not in this image), linked with tests/h5stub/h5stub_driver.cpp
and libsz3b200.  tests/test_h5z_filter.py runs it.  Nothing of the reference is copied into the repository."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build(ref="/root/reference"):
    src = os.path.join(ref, "tools", "H5Z-SZ3", "src", "H5Z_SZ3.cpp")
    inc = os.path.join(ref, "tools", "H5Z-SZ3", "include")
    if not os.path.exists(src):
        print("[h5z] %s absent: keeping whatever build/h5z holds" % src)
        return False
    out = os.path.join(ROOT, "build", "h5z")
    os.makedirs(out, exist_ok=True)
    exe = os.path.join(out, "h5z_roundtrip")
    lib = os.path.join(ROOT, "sz3_b200", "lib")
    stub = os.path.join(ROOT, "tests", "h5stub")
    deps = [src, os.path.join(inc, "H5Z_SZ3.hpp"), os.path.join(stub, "hdf5.h"), os.path.join(stub, "h5stub_driver.cpp"),
            os.path.join(lib, "libsz3b200.so"), os.path.join(ROOT, "include", "SZ3", "api", "sz.hpp"),
            os.path.join(ROOT, "include", "SZ3", "utils", "Config.hpp")]
    if os.path.exists(exe) and all(os.path.getmtime(exe) >= os.path.getmtime(d) for d in deps):
        return True
    subprocess.run(["g++", "-std=c++17", "-O1", "-w", "-I" + stub, "-I" + os.path.join(ROOT, "include"), "-I" + inc, src,
                    os.path.join(stub, "h5stub_driver.cpp"), "-o", exe, "-L" + lib, "-lsz3b200", "-Wl,-rpath,$ORIGIN/../../sz3_b200/lib"],
                   check=True)
    print("[h5z] built", exe)
    return True


if __name__ == "__main__":
    build(*sys.argv[1:2])
