#!/usr/bin/env python
"""Builds the reference's Python binding `pysz` (tools/pysz/src/pysz/sz.pyx, sz.pxd -- UNMODIFIED, read where they lie
under /root/reference) against this repo's drop-in headers (include/SZ3/api/sz.hpp, include/SZ3/utils/Config.hpp) and
libsz3b200.so: the binding a pysz user has, with the GPU path behind it.

    python tools/build_pysz.py [reference root]      ->  build/pysz/pysz/{__init__.py, sz.<abi>.so}

Nothing of the reference is copied into the repository: the Cython translation (build/pysz/sz.cpp) and the extension
are build artefacts under the git-ignored build/ tree (they travel to the GPU box with the snapshot, like every other
built library).  Usage afterwards:  sys.path.insert(0, "build/pysz");  from pysz import sz, szConfig
"""
import os
import subprocess
import sys
import sysconfig

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build(ref="/root/reference"):
    src = os.path.join(ref, "tools", "pysz", "src")
    pyx = os.path.join(src, "pysz", "sz.pyx")
    if not os.path.exists(pyx):
        print("[pysz] %s absent: keeping whatever build/pysz holds" % pyx)
        return False
    import numpy
    out = os.path.join(ROOT, "build", "pysz")
    pkg = os.path.join(out, "pysz")
    os.makedirs(pkg, exist_ok=True)
    cpp = os.path.join(out, "sz.cpp")
    so = os.path.join(pkg, "sz" + sysconfig.get_config_var("EXT_SUFFIX"))
    lib = os.path.join(ROOT, "sz3_b200", "lib", "libsz3b200.so")
    deps = [pyx, os.path.join(src, "pysz", "sz.pxd"), lib, os.path.join(ROOT, "include", "SZ3", "api", "sz.hpp"),
            os.path.join(ROOT, "include", "SZ3", "utils", "Config.hpp"), os.path.join(ROOT, "include", "sz3b.h")]
    with open(os.path.join(pkg, "__init__.py"), "w") as f:   # the package's four public names (pysz/__init__.py)
        f.write('from pysz.sz import sz, szConfig, szErrorBoundMode, szAlgorithm\n\n'
                '__all__ = ["sz", "szConfig", "szErrorBoundMode", "szAlgorithm"]\n')
    if os.path.exists(so) and all(os.path.getmtime(so) >= os.path.getmtime(d) for d in deps):
        return True
    subprocess.run([sys.executable, "-m", "cython", "--cplus", "-3", "-I", src, pyx, "-o", cpp], check=True)
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-w", "-I" + os.path.join(ROOT, "include"),
                    "-I" + numpy.get_include(), "-I" + sysconfig.get_paths()["include"], cpp, "-o", so,
                    "-L" + os.path.dirname(lib), "-lsz3b200", "-Wl,-rpath,$ORIGIN/../../../sz3_b200/lib"], check=True)
    print("[pysz] built", so)
    return True


if __name__ == "__main__":
    build(*sys.argv[1:2])
