"""The reference's HDF5 filter plugin (tools/H5Z-SZ3/src/H5Z_SZ3.cpp, UNMODIFIED) on this repo's drop-in headers.
HDF5 is not in the image, so the plugin is compiled against a stub of the HDF5 plugin API (tests/h5stub) and driven the
way libhdf5 drives a filter: plugin info, set_SZ3_conf_to_H5, the set_local callback, then the filter callback on a
chunk, forwards and with H5Z_FLAG_REVERSE (tests/h5stub/h5stub_driver.cpp; built by tools/build_h5z_test.py from
build()).  CPU part: plugin info and set_local (Config through cd_values).  GPU part: chunks of float / double / int32 /
int64 through the filter callback; an element type outside the GPU path is refused."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "build", "h5z", "h5z_roundtrip")
needs_exe = pytest.mark.skipif(not os.path.exists(EXE), reason="build/h5z/h5z_roundtrip not built (tools/build_h5z_test.py needs /root/reference)")


@needs_exe
def test_h5z_plugin_info_and_set_local():
    r = subprocess.run([EXE], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "set_local ok" in r.stdout


@needs_exe
@pytest.mark.gpu
def test_h5z_filter_roundtrip_on_gpu():
    r = subprocess.run([EXE, "gpu"], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "h5z roundtrip ok" in r.stdout and "uint16 chunk refused" in r.stdout
