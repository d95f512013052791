"""Regenerates the committed golden fixtures of tests/golden/ from the UNMODIFIED reference (oracle/_ref/libsz3ref.so,
built from /root/reference by `make -C oracle ref`).  Run in the build container:  python tests/golden/make_golden.py

Each fixture is one .npz: the input array (stored, not regenerated: numpy's sin / normal sampling are not guaranteed
to be bit-reproducible across CPUs) and the stream SZ_compress of the reference wrote for it; cases.json holds the
configuration of every case and the SHA-256 of the array the reference's SZ_decompress returns for that stream.
tests/test_golden.py checks the oracle restatement (CPU suite) and the CUDA path (GPU suite) against them, so the
pins hold on a box that has neither /root/reference nor the prebuilt reference library."""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from common import (ALGO_INTERP, ALGO_INTERP_LORENZO, ALGO_LORENZO_REG, EB_REL, field_g1, field_nd, make_config,  # noqa: E402
                    ref_lib)
from test_gpu_compress import ref_compress, ref_decompress  # noqa: E402

REG_ONLY = dict(lorenzo=0, lorenzo2=0, regression=1)
# name, input, config keywords, what the GPU suite compares ("stream": byte-identical compression + decode; "decode":
# bit-identical decode of the reference's stream).  Every case mirrors a live reference comparison of the GPU suite.
CASES = [
    ("interp_3d_f32", field_nd((8, 8, 128), np.float32), dict(cmprAlgo=ALGO_INTERP, absErrorBound=1.0), "stream"),
    ("tuned_3d_f32", field_nd((8, 8, 128), np.float32), dict(cmprAlgo=ALGO_INTERP_LORENZO, absErrorBound=1.0), "stream"),
    ("tuned_1d_f32", field_g1(20000), dict(cmprAlgo=ALGO_INTERP_LORENZO, absErrorBound=1e-4), "stream"),
    ("regression_2d_f32", field_nd((96, 96), np.float32), dict(cmprAlgo=ALGO_LORENZO_REG, absErrorBound=1e-3, **REG_ONLY), "stream"),
    ("interp_3d_anchor0_f32", field_nd((8, 8, 128), np.float32), dict(cmprAlgo=ALGO_INTERP, absErrorBound=1e-3, interpAnchorStride=0), "decode"),
    ("interp_1d_f32", field_nd((5000,), np.float32), dict(cmprAlgo=ALGO_INTERP, absErrorBound=1e-3), "decode"),
    ("regression_small_2d_f32", field_nd((40, 45), np.float32), dict(cmprAlgo=ALGO_LORENZO_REG, absErrorBound=1e-3, **REG_ONLY), "decode"),
    ("lorenzo_3d_f32", field_nd((24, 30, 36), np.float32), dict(cmprAlgo=ALGO_LORENZO_REG, regression=0, absErrorBound=1e-3), "decode"),
    ("lorenzo_reg_1d_f32", field_nd((3000,), np.float32), dict(cmprAlgo=ALGO_LORENZO_REG, blockSize=128, absErrorBound=1e-3), "decode"),
    ("composed_2d_f64", field_nd((130, 77), np.float64), dict(cmprAlgo=ALGO_LORENZO_REG, lorenzo2=1, blockSize=16, absErrorBound=1e-3), "decode"),
    ("lorenzo_reg_4d_f32", field_nd((9, 12, 20, 18), np.float32), dict(cmprAlgo=ALGO_LORENZO_REG, absErrorBound=1e-3), "decode"),
    ("regression_rel_3d_f64", field_nd((20, 24, 28), np.float64), dict(cmprAlgo=ALGO_LORENZO_REG, errorBoundMode=EB_REL, relErrorBound=1e-4, **REG_ONLY), "cpu"),
    ("lossless_2d_f32", field_nd((64, 64), np.float32), dict(cmprAlgo=ALGO_INTERP, absErrorBound=0.0), "cpu"),
    # OpenMP container: the slab count is the generating box's thread count and is read back from the stream
    ("omp_container_3d_f32", field_nd((24, 30, 36), np.float32), dict(cmprAlgo=ALGO_INTERP_LORENZO, absErrorBound=1e-3, openmp=1), "cpu"),
]


def main():
    assert ref_lib() is not None, "build the reference first: make -C oracle ref"
    index = []
    for name, data, kw, gpu in CASES:
        conf = make_config(data.shape, **kw)
        stream = ref_compress(data, conf)
        dec, dconf = ref_decompress(stream, data)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), data=data, stream=stream)
        index.append({"name": name, "shape": list(data.shape), "dtype": data.dtype.name, "config": kw, "gpu": gpu,
                      "stream_bytes": int(stream.size), "stream_sha256": hashlib.sha256(stream.tobytes()).hexdigest(),
                      "decoded_sha256": hashlib.sha256(dec.tobytes()).hexdigest(), "abs_error_bound": float(dconf.absErrorBound),
                      "algo_in_stream": int(dconf.cmprAlgo)})
        print(f"{name:28s} {str(data.shape):18s} {data.dtype.name:8s} -> {stream.size:7d} bytes, algo {dconf.cmprAlgo}")
    with open(os.path.join(HERE, "cases.json"), "w") as f:
        json.dump(index, f, indent=1)
        f.write("\n")


if __name__ == "__main__":
    main()
