# GPU-box script: A/B of differently built copies of the library (build/variants/*.so, SZ3B_LIB_PATH): parity of the box
# schedule, then the stage times of predict+quantize at 512^3.  usage: gpurun -- 'bash tests/gpu_variants.sh'
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for lib in sz3_b200/lib/libsz3b200.so build/variants/*.so; do
  echo "== $lib"
  export SZ3B_LIB_PATH=$PWD/$lib
  timeout 300 python -m pytest tests/test_gpu_interp.py -m gpu -x -q -k "box" 2>&1 | tail -1
  timeout 120 python tools/prof_decompose.py 0 5 2>&1 | tail -3 | cut -c1-400
done 2>&1 | tee gpurun_out/variants.log
