# GPU-box script: blockwise / decompress parity tests, then Lorenzo-stack timings.  usage: gpurun -- 'bash tests/gpu_lz.sh TAG'
TAG=${1:-lz}
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_blockwise.py tests/test_gpu_decompress.py -m gpu -x -q 2>&1 | tail -5
SZ3B_VERBOSE=1 timeout 900 python tests/bench_configs.py lz lz512 > gpurun_out/lz_$TAG.json 2>gpurun_out/lz_$TAG.err; cut -c1-330 gpurun_out/lz_$TAG.json; grep sz3b gpurun_out/lz_$TAG.err | sort | uniq -c
