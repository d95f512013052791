# GPU-box script: stream parity tests + launch list of the encode stage.  usage: gpurun -- 'bash tests/gpu_pack.sh TAG'
TAG=${1:-pack}
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_compress.py tests/test_golden.py tests/test_gpu_blockwise.py tests/test_gpu_decompress.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/prof_decompose.py 0 4 2>&1 | tail -2 | cut -c1-300
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_$TAG.csv \
    python tools/prof_decompose.py 0 2 > /dev/null 2>&1
