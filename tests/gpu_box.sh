# GPU-box script: interpolation parity tests, launch list and a full ncu capture of the level-1 box launch.
# usage: gpurun --timeout 900 -- 'bash tests/gpu_box.sh TAG'
TAG=${1:-box}
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_interp.py tests/test_gpu_compress.py -m gpu -x -q 2>&1 | tail -4
timeout 300 python tools/prof_decompose.py 0 4 2>&1 | tail -2 | cut -c1-300
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_$TAG.csv \
    python tools/prof_decompose.py 0 2 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_interp_box -s 4 -c 1 -o gpurun_out/prof_$TAG \
    python tools/prof_decompose.py 0 1 > gpurun_out/ncu_$TAG.log 2>&1
