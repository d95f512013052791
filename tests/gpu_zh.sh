# GPU-box script: lossless-stage + blockwise parity tests, per-step profile, ncu launch list of one policy-2 step, Lorenzo timings
TAG=${1:-zh}
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_zhuf.py tests/test_gpu_blockwise.py tests/test_gpu_decompress.py -m gpu -x -q 2>&1 | tail -6
timeout 300 python tests/step_profile.py 512 6 2 2>&1 | tail -12
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_$TAG.csv python tests/step_profile.py 512 2 2 > gpurun_out/ncu_$TAG.log 2>&1
python tools/ncu_summary.py launches gpurun_out/launches_$TAG.csv 2>/dev/null | head -60
SZ3B_VERBOSE=1 timeout 900 python tests/bench_configs.py lz lz512 > gpurun_out/lz_$TAG.json 2>gpurun_out/lz_$TAG.err; cut -c1-420 gpurun_out/lz_$TAG.json; grep sz3b gpurun_out/lz_$TAG.err | sort | uniq -c
