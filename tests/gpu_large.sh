# GPU-box script: large-shape checks (one slab of BASELINE configs #4 and #5) and the other configurations' timings
TAG=${1:-lg}
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python tests/large_check.py --c4 2>&1 | tail -8 | cut -c1-700
timeout 900 python tests/large_check.py 2>&1 | tail -8 | cut -c1-700
timeout 900 python tests/bench_configs.py c3 c4s > gpurun_out/configs_$TAG.json 2> gpurun_out/configs_$TAG.err; cut -c1-900 gpurun_out/configs_$TAG.json
