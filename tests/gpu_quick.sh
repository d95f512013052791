# GPU-box script: parity tests + per-step profile + one bench line (no profiler).  usage: gpurun --timeout 900 -- 'bash tests/gpu_quick.sh TAG'
TAG=${1:-q}
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 300 python tests/step_profile.py 512 6 2 2>&1 | tail -7 | cut -c1-250
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -3 gpurun_out/bench_$TAG.err; cat gpurun_out/bench_$TAG.json
