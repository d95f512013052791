# GPU-box script: blockwise / decompress parity tests first, then the whole suite, then Lorenzo timings and a bench line.
TAG=${1:-lz}
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader; nproc
timeout 600 python -m pytest tests/test_gpu_blockwise.py tests/test_gpu_decompress.py -m gpu -x -q 2>&1 | tail -15
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 600 python tests/bench_configs.py lz > gpurun_out/lz_$TAG.json 2>gpurun_out/lz_$TAG.err; cat gpurun_out/lz_$TAG.json; tail -3 gpurun_out/lz_$TAG.err
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -3 gpurun_out/bench_$TAG.err; cat gpurun_out/bench_$TAG.json
