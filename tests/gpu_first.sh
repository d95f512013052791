cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,memory.total --format=csv
nproc; free -g | head -2
timeout 900 python -m pytest tests/test_gpu_interp.py tests/test_gpu_compress.py -m gpu -x -q 2>&1 | tail -30
timeout 600 python tests/quick_bench.py 512 2>&1 | tail -80
