# GPU-box script: compute-sanitizer over small cases of the kernels added last (memcheck, then racecheck on tiny inputs)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
timeout 900 $S --tool memcheck --error-exitcode 9 --print-limit 5 python -m pytest tests/test_zhuf.py -m gpu -x -q -k "gpu_frames and (5000 or tail or two or 100 or symbols)" 2>&1 | tail -6
timeout 900 $S --tool memcheck --error-exitcode 9 --print-limit 5 python -m pytest tests/test_gpu_blockwise.py -m gpu -x -q -k "lorenzo_stack or special" 2>&1 | tail -6
timeout 900 $S --tool memcheck --error-exitcode 9 --print-limit 5 python -m pytest tests/test_gpu_decompress.py -m gpu -x -q -k "shape12 or shape13 or shape14 or shape17" 2>&1 | tail -6
timeout 900 $S --tool racecheck --error-exitcode 9 --print-limit 5 python -m pytest tests/test_zhuf.py -m gpu -x -q -k "gpu_frames and (5000 or tail)" 2>&1 | tail -6
timeout 900 $S --tool racecheck --error-exitcode 9 --print-limit 5 python -m pytest tests/test_gpu_blockwise.py -m gpu -x -q -k "lorenzo_stack and (shape2 or shape6)" 2>&1 | tail -6
