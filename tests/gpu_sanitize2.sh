# GPU-box script (round 2): compute-sanitizer over the kernels added or rewritten in round 2 -- box schedule (TMA
# planes, CTA split of coarse tiles), k_hist_u16, the packer, the Huffman write pass / tile scan, integer element types.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
timeout 900 $S --tool memcheck --error-exitcode 9 --print-limit 5 python -m pytest tests/test_gpu_interp.py -m gpu -x -q -k "box and not 512 and not rejects" 2>&1 | tail -5
timeout 900 $S --tool memcheck --error-exitcode 9 --print-limit 5 python -m pytest tests/test_gpu_compress.py -m gpu -x -q -k "stream_identical_small" 2>&1 | tail -5
timeout 900 $S --tool memcheck --error-exitcode 9 --print-limit 5 python -m pytest tests/test_gpu_decompress.py -m gpu -x -q -k "shape0 or shape3 or shape12 or shape17" 2>&1 | tail -5
timeout 900 $S --tool memcheck --error-exitcode 9 --print-limit 5 python -m pytest tests/test_gpu_int_types.py -m gpu -x -q -k "interp_variants or regression_stacks" 2>&1 | tail -5
timeout 900 $S --tool racecheck --error-exitcode 9 --print-limit 5 python -m pytest tests/test_gpu_interp.py -m gpu -x -q -k "interp3d_box_variants" 2>&1 | tail -5
timeout 900 $S --tool racecheck --error-exitcode 9 --print-limit 5 python -m pytest tests/test_gpu_compress.py -m gpu -x -q -k "stream_identical_small and shape0" 2>&1 | tail -5
