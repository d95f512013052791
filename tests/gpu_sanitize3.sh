# GPU-box script (round 2, late): compute-sanitizer over the decompression kernels added after r2d -- k_box_recover_x
# (TMA planes from the output array), the two-level Huffman decode tables, the staged regression coefficient
# recurrence, the streaming frame mirror -- and the OpenMP container paths.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
{
timeout 900 $S --tool memcheck --error-exitcode 9 --print-limit 5 python -m pytest tests/test_gpu_decompress.py -m gpu -x -q -k "bit_identical and gpu" 2>&1 | tail -4
timeout 900 $S --tool memcheck --error-exitcode 9 --print-limit 5 python -m pytest tests/test_gpu_decompress.py -m gpu -x -q -k "omp_container or special_values or device_pointer or lossless_and_noise" 2>&1 | tail -4
timeout 900 $S --tool racecheck --error-exitcode 9 --print-limit 5 python -m pytest tests/test_gpu_decompress.py -m gpu -x -q -k "bit_identical and gpu and (shape20 or shape21 or shape22 or shape23 or shape9 or shape12)" 2>&1 | tail -4
timeout 900 $S --tool synccheck --error-exitcode 9 --print-limit 5 python -m pytest tests/test_gpu_decompress.py -m gpu -x -q -k "bit_identical and gpu and (shape20 or shape22 or shape24)" 2>&1 | tail -4
} | tee gpurun_out/sanitize3.log
