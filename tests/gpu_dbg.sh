# GPU-box script (diagnostics): decoder rounds on the 2^31-element 4-D case, then the decompression tests
cd $GRAFT_REPO_ROOT
echo "== 4-D 32x256x512x512"; SZ3B_VERBOSE=1 timeout 900 python tests/large_check.py --c4 32 256 512 512 2>&1 | grep -v "lorenzo stack" | tail -7 | cut -c1-260
timeout 600 python -m pytest tests/test_gpu_decompress.py tests/test_gpu_interp.py -m gpu -x -q 2>&1 | tail -4
