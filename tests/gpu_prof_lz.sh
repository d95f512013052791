# GPU-box script: decompression parity (decoder rounds) + full ncu capture of one large front of the Lorenzo wavefront
TAG=${1:-lzp}
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_decompress.py tests/test_gpu_interp.py -m gpu -x -q 2>&1 | tail -4
echo "== 4-D 24x256x512x512"; SZ3B_VERBOSE=1 timeout 900 python tests/large_check.py --c4 24 256 512 512 2>&1 | grep -v "lorenzo stack" | tail -5 | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_bw_front -s 190 -c 1 -o gpurun_out/prof_bwfront_$TAG python tests/lz_one.py 256 0 > gpurun_out/ncu_bwfront_$TAG.log 2>&1
tail -2 gpurun_out/ncu_bwfront_$TAG.log
