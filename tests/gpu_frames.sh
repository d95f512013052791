# GPU-box script: the GPU frame decoder (zhuf_dec.cuh) -- parity tests, then decompression of the 512^3 bench stream
# with the decoder off and on.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
{
timeout 300 python -m pytest tests/test_gpu_decompress.py -m gpu -x -q -k "frames or bit_identical" 2>&1 | tail -3
for m in ${MODES:-0 1}; do echo "== SZ3B_FRAME_DECODER=$m"; SZ3B_FRAME_DECODER=$m timeout 120 python tools/prof_decompress.py 6 2>&1 | tail -2 | cut -c1-460; done
} 2>&1 | tee gpurun_out/frames.log
