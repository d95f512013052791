# GPU-box script (diagnostics): 1-D tests, full ncu capture of k_zhuf_build and k_bw_front
TAG=${1:-dbg}
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_compress.py -m gpu -x -q -k "one_dimensional" 2>&1 | tail -15
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_zhuf_build -s 2 -c 1 -o gpurun_out/prof_zbuild_$TAG python tests/step_profile.py 512 3 2 > gpurun_out/ncu_zbuild_$TAG.log 2>&1
tail -2 gpurun_out/ncu_zbuild_$TAG.log
ls -la gpurun_out/*.ncu-rep
