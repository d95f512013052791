cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_compress.py tests/test_gpu_decompress.py -m gpu -x -q 2>&1 | tail -2
timeout 300 python tools/prof_decompose.py 0 3 2>&1 | tail -1 | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_pack_count|k_pack_write|k_pack_scan|k_hist_u16' -s 4 -c 4 -o gpurun_out/prof_pack2 python tools/prof_decompose.py 0 2 > gpurun_out/ncu_pack2.log 2>&1
