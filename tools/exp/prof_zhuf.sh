cd $GRAFT_REPO_ROOT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_zhuf_build -s 2 -c 1 -o gpurun_out/prof_zhuf python tests/step_profile.py 512 3 2 > gpurun_out/ncu_zhuf.log 2>&1
timeout 300 python tests/step_profile.py 512 4 2 2>&1 | grep "device" | tail -2 | cut -c1-400
