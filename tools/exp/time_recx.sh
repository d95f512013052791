cd $GRAFT_REPO_ROOT
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"k_box_recover_x" -c 2 --csv python tools/prof_decompress.py 2 2>/dev/null | grep -E "k_box_recover" | awk -F'","' '{print $(NF-2), $NF}' | tr -d '"'
timeout 300 python tools/prof_decompress.py 4 2>&1 | tail -1 | cut -c1-300
