# GPU-box script (diagnostics): where does the 2^31-element 4-D case fail?
cd $GRAFT_REPO_ROOT
free -g | head -2
echo "== 4-D 31x256x512x512 (just below 2^31)"; timeout 600 python tests/large_check.py --c4 31 256 512 512 2>&1 | tail -3 | cut -c1-200
echo "== 4-D 32x256x512x512 with the reference decoder"; timeout 1500 python tests/large_check.py --c4 --ref 32 256 512 512 2>&1 | tail -6 | cut -c1-200
