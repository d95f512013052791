# GPU-box script (diagnostics): where does the 2^31-element decompression fail?
cd $GRAFT_REPO_ROOT
echo "== A: 4-D 2^30 elements, policy 2"; timeout 600 python tests/large_check.py --c4 16 256 512 512 2>&1 | tail -4 | cut -c1-300
echo "== B: 4-D 2^31 elements, policy 0"; SZ3B_POLICY=0 timeout 900 python tests/large_check.py --c4 32 256 512 512 2>&1 | tail -4 | cut -c1-300
echo "== C: 3-D 2^31 elements, policy 2"; timeout 900 python tests/large_check.py 512 2048 2048 2>&1 | tail -4 | cut -c1-300
