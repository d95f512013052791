cd $GRAFT_REPO_ROOT
for cfg in "4 4 1" "8 4 1" "8 8 1" "4 4 2" "2 2 1" "4 2 2"; do
set -- $cfg
echo "== tiny=$1 small=$2 mid=$3"
SZ3B_BOX_SPLIT_TINY=$1 SZ3B_BOX_SPLIT_SMALL=$2 SZ3B_BOX_SPLIT_MID=$3 ncu --metrics gpu__time_duration.sum --clock-control none -c 10 --csv python tools/prof_decompose.py 0 1 2>/dev/null | grep -E "k_interp_box" | awk -F'","' '{print $5, $NF}' | tr -d '"' | tr '\n' ';'
echo
done
