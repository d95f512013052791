# GPU-box script (1 GPU): the N = 1 bench under different host pool sizes.
cd $GRAFT_REPO_ROOT
for t in 0 12 8 6; do
  echo "== host-threads $t"
  timeout 200 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline $( [ $t != 0 ] && echo --host-threads $t ) 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), 'ms', round(d['e2e']['ms_per_step'],2), d['stages_ms'].get('tune_wall'))"
done
