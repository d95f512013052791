# GPU-box script (2 GPUs): the N = 2 bench with the per-step trace, under a few host settings.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() {
  echo "== $PIN $*"
  SZ3B_BENCH_PIN_CORES=$PIN SZ3B_STEP_TRACE=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) \
     bench.py --gpus 2 --steps 8 --warmup 3 --no-extras --no-cpu-baseline "$@" 2> gpurun_out/probe.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('value', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), 'ms', round(d['e2e']['ms_per_step'],2), d['config']['host'])"
  grep trace gpurun_out/probe.err | awk '{print $3, $6}' | tr '\n' ';'; echo
}
PIN= run
PIN= run
PIN=6 run
PIN= run --no-pin
