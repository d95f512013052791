# GPU-box experiment: how the bench line depends on the host cores one rank gets (stands in for 8 ranks on one host).
# usage: gpurun --timeout 900 -- 'bash tests/gpu_cores.sh'
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nproc
run() {  # label, cpu list, extra args
    timeout 200 taskset -c $2 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $3 2>/dev/null |
        python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', 'cores', '$2', round(d['value'],1), round(d['e2e']['value'],1), d['config']['host'])"
}
run old_spin16 0-1 "--host-threads 16 --host-wait 0"
run spin2 0-1 "--host-threads 2 --host-wait 0"
run yield2 0-1 "--host-threads 2 --host-wait 1"
run yield16 0-1 "--host-threads 16 --host-wait 1"
run yield_all 0-15 "--host-wait 1"
run spin_all 0-15 "--host-wait 0"
