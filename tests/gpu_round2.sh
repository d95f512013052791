# GPU-box script (round 2): parity tests, bench lines of both arms, ncu launch list of a bench step and one full
# capture of the level-1 box launch.  usage: gpurun --timeout 2400 -- 'bash tests/gpu_round2.sh TAG'
TAG=${1:-r2}
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader; nproc
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -3 gpurun_out/bench_$TAG.err; cat gpurun_out/bench_$TAG.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>/dev/null; cat gpurun_out/bench_ref_$TAG.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_interp_box -s 4 -c 1 -o gpurun_out/prof_box_$TAG \
    python tools/prof_decompose.py 0 1 > gpurun_out/ncu_full_$TAG.log 2>&1
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k regex:'k_box_compact|k_interp_anchor|k_interp_box' -s 7 -c 7 -o gpurun_out/traffic_$TAG python tools/prof_decompose.py 0 2 > /dev/null 2>&1
