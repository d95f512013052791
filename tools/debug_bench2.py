# per-step wall times of the N-rank value / e2e legs (debug)
import os, sys, time, runpy
sys.argv = ["bench.py", "--gpus", os.environ.get("WORLD_SIZE", "1"), "--steps", "8", "--warmup", "3", "--no-extras", "--no-cpu-baseline", "--diag"]
import bench
orig = bench.run_ours
runpy.run_path("bench.py", run_name="__main__")
