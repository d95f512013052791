#!/bin/bash
# usage: run_variant.sh <variant .so>  -- ncu timing of the level-1 box launch with a variant library swapped in
# (experiment variants may produce unusable histograms: only the kernel under ncu counts, later failures are ignored)
cp sz3_b200/lib/libsz3b200.so /tmp/libsz3b200_orig.so
[ "$1" != "sz3_b200/lib/libsz3b200.so" ] && cp "$1" sz3_b200/lib/libsz3b200.so
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:k_interp_box -s 4 -c 1 python tools/prof_decompose.py 6 1 2>&1 | grep -E "duration|inst_executed|issue_active"
cp /tmp/libsz3b200_orig.so sz3_b200/lib/libsz3b200.so
