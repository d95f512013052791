# GPU-box script (diagnostics): plane-ordered copy tests, smoke(), timing, per-front times of the Lorenzo wavefront
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_compress.py -m gpu -x -q 2>&1 | tail -5
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 300 python tests/step_profile.py 512 6 2 2>&1 | sed -n 4,12p | cut -c1-230
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_lz.csv python tests/lz_one.py 256 0 > gpurun_out/ncu_lz.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_lz.csv')) if len(r)>8]
h=rows[0]; ik=h.index('Kernel Name'); iv=h.index('Metric Value'); ig=h.index('Grid Size'); iu=h.index('Metric Unit')
fr=[(r[ig], float(r[iv].replace(',','')), r[iu]) for r in rows[1:] if 'k_bw_front' in r[ik]]
print('k_bw_front launches', len(fr), 'sum', sum(v for _,v,_ in fr), fr[0][2] if fr else '')
sel=fr[len(fr)//2:]
for i in list(range(0,6))+list(range(len(sel)//2-2,len(sel)//2+2))+list(range(len(sel)-4,len(sel))):
    print(i, sel[i])
PY
