# GPU-box script (diagnostics): smoke(), lossless-stage timing, per-front times of the Lorenzo wavefront
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 300 python tests/step_profile.py 512 6 2 2>&1 | sed -n 4,6p | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_lz.csv python tests/lz_one.py 256 0 > gpurun_out/ncu_lz.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_lz.csv')) if len(r)>8]
h=rows[0]; ik=h.index('Kernel Name'); iv=h.index('Metric Value'); ig=h.index('Grid Size')
fr=[(r[ig], float(r[iv].replace(',',''))) for r in rows[1:] if 'k_bw_front' in r[ik]]
print('k_bw_front launches', len(fr), 'total us', sum(v for _,v in fr)/1000 if fr and fr[0][1]>1000 else sum(v for _,v in fr))
half=len(fr)//2
sel=fr[half:]  # second compression
for i in list(range(0,6))+list(range(len(sel)//2-2,len(sel)//2+2))+list(range(len(sel)-4,len(sel))):
    print(i, sel[i])
PY
