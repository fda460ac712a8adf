# per-kernel durations of one frame (frame 20 of cfg4), about 1 GPU-minute
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -s 339 -c 20 --csv --log-file gpurun_out/launch_list_cur.csv python scratch/prof_run.py cfg4 21 > gpurun_out/launch_list_cur.log 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/launch_list_cur.csv')))
hi=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
h=rows[hi]; ik=h.index('Kernel Name'); iv=h.index('Metric Value')
tot=0
for r in rows[hi+1:]:
    if len(r)<len(h): continue
    name=r[ik].split('(')[0].replace('<unnamed>::','').replace('void ','')
    v=float(r[iv].replace(',',''))/1000; tot+=v
    print(f"{name:28s} {v:8.2f} us")
print('sum', round(tot,1))
PY
