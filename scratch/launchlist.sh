CMD="python bench.py --steps 2 --warmup 1 --rays-per-gpu 131072 --no-extra --cpu-rays 256"
$CMD > gpurun_out/plain_ll.log 2>&1 || { tail -5 gpurun_out/plain_ll.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_ll.csv $CMD > gpurun_out/ncu_ll.log 2>&1
python - <<'PY'
import csv, io, collections
lines=open('gpurun_out/launches_ll.csv').read().splitlines()
start=[i for i,l in enumerate(lines) if l.startswith('"ID"')][0]
rd=csv.DictReader(io.StringIO("\n".join(lines[start:])))
agg=collections.defaultdict(lambda:[0,0.0]); order=[]
for r in rd:
    if r.get('Metric Name')!='gpu__time_duration.sum': continue
    name=r['Kernel Name'].split('(')[0][:60]; v=float(r['Metric Value'].replace(',','')); u=r['Metric Unit']
    v = v/1e3 if u=='ns' else (v*1e3 if u=='ms' else v)
    agg[name][0]+=1; agg[name][1]+=v
tot=sum(v[1] for v in agg.values())
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1])[:22]:
    print(f"{v[1]:12.1f} us {v[0]:4d}x {100*v[1]/tot:5.1f}%  {k}")
print('total', tot)
PY
