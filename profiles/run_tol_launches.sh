ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r3_launches_tol.csv python profiles/run_full.py tol 3e-12 > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r3_launches_tol.csv')) if len(r)>5]
hdr=[i for i,r in enumerate(rows) if r[0]=='ID'][0]
H=rows[hdr]; k=H.index('Kernel Name'); v=H.index('Metric Value')
data=rows[hdr+1:]
half=len(data)//2
tot=collections.OrderedDict()
for r in data[half:]:
    n=r[k].split('(')[0][:70]
    tot.setdefault(n,[0,0.0]); tot[n][0]+=1; tot[n][1]+=float(r[v].replace(',',''))
print(sum(t for c,t in tot.values())/1e6)
for n,(c,t) in sorted(tot.items(), key=lambda x:-x[1][1])[:10]: print(f"{t/1e6:9.3f} ms x{c:4d}  {n}")
PY
