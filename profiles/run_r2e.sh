# round 2e: fast (fp32) bin walk + separable interpolant in the control-variate path
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_cv.py tests/test_gpu_full_size.py tests/test_gpu_fubini.py tests/test_gpu_examples.py -m gpu -q --durations=5 > gpurun_out/r2e_tests.log 2>&1; tail -12 gpurun_out/r2e_tests.log
timeout 300 python bench.py --workload c4 --no-cpu-baseline > gpurun_out/r2e_bench_c4.json 2> gpurun_out/r2e_bench_c4.err; tail -c 400 gpurun_out/r2e_bench_c4.json; tail -5 gpurun_out/r2e_bench_c4.err
BATCH=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2e_launches_c4.csv python profiles/run_full.py c4 > gpurun_out/r2e_c4_run.log 2>&1; tail -3 gpurun_out/r2e_c4_run.log
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r2e_launches_c4.csv')) if len(r)>5]
hdr=[i for i,r in enumerate(rows) if r[0]=='ID'][0]
H=rows[hdr]; k=H.index('Kernel Name'); v=H.index('Metric Value')
data=rows[hdr+1:]
half=len(data)//2
tot=collections.OrderedDict()
for r in data[half:]:
    n=r[k].split('(')[0][:70]
    tot.setdefault(n,[0,0.0]); tot[n][0]+=1; tot[n][1]+=float(r[v].replace(',',''))
print('total ms', sum(t for c,t in tot.values())/1e6)
for n,(c,t) in sorted(tot.items(), key=lambda x:-x[1][1])[:16]: print(f"{t/1e6:9.3f} ms x{c:4d}  {n}")
PY
