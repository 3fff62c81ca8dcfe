# experiment: window walk kernel compiled with __launch_bounds__(256, 5) (48 registers, 5 CTAs/SM) against the shipped 4-CTA build
set -x
mkdir -p gpurun_out
for rep in 1 2; do
timeout 200 python bench.py --workload c5 --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/r1d_occ4_$rep.json 2>/dev/null
done
cp viltrum_b200/libviltrum_b200.so /tmp/orig.so
cp profiles/exp/occ5/libviltrum_b200_occ5.so viltrum_b200/libviltrum_b200.so
for rep in 1 2; do
timeout 200 python bench.py --workload c5 --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/r1d_occ5_$rep.json 2>/dev/null
done
timeout 300 python -m pytest tests/test_gpu_walk.py -x -q -m gpu 2>&1 | tail -2
cp /tmp/orig.so viltrum_b200/libviltrum_b200.so
python - <<'PY'
import json
for n in ("occ4_1", "occ4_2", "occ5_1", "occ5_2"):
    try:
        d = json.loads(open(f'gpurun_out/r1d_{n}.json').read().strip().splitlines()[-1])
        print(n, 'c5 value %.1f G paths/s' % (d['value'] / 1e9), 'ms', d['ms_per_step'], 'e2e %.1f' % (d['e2e']['value'] / 1e9))
    except Exception as e:
        print(n, "failed", e)
PY
