# experiment harness (not product): bench.py against variant builds of the library (profiles/exp/libv_*.so)
mkdir -p gpurun_out
cp viltrum_b200/libviltrum_b200.so /tmp/lib_product.so
for v in product "$@"; do
  if [ "$v" = product ]; then cp /tmp/lib_product.so viltrum_b200/libviltrum_b200.so; else cp profiles/exp/libv_$v.so viltrum_b200/libviltrum_b200.so; fi
  python bench.py --no-cpu-baseline --steps 30 --warmup 5 > gpurun_out/var_$v.json 2>&1
  python - "$v" <<'PY'
import json,sys
v=sys.argv[1]
try:
    d=json.loads(open(f'gpurun_out/var_{v}.json').read().strip().splitlines()[-1])
    print(v, 'value %.1f G'%(d['value']/1e9), 'ms %.4f'%d['ms_per_step'], 'e2e %.1f G'%(d['e2e']['value']/1e9), 'frac %.4f'%d['roofline']['frac'])
except Exception as e:
    print(v, 'FAILED', e)
PY
done
cp /tmp/lib_product.so viltrum_b200/libviltrum_b200.so
