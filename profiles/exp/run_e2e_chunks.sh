# experiment: end-to-end chunk size (VB200_E2E_CHUNK_BINS) and helper threads
for cb in 65536 32768 16384 8192; do for th in 8 12; do
VB200_E2E_CHUNK_BINS=$cb VB200_HOST_THREADS=$th python bench.py --no-cpu-baseline --steps 40 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('chunk_bins $cb threads $th: value %.1f G  e2e %.1f G (%.4f ms)  pinned %.1f G'%(d['value']/1e9, d['e2e']['value']/1e9, d['e2e']['ms_per_step'], d['e2e']['pinned_value']/1e9))"
done; done
