# end-to-end C2 (host bins through vb200_mc_per_bin): chunk size x host threads sweep
mkdir -p gpurun_out
for chunk in 16384 32768 65536 131072 262144; do
for thr in 4 8 16; do
VB200_E2E_CHUNK_BINS=$chunk VB200_HOST_THREADS=$thr timeout 120 python bench.py --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/r1d_e2e_${chunk}_${thr}.json 2>/dev/null
python - <<PY
import json
d = json.loads(open('gpurun_out/r1d_e2e_${chunk}_${thr}.json').read().strip().splitlines()[-1])
print('chunk', $chunk, 'threads', $thr, 'resident %.1f' % (d['value'] / 1e9), 'e2e %.1f G (%.4f ms)' % (d['e2e']['value'] / 1e9, d['e2e']['ms_per_step']), 'pinned %.1f' % (d['e2e']['pinned_value'] / 1e9))
PY
done
done
nproc
