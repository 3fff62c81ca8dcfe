# end-to-end C2 after kMaxChunks 64 -> 1024: smaller chunks, default (16 Ki bins, 8 threads) last; then walk/mc parity tests on the new chunking
mkdir -p gpurun_out
for cfg in "2048 8" "4096 8" "4096 16" "8192 8" "8192 16" "16384 16"; do
set -- $cfg
VB200_E2E_CHUNK_BINS=$1 VB200_HOST_THREADS=$2 timeout 120 python bench.py --no-cpu-baseline --steps 30 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('chunk $1 threads $2: resident %.4f ms  e2e %.4f ms (%.1f G)  pinned %.4f ms' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['value']/1e9, d['e2e']['pinned_ms_per_step']))"
done
for rep in 1 2; do
timeout 120 python bench.py --no-cpu-baseline --steps 30 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('default: resident %.4f ms  e2e %.4f ms (%.1f G)  pinned %.4f ms' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['value']/1e9, d['e2e']['pinned_ms_per_step']))"
done
timeout 120 python bench.py --workload c5 --no-cpu-baseline --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c5 default: resident %.4f ms  e2e %.4f ms (%.1f G)' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['value']/1e9))"
timeout 400 python -m pytest tests/test_gpu_mc.py tests/test_gpu_walk.py tests/test_gpu_full_size.py tests/test_gpu_examples.py -x -q -m gpu 2>&1 | tail -2
