# where the end-to-end C2 call spends its time: VB200_E2E_TRACE timestamps (us after the launch returned) for a few chunk/thread settings
mkdir -p gpurun_out
for cfg in "65536 8" "16384 16" "16384 8" "8192 16"; do
set -- $cfg
echo "== chunk $1 threads $2"
VB200_E2E_TRACE=1 VB200_E2E_CHUNK_BINS=$1 VB200_HOST_THREADS=$2 timeout 120 python bench.py --no-cpu-baseline --steps 20 --warmup 5 2> gpurun_out/r1d_trace_$1_$2.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('resident %.4f ms  e2e %.4f ms  pinned %.4f ms' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['pinned_ms_per_step']))"
grep "vb200 e2e" gpurun_out/r1d_trace_$1_$2.err | tail -25 | head -8
done
