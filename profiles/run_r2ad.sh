# round 2ad: where the end-to-end C2 call spends its time with the round-2 kernel (VB200_E2E_TRACE timestamps, us after the launch returned)
mkdir -p gpurun_out
for cfg in "16384 8" "16384 12" "8192 8" "32768 8"; do
set -- $cfg
echo "== chunk $1 threads $2"
VB200_E2E_TRACE=1 VB200_E2E_CHUNK_BINS=$1 VB200_HOST_THREADS=$2 timeout 120 python bench.py --no-cpu-baseline --no-cv --steps 20 --warmup 5 --sustain 0.2 2> gpurun_out/r2ad_trace_$1_$2.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('resident %.4f ms  e2e %.4f ms  pinned %.4f ms' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['pinned_ms_per_step']))"
grep "vb200 e2e" gpurun_out/r2ad_trace_$1_$2.err | tail -25 | head -6
done
