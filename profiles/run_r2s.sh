# round 2s: occupancy sweep of the tile-major residual kernel (C4): samples per pass J x resident CTAs the kernel is compiled for
set -x
mkdir -p gpurun_out
cp viltrum_b200/libviltrum_b200.so /tmp/lib_default.so
run() { # tag, lib, J
  cp $2 viltrum_b200/libviltrum_b200.so
  VB200_CVT_J=$3 timeout 300 python bench.py --workload c4 --no-cpu-baseline --steps 10 --warmup 3 --sustain 0.5 2> gpurun_out/r2s_$1.err | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=j['roofline']
print('$1', 'step ms %.3f' % j['ms_per_step'], 'kernel ms/launch %.3f' % r['kernel_ms_per_launch'], 'frac %.3f' % r['frac'])" | tee -a gpurun_out/r2s_sweep.txt
}
: > gpurun_out/r2s_sweep.txt
run minb1_J64 /tmp/lib_default.so 64
run minb1_J32 /tmp/lib_default.so 32
run minb2_J64 viltrum_b200/build/libvariant_minb2.so 64
run minb2_J32 viltrum_b200/build/libvariant_minb2.so 32
run minb3_J32 viltrum_b200/build/libvariant_minb3.so 32
run minb3_J16 viltrum_b200/build/libvariant_minb3.so 16
run minb3_J64 viltrum_b200/build/libvariant_minb3.so 64
cp /tmp/lib_default.so viltrum_b200/libviltrum_b200.so
cat gpurun_out/r2s_sweep.txt
