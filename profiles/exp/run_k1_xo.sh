#!/bin/bash
# xoshiro-stream variants of K1 against the Philox kernel, then one ncu --set full capture of each
mkdir -p gpurun_out; out=gpurun_out/k1_mix_r2b.txt; : > $out
for rep in 1 2; do for b in profiles/exp/bin/k1_mix_*; do $b 30 >> $out; done; done
cat $out
for v in xoshiro_minb1 xoshiro_minb2 p5t0_r12; do
  timeout 120 ncu --set full --clock-control none --import-source on -k regex:mc_per_bin -s 6 -c 1 -o gpurun_out/r2b_k1_$v -f profiles/exp/bin/k1_mix_$v 3 > gpurun_out/r2b_ncu_$v.log 2>&1; tail -2 gpurun_out/r2b_ncu_$v.log
done
