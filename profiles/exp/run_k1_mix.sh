#!/bin/bash
# runs every profiles/exp/bin/k1_mix_* twice (order reversed the second time) -> gpurun_out/k1_mix.txt
mkdir -p gpurun_out; out=gpurun_out/k1_mix${1:+_$1}.txt; : > $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> $out
for b in $(ls profiles/exp/bin/k1_mix_*); do $b 30 >> $out; done
for b in $(ls -r profiles/exp/bin/k1_mix_*); do $b 30 >> $out; done
cat $out
