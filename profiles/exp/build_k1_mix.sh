#!/bin/bash
# builds profiles/exp/bin/k1_mix_<variant> (sm_100a) for every generator mix; run_k1_mix.sh runs them on the GPU box
cd "$(dirname "$0")/../.." && mkdir -p profiles/exp/bin
build() {  # name, flags...
  name=$1; shift
  nvcc -std=c++17 -O3 -lineinfo --expt-relaxed-constexpr -gencode arch=compute_100a,code=sm_100a -I include "$@" -DVARIANT="\"$name\"" \
      -o profiles/exp/bin/k1_mix_$name profiles/exp/k1_mix.cu &
}
rm -f profiles/exp/bin/k1_mix_*
for b in 1 2 3 4; do build xoshiro_minb$b -DK1_RNG=0 -DVB200_MC_MINB=$b; done
if [ "$1" = xo ]; then
  build p5t0_r12 -DVB200_MC_TF_NUM=0 -DVB200_MC_MINB=1
  for b in 2 3 4; do build xoshiro_pipe_minb$b -DK1_RNG=0 -DVB200_MC_MINB=$b -DVB200_MC_PIPELINE=1; build xoshiro_ch4_minb$b -DK1_RNG=0 -DVB200_MC_MINB=$b -DVB200_MC_CHAINS=4;
    build xoshiro_pipe_ch4_minb$b -DK1_RNG=0 -DVB200_MC_MINB=$b -DVB200_MC_CHAINS=4 -DVB200_MC_PIPELINE=1; done
  build philox_pipe_minb3 -DK1_RNG=1 -DVB200_MC_MINB=3 -DVB200_MC_PIPELINE=1
  wait; ls profiles/exp/bin; exit 0; fi
for n in 0 1 2 3 4 5; do build p$((5-n))t${n}_r12 -DVB200_MC_MINB=1 -DVB200_MC_TF_NUM=$n -DVB200_MC_TF_ROUNDS=12; done
wait
for n in 1 2 3; do build p$((5-n))t${n}_r20 -DVB200_MC_MINB=1 -DVB200_MC_TF_NUM=$n -DVB200_MC_TF_ROUNDS=20; done
for n in 2 3; do for b in 2 3; do build p$((5-n))t${n}_r12_minb$b -DVB200_MC_TF_NUM=$n -DVB200_MC_TF_ROUNDS=12 -DVB200_MC_MINB=$b; done; done
wait
ls -la profiles/exp/bin
