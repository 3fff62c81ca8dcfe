#!/bin/bash
# builds viltrum_b200/build/libvariant_minb<N>.so: the library with the tile-major residual kernel compiled for N resident CTAs per SM (experiment)
cd "$(dirname "$0")/../.." || exit 1
for n in "$@"; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo --expt-relaxed-constexpr -Xcompiler -fPIC -I include --fmad=false -DVB200_CVT_MINB=$n \
      -Xptxas -v -c viltrum_b200/csrc/cv.cu -o viltrum_b200/build/cv_minb$n.o 2> viltrum_b200/build/cv_minb$n.ptxas.txt || exit 1
  objs=$(ls viltrum_b200/build/*.o | grep -v "cv_minb" | grep -v "/cv.o")
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o viltrum_b200/build/libvariant_minb$n.so $objs viltrum_b200/build/cv_minb$n.o -lcudart -ldl || exit 1
  grep -A2 "cv_tile_samples_kernelILi3ELi5" viltrum_b200/build/cv_minb$n.ptxas.txt | grep -E "spill|registers"
done
