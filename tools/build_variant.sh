#!/bin/bash
# usage: tools/build_variant.sh NAME "-DMACRO=1 ..." [source.cu] : builds _ab/libb200mc_NAME.so (one translation unit, default ising.cu,
# recompiled with the extra flags; the other objects come from the regular build); run a tool against it with B200MC_SO=_ab/libb200mc_NAME.so
set -e
cd "$(dirname "$0")/../cuda_fortran_mc_simulation_spin_b200/csrc"
SRC=${3:-ising.cu}
BASE=${SRC%.cu}
make -s
mkdir -p ../../_ab/build_$1
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-fvisibility=hidden -Xptxas -v $2 -c -o ../../_ab/build_$1/$BASE.o $SRC 2> ../../_ab/build_$1/$BASE.log
OBJS=""
for o in ring ising ising_torus ising_bits clock sixclock xy xy_helical; do
  if [ "$o" = "$BASE" ]; then OBJS="$OBJS ../../_ab/build_$1/$o.o"; else OBJS="$OBJS build/$o.o"; fi
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../_ab/libb200mc_$1.so $OBJS
grep -c "spill stores" ../../_ab/build_$1/$BASE.log
