#!/bin/bash
# usage: tools/build_variant.sh NAME "-DMACRO=1 ..." : builds _ab/libb200mc_NAME.so (ising.cu recompiled with the extra flags, the other
# objects taken from the regular build); run a tool against it with B200MC_SO=_ab/libb200mc_NAME.so
set -e
cd "$(dirname "$0")/../cuda_fortran_mc_simulation_spin_b200/csrc"
make -s
mkdir -p ../../_ab/build_$1
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-fvisibility=hidden -Xptxas -v $2 -c -o ../../_ab/build_$1/ising.o ising.cu 2> ../../_ab/build_$1/ising.log
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../_ab/libb200mc_$1.so ../../_ab/build_$1/ising.o build/ring.o build/ising_torus.o build/ising_bits.o build/clock.o build/sixclock.o build/xy.o build/xy_helical.o
grep -A3 "Compiling entry function.*ising_pass_kernelILi[46]ELi0ELb1ELb0ELb[01]ELb0E" ../../_ab/build_$1/ising.log | grep "Used\|spill" | tr '\n' ' '; echo
