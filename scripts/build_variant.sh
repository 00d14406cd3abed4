#!/bin/bash
# Developer tool: build libfv2d_b200 with extra -D flags on the sweep kernel into
# scratch/lib_<name>.so (scratch/ is git-ignored but travels to the GPU box).
#   scripts/build_variant.sh nt128 -DFV2D_NT=128
set -e
name=$1; shift
cd "$(dirname "$0")/.."
mkdir -p scratch build/var_$name
make -s build/fv2d_ops.o build/fv2d_stream.o build/fv2d_capi.o
make -s -j4 OBJDIR=build/var_$name SWEEPFLAGS="$*" build/var_$name/fv2d_sweep.o build/var_$name/fv2d_sweep_s0.o build/var_$name/fv2d_sweep_s1.o build/var_$name/fv2d_sweep_s2.o
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o scratch/lib_$name.so build/fv2d_ops.o build/fv2d_stream.o build/fv2d_capi.o \
  build/var_$name/fv2d_sweep.o build/var_$name/fv2d_sweep_s0.o build/var_$name/fv2d_sweep_s1.o build/var_$name/fv2d_sweep_s2.o -cudart static -Xcompiler -fopenmp
grep -A1 "k_sweepILi[0-9]*ELb1ELi1ELi0ELb0ELi0E" build/var_$name/fv2d_sweep_s1.ptxas.log | grep -E "Used|spill" | tr '\n' ' '; echo
