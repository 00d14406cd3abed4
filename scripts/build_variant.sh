#!/bin/bash
# Developer tool: build libfv2d_b200 with extra -D flags on the sweep kernel into
# scratch/lib_<name>.so (scratch/ is git-ignored but travels to the GPU box).
#   scripts/build_variant.sh nt128 -DFV2D_NT=128
set -e
name=$1; shift
cd "$(dirname "$0")/.."
mkdir -p scratch build
NVCC=/usr/local/cuda/bin/nvcc
ARCH="-gencode arch=compute_100a,code=sm_100a"
$NVCC -std=c++17 -O3 $ARCH -lineinfo -Xcompiler -fPIC -Xcudafe --diag_suppress=177 "$@" -Xptxas -v \
  -c fv2d_b200/csrc/fv2d_sweep.cu -o build/fv2d_sweep_$name.o 2> build/fv2d_sweep_$name.ptxas.log
make -s build/fv2d_ops.o build/fv2d_capi.o
$NVCC $ARCH -shared -o scratch/lib_$name.so build/fv2d_ops.o build/fv2d_sweep_$name.o build/fv2d_capi.o -cudart static -Xcompiler -fopenmp
grep -A1 "k_sweepILi[0-9]*ELb1ELi1ELi0ELb0ELb1" build/fv2d_sweep_$name.ptxas.log | grep -E "Used|spill" | tr '\n' ' '; echo
