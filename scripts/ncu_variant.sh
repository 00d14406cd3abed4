#!/bin/bash
# Developer tool (GPU box): one `ncu --set full` capture of the fused sweep for a library
# variant -> gpurun_out/sweep_<name>.ncu-rep
#   scripts/ncu_variant.sh [--workload W] name [name...]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
WL=kelvin_helmholtz_8192_plm_hllc
if [ "$1" == "--workload" ]; then WL=$2; shift 2; fi
for n in "$@"; do
  lib=scratch/lib_$n.so; [ "$n" == "main" ] && lib=fv2d_b200/libfv2d_b200.so
  FV2D_B200_LIB=$PWD/$lib timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 4 -c 1 \
    -f -o gpurun_out/sweep_${WL}_$n python bench.py --workload $WL --steps 3 --warmup 3 --e2e-steps 0 --no-cpu-baseline --reps 1 --sustained-steps 0 --no-scaling-blocks > gpurun_out/ncu_$n.log 2>&1
  tail -2 gpurun_out/ncu_$n.log
done
