#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== 2-D copy probe"
timeout 300 ./scratch/pcie2d 2>&1 | tee gpurun_out/r2_pcie2d_probe.txt
echo "== timing build, 1024-row slab"
FV2D_B200_LIB=$PWD/scratch/lib_timing.so timeout 300 python scripts/sweep_timing.py kelvin_helmholtz_8192_plm_hllc 6 1024 2>&1 | grep -v WARNING
FV2D_B200_LIB=$PWD/scratch/lib_timing.so timeout 300 python scripts/sweep_timing.py kelvin_helmholtz_8192_plm_hllc 6 2>&1 | grep -v WARNING
