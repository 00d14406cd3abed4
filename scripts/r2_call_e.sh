#!/bin/bash
# round-2 GPU call E/F (1 GPU): persistent sweep variants
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py tests/test_gpu_fuzz.py -x -q 2>&1 | tail -4
echo "== timing KH auto"; FV2D_B200_LIB=$PWD/scratch/lib_timing.so timeout 100 python scripts/sweep_timing.py
echo "== timing blast auto"; FV2D_B200_LIB=$PWD/scratch/lib_timing.so timeout 100 python scripts/sweep_timing.py blast_4096_pcm_hllc
echo "== variants"
scripts/bench_variants.sh main main
for wl in blast_4096_pcm_hllc c91_8192_pcm_hllc_tc_visc rayleigh_taylor_16384_plm_hllc; do scripts/bench_variants.sh --workload $wl main; done
echo "== KH 8 slabs' worth: 8192 x 1024"; BENCH_EXTRA="--ny 1024" scripts/bench_variants.sh main main
