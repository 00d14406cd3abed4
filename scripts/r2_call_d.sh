#!/bin/bash
# round-2 GPU call D (1 GPU): persistent sweep with depth-1 reservation and inline next-item staging
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py tests/test_gpu_fuzz.py -x -q 2>&1 | tail -4
echo "== timing KH auto"; FV2D_B200_LIB=$PWD/scratch/lib_timing.so timeout 100 python scripts/sweep_timing.py
echo "== timing blast auto"; FV2D_B200_LIB=$PWD/scratch/lib_timing.so timeout 100 python scripts/sweep_timing.py blast_4096_pcm_hllc
echo "== variants"
scripts/bench_variants.sh main main
for wl in blast_4096_pcm_hllc c91_8192_pcm_hllc_tc_visc rayleigh_taylor_16384_plm_hllc; do scripts/bench_variants.sh --workload $wl main; done
echo "== KH schedule knobs"
for c in 100 200 300; do FV2D_SCHED_C100=$c scripts/bench_variants.sh main | sed "s/^/C100=$c /"; done
for h in 64 128 192; do FV2D_SCHED_HMAX=$h scripts/bench_variants.sh main | sed "s/^/HMAX=$h /"; done
echo "== blast schedule knobs"
for c in 100 200 300; do FV2D_SCHED_C100=$c scripts/bench_variants.sh --workload blast_4096_pcm_hllc main | sed "s/^/C100=$c /"; done
