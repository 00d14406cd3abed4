#!/bin/bash
# round-2 GPU call G (1 GPU): persistent sweep, late binding + round schedule
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py tests/test_gpu_fuzz.py -x -q 2>&1 | tail -3
echo "== timing KH auto"; FV2D_B200_LIB=$PWD/scratch/lib_timing.so timeout 100 python scripts/sweep_timing.py 2>/dev/null
echo "== timing blast auto"; FV2D_B200_LIB=$PWD/scratch/lib_timing.so timeout 100 python scripts/sweep_timing.py blast_4096_pcm_hllc 2>/dev/null
echo "== variants"
scripts/bench_variants.sh main main 2>/dev/null
for wl in blast_4096_pcm_hllc c91_8192_pcm_hllc_tc_visc rayleigh_taylor_16384_plm_hllc; do scripts/bench_variants.sh --workload $wl main 2>/dev/null; done
echo "== KH 8192 x 1024"; BENCH_EXTRA="--ny 1024" scripts/bench_variants.sh main main 2>/dev/null
for c in 150 300; do echo "== C100=$c"; FV2D_SCHED_C100=$c BENCH_EXTRA="--ny 1024" scripts/bench_variants.sh main 2>/dev/null; FV2D_SCHED_C100=$c scripts/bench_variants.sh --workload blast_4096_pcm_hllc main 2>/dev/null; done
echo "== HMIN=16"; FV2D_SCHED_HMIN=16 BENCH_EXTRA="--ny 1024" scripts/bench_variants.sh main 2>/dev/null; FV2D_SCHED_HMIN=16 scripts/bench_variants.sh --workload blast_4096_pcm_hllc main 2>/dev/null
