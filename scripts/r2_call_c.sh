#!/bin/bash
# round-2 GPU call C (1 GPU): why is the persistent sweep slower?  chunk sweep, in-kernel timing, ncu
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== KH chunk sweep"; scripts/chunk_sweep.sh auto 48 93 186 372 1024
echo "== timing KH auto"; FV2D_B200_LIB=$PWD/scratch/lib_timing.so timeout 100 python scripts/sweep_timing.py
echo "== timing KH 1024"; FV2D_CHUNK_ROWS=1024 FV2D_B200_LIB=$PWD/scratch/lib_timing.so timeout 100 python scripts/sweep_timing.py
echo "== timing blast auto"; FV2D_B200_LIB=$PWD/scratch/lib_timing.so timeout 100 python scripts/sweep_timing.py blast_4096_pcm_hllc
echo "== timing blast 241"; FV2D_CHUNK_ROWS=241 FV2D_B200_LIB=$PWD/scratch/lib_timing.so timeout 100 python scripts/sweep_timing.py blast_4096_pcm_hllc
echo "== blast chunks"; for cr in auto 80 241; do if [ $cr == auto ]; then unset FV2D_CHUNK_ROWS; else export FV2D_CHUNK_ROWS=$cr; fi; scripts/bench_variants.sh --workload blast_4096_pcm_hllc main; done; unset FV2D_CHUNK_ROWS
echo "== ncu"; scripts/ncu_variant.sh main
