#!/bin/bash
# round-2 GPU call L (1 GPU): U^{n+1} through TMA stores, A/B against per-thread stores
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py tests/test_gpu_fuzz.py tests/test_gpu_multigpu.py -x -q 2>&1 | tail -3
scripts/bench_variants.sh main nostore main nostore 2>/dev/null
for wl in blast_4096_pcm_hllc c91_8192_pcm_hllc_tc_visc rayleigh_taylor_16384_plm_hllc; do scripts/bench_variants.sh --workload $wl main nostore 2>/dev/null; done
echo "== sustained (200 steps)"
for v in main nostore; do lib=scratch/lib_$v.so; [ $v == main ] && lib=fv2d_b200/libfv2d_b200.so
FV2D_B200_LIB=$PWD/$lib python bench.py --steps 200 --warmup 10 --e2e-steps 0 --no-cpu-baseline --reps 1 --sustained-steps 0 --no-scaling-blocks 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$v', round(d['value']), round(d['ms_per_step'],4), round(d['roofline']['frac'],4), d['clocks'])"
done
