#!/bin/bash
# round-2 GPU call R (1 GPU): racecheck after the fixes; schedule knobs at full size
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== racecheck"; timeout 900 compute-sanitizer --tool racecheck python scripts/sanitize_check.py 2>&1 | grep -v "WARNING: " > gpurun_out/r2_sanitizer_racecheck.txt; tail -8 gpurun_out/r2_sanitizer_racecheck.txt
timeout 300 python -m pytest tests/test_gpu_multigpu.py tests/test_gpu_parity.py -x -q 2>&1 | tail -2
echo "== schedule knobs, KH 8192^2"
for spec in "96 220" "192 220" "192 350" "128 300" "256 300" "96 220"; do set -- $spec; FV2D_SCHED_HMAX=$1 FV2D_SCHED_C100=$2 scripts/bench_variants.sh main 2>/dev/null | sed "s/^/HMAX=$1 C=$2 /"; done
echo "== RT 16384^2"
for spec in "96 220" "192 300"; do set -- $spec; FV2D_SCHED_HMAX=$1 FV2D_SCHED_C100=$2 scripts/bench_variants.sh --workload rayleigh_taylor_16384_plm_hllc main 2>/dev/null | sed "s/^/HMAX=$1 C=$2 /"; done
