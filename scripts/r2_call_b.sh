#!/bin/bash
# round-2 GPU call B (1 GPU): first run of the persistent sweep: smoke, GPU test suite, variants
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== pytest"
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_fullsize_reference.py 2>&1 | tail -15
echo "== variants"
scripts/bench_variants.sh main nouni main nouni
for wl in blast_4096_pcm_hllc c91_8192_pcm_hllc_tc_visc rayleigh_taylor_16384_plm_hllc; do scripts/bench_variants.sh --workload $wl main; done
timeout 300 python bench.py --steps 200 --warmup 10 --e2e-steps 0 --no-cpu-baseline --reps 1 --sustained-steps 0 --no-scaling-blocks 2>&1 | tail -1 > gpurun_out/r2_v7_sustained200.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_v7_sustained200.json'))
print('sustained200', d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['ms_per_launch'], d['gpu_launches'], d['clocks'])
PY
