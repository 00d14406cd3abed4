#!/bin/bash
# round-2 GPU call Q (1 GPU): racecheck (full log) + ncu evidence of the shipped sweep
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== racecheck"; timeout 900 compute-sanitizer --tool racecheck python scripts/sanitize_check.py 2>&1 | grep -v "WARNING: " > gpurun_out/r2_sanitizer_racecheck.txt; tail -5 gpurun_out/r2_sanitizer_racecheck.txt; grep -c "Race reported" gpurun_out/r2_sanitizer_racecheck.txt
grep "Race reported" -A1 gpurun_out/r2_sanitizer_racecheck.txt | grep -o "fv2d_sweep.cu:[0-9]*\|Error\|Warning" | sort | uniq -c | sort -rn | head -20
echo "== ncu full"
for wl in kelvin_helmholtz_8192_plm_hllc c91_8192_pcm_hllc_tc_visc; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 19 -c 1 -f -o gpurun_out/r2_sweep_$wl \
    python bench.py --workload $wl --steps 3 --warmup 3 --e2e-steps 0 --no-cpu-baseline --reps 1 --sustained-steps 0 --no-scaling-blocks > gpurun_out/r2_ncu_$wl.log 2>&1
  grep -E "PROF|Report" gpurun_out/r2_ncu_$wl.log | tail -2 | cut -c1-200
done
ls -la gpurun_out/r2_sweep_*.ncu-rep
