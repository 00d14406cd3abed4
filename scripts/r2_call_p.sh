#!/bin/bash
# round-2 GPU call P (1 GPU): decomposition-independence test, compute-sanitizer, ncu evidence of the shipped sweep
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_properties.py -x -q 2>&1 | tail -3
echo "== racecheck"; timeout 900 compute-sanitizer --tool racecheck python scripts/sanitize_check.py 2>&1 | grep -v WARNING | tail -12 | tee gpurun_out/r2_sanitizer_racecheck.txt
echo "== memcheck"; timeout 900 compute-sanitizer --tool memcheck python scripts/sanitize_check.py 2>&1 | grep -v WARNING | tail -12 | tee gpurun_out/r2_sanitizer_memcheck.txt
echo "== launch list"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_kh8192.csv \
  python bench.py --steps 4 --warmup 3 --e2e-steps 0 --no-cpu-baseline --reps 1 --sustained-steps 0 --no-scaling-blocks > gpurun_out/r2_launches.log 2>&1
grep -c k_sweep gpurun_out/r2_launches_kh8192.csv
echo "== ncu full"
for wl in kelvin_helmholtz_8192_plm_hllc c91_8192_pcm_hllc_tc_visc; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 24 -c 1 -f -o gpurun_out/r2_sweep_$wl \
    python bench.py --workload $wl --steps 3 --warmup 3 --e2e-steps 0 --no-cpu-baseline --reps 1 --sustained-steps 0 --no-scaling-blocks > gpurun_out/r2_ncu_$wl.log 2>&1
  tail -1 gpurun_out/r2_ncu_$wl.log | cut -c1-200
done
ls -la gpurun_out/*.ncu-rep | tail -3
