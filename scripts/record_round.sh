#!/bin/bash
# Developer tool (GPU box, 1 GPU): the round's record of the shipped HEAD -> gpurun_out/<tag>_*: GPU suite, smoke,
# compute-sanitizer, the bench lines (driver-like and default), the reference arm, the other workloads, the ncu launch
# list and `ncu --set full` captures of the sweep, full-size parity against the reference binary.
#   scripts/record_round.sh [tag]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${1:-r2v9}
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/${T}_pytest_gpu_1gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v WARNING | tail -3
echo "== racecheck"; timeout 900 compute-sanitizer --tool racecheck python scripts/sanitize_check.py 2>&1 | grep -v WARNING | tail -9 | tee gpurun_out/${T}_sanitizer_racecheck.txt
echo "== memcheck"; timeout 900 compute-sanitizer --tool memcheck python scripts/sanitize_check.py 2>&1 | grep -v WARNING | tail -9 | tee gpurun_out/${T}_sanitizer_memcheck.txt
echo "== bench, driver-like (--steps 20 --warmup 5)"
timeout 900 python bench.py --steps 20 --warmup 5 2> gpurun_out/${T}_bench_steps20.err | tail -1 > gpurun_out/${T}_bench_steps20.json
echo "== bench, default"
timeout 900 python bench.py 2> gpurun_out/${T}_bench_default.err | tail -1 > gpurun_out/${T}_bench_default.json
echo "== reference arm"
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 2>> gpurun_out/${T}_bench_default.err | tail -1 > gpurun_out/${T}_bench_reference_arm.json
for wl in blast_4096_pcm_hllc rayleigh_taylor_16384_plm_hllc c91_8192_pcm_hllc_tc_visc; do
  timeout 600 python bench.py --workload $wl --steps 40 --warmup 5 --e2e-steps 0 --no-cpu-baseline --reps 1 --sustained-steps 0 --no-scaling-blocks 2>> gpurun_out/${T}_bench_default.err | tail -1 > gpurun_out/${T}_bench_$wl.json
done
for f in gpurun_out/${T}_bench*.json; do python - $f <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    r=d.get('roofline') or {}
    print(f"{sys.argv[1]}: {d['value']:.0f} {d['unit']} ms/step={d['ms_per_step']:.4f} frac={r.get('frac')} e2e={(d.get('e2e') or {}).get('value')} sustained={(d.get('sustained') or {}).get('frac')} clocks={d.get('clocks')}")
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
echo "== launch list"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_kh8192.csv \
  python bench.py --steps 4 --warmup 3 --e2e-steps 2 --no-cpu-baseline --reps 1 --sustained-steps 0 --no-scaling-blocks > gpurun_out/${T}_launches.log 2>&1
grep -c k_sweep gpurun_out/${T}_launches_kh8192.csv
echo "== ncu full"
for wl in kelvin_helmholtz_8192_plm_hllc c91_8192_pcm_hllc_tc_visc; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 24 -c 1 -f -o gpurun_out/${T}_sweep_$wl \
    python bench.py --workload $wl --steps 3 --warmup 3 --e2e-steps 0 --no-cpu-baseline --reps 1 --sustained-steps 0 --no-scaling-blocks > gpurun_out/${T}_ncu_$wl.log 2>&1
  tail -1 gpurun_out/${T}_ncu_$wl.log | cut -c1-160
done
ls -la gpurun_out/${T}*.ncu-rep
echo "== full-size parity against the reference binary"
timeout 1500 python scripts/parity_fullsize.py --out gpurun_out/${T}_parity_fullsize.json 2>&1 | grep -v WARNING | tail -12
