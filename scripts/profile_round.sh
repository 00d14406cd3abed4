#!/bin/bash
# Developer tool (GPU box, 1 GPU): everything profiles/ needs for one kernel version ->
# gpurun_out/<tag>_*: the default bench line, the reference arm, one kernel-only line per other
# workload, the ncu launch list of the default bench command and `ncu --set full` captures of the
# fused sweep (KH 8192^2 and C91 8192^2).
#   scripts/profile_round.sh v4
tag=${1:-vX}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python bench.py 2> gpurun_out/${tag}_bench.err | tail -1 > gpurun_out/${tag}_bench.json
python bench.py --impl reference --steps 10 --warmup 3 2>> gpurun_out/${tag}_bench.err | tail -1 > gpurun_out/${tag}_bench_reference_arm.json
for wl in blast_4096_pcm_hllc rayleigh_taylor_16384_plm_hllc c91_8192_pcm_hllc_tc_visc; do
  python bench.py --workload $wl --steps 40 --warmup 5 --e2e-steps 0 --no-cpu-baseline --reps 1 --sustained-steps 0 --no-scaling-blocks 2>> gpurun_out/${tag}_bench.err | tail -1 > gpurun_out/${tag}_bench_$wl.json
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_kh8192.csv \
  python bench.py --steps 4 --warmup 3 --e2e-steps 0 --no-cpu-baseline --reps 1 --sustained-steps 0 --no-scaling-blocks > gpurun_out/${tag}_launches.log 2>&1
for wl in kelvin_helmholtz_8192_plm_hllc c91_8192_pcm_hllc_tc_visc; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 4 -c 1 -f -o gpurun_out/${tag}_sweep_$wl \
    python bench.py --workload $wl --steps 3 --warmup 3 --e2e-steps 0 --no-cpu-baseline --reps 1 --sustained-steps 0 --no-scaling-blocks > gpurun_out/${tag}_ncu_$wl.log 2>&1
done
for f in gpurun_out/${tag}_bench*.json; do python - $f <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    r=d.get('roofline') or {}
    print(f"{sys.argv[1]}: {d['value']:.0f} {d['unit']} ms/step={d['ms_per_step']:.4f} frac={r.get('frac')} e2e={(d.get('e2e') or {}).get('value')} clocks={d.get('clocks')}")
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
