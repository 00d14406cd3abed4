#!/bin/bash
# Developer tool (GPU box): kernel-only bench line for each scratch/lib_*.so given by name.
#   [BENCH_EXTRA='--nx 8192 --ny 8192'] scripts/bench_variants.sh [--workload W] name1 name2 ...
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
WL=kelvin_helmholtz_8192_plm_hllc
if [ "$1" == "--workload" ]; then WL=$2; shift 2; fi
for n in "$@"; do
  lib=scratch/lib_$n.so; [ "$n" == "main" ] && lib=fv2d_b200/libfv2d_b200.so
  FV2D_B200_LIB=$PWD/$lib timeout 150 python bench.py --workload $WL --steps 40 --warmup 5 --e2e-steps 0 --no-cpu-baseline --reps 1 --sustained-steps 0 --no-scaling-blocks $BENCH_EXTRA 2>&1 | tail -1 > gpurun_out/var_${WL}_$n.json
  python - "$n" gpurun_out/var_${WL}_$n.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[2]))
    print(f"{sys.argv[1]:14s} {d['value']:9.0f} Mcell/s  frac={d['roofline']['frac']:.4f}  ms/launch={d['roofline']['ms_per_launch']:.4f}  clk={d['clocks']['sm_mhz']} {d['clocks']['reasons']} pw={d['clocks'].get('power_w_max')}")
except Exception as e:
    print(sys.argv[1], "FAILED", e, open(sys.argv[2]).read()[-400:])
PY
done
