#!/usr/bin/env python
"""Developer tool (GPU box): where a persistent sweep spends its time, per CTA (needs a library
variant built with -DFV2D_TIMING: scripts/build_variant.sh timing -DFV2D_TIMING).
    FV2D_B200_LIB=$PWD/scratch/lib_timing.so python scripts/sweep_timing.py [workload] [steps]"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from fv2d_b200 import capi  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "kelvin_helmholtz_8192_plm_hllc"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
base, ov = bench.WORKLOADS[wl]
ov = dict(ov)
if len(sys.argv) > 3:  # rows of the slab (development: what one of N y-slabs sees)
    ov["mesh.Ny"] = int(sys.argv[3])
dev, run = capi.params_from_ini(ROOT / "settings" / base, ov)
Q0 = capi.init_problem(dev, run)
with capi.Context(dev, run.time_stepping, run.epsilon_reset_negative) as ctx:
    ctx.upload_Q(Q0)
    ctx.prim_to_cons()
    ctx.compute_dt()
    ctx.run_steps(steps)
    t = ctx.sweep_timing(296).astype(np.float64)
t = t[t[:, 3] > 0]
us = 1.0 / 1965.0  # cycles -> microseconds at the boost clock
raw = t[:, 2].astype(np.int64)
gap, loop, items, rows_cta, tot = t[:, 0] * us, t[:, 1] * us, (raw & 0xffff).astype(float), (raw >> 16).astype(float), t[:, 3] * us
print(f"{wl}: {len(t)} CTAs; items/CTA mean {items.mean():.1f} min {items.min():.0f} max {items.max():.0f}")
print(f"  total    mean {tot.mean():8.1f} us  min {tot.min():8.1f}  max {tot.max():8.1f}")
print(f"  in loops mean {loop.mean():8.1f} us  min {loop.min():8.1f}  max {loop.max():8.1f}")
print(f"  outside  mean {gap.mean():8.1f} us  min {gap.min():8.1f}  max {gap.max():8.1f}   per item {(gap / items).mean():.2f} us")
rows = dev.Ny * ((dev.Nx + 251) // 252) / len(t)
print(f"  rows/CTA {rows:.0f} (min {rows_cta.min():.0f} max {rows_cta.max():.0f}): {loop.mean() / rows:.3f} us per row inside the loops; "
      f"per-CTA us/row min {(loop / rows_cta).min():.3f} max {(loop / rows_cta).max():.3f}")
slow = np.argsort(tot)[-5:]
for k in slow:
    print(f"    slowest: total {tot[k]:.1f} us items {items[k]:.0f} rows {rows_cta[k]:.0f} us/row {loop[k] / rows_cta[k]:.3f}")
