#!/usr/bin/env python
"""Developer tool (GPU box): a few fused steps of the main kernel families at sizes spanning
several strips and chunks, meant to be run under compute-sanitizer:
    compute-sanitizer --tool racecheck python scripts/sanitize_check.py
    compute-sanitizer --tool memcheck  python scripts/sanitize_check.py
"""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from conftest import load_golden  # noqa: E402
from fv2d_b200 import capi  # noqa: E402

CASES = [("kh_plm_128x64", {"mesh.Nx": 600, "mesh.Ny": 150}), ("c91_64x32", {"mesh.Nx": 300, "mesh.Ny": 90}),
         ("c91_plm_64x32", {"mesh.Nx": 260, "mesh.Ny": 40}), ("gresho_rk2_32", {"mesh.Nx": 280, "mesh.Ny": 70}),
         ("rt_fslp_32x96", {"mesh.Nx": 270, "mesh.Ny": 60}), ("blast_64", {"mesh.Nx": 300, "mesh.Ny": 64})]
os.environ["FV2D_CHUNK_ROWS"] = "11"
os.environ["FV2D_STREAM_ROWS"] = "16"
os.environ["FV2D_MAX_CTAS"] = "3"  # few CTAs, several work items each: the cross-item TMA streams are exercised
for name, ov in CASES:
    dev, run = capi.params_from_ini(load_golden(name).ini_path(), ov)
    Q0 = capi.init_problem(dev, run)
    with capi.Context(dev, run.time_stepping, run.epsilon_reset_negative) as ctx:
        ctx.upload_Q(Q0)
        ctx.prim_to_cons()
        ctx.compute_dt()
        ctx.run_steps(2)
        U = ctx.download_U()
    # the streamed host path: row blocks of 16 rows, partial sweeps, three streams
    a, b, hint = Q0.copy(), np.empty_like(Q0), 0.0
    with capi.Context(dev, run.time_stepping, run.epsilon_reset_negative) as ctx:
        for _ in range(3):
            _, hint, _ = ctx.advance_host_stream(a, b, hint)
            a, b = b, a
    print(name, "finite:", bool(np.all(np.isfinite(U))) and bool(np.all(np.isfinite(a))), flush=True)
