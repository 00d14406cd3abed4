#!/usr/bin/env python
"""Quick on-GPU parity report: operator-level path and fused path vs the CPU oracle on every
golden fixture.  (Developer tool; the real checks are tests/test_gpu_*.py.)"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import oracle_lib as O  # noqa: E402
from conftest import GOLDEN_NAMES, Golden, rel_l1  # noqa: E402
from fv2d_b200 import capi  # noqa: E402


def main():
    bad = 0
    for name in GOLDEN_NAMES:
        g = Golden(name)
        dev, run = capi.params_from_ini(g.ini_path())
        Q0 = capi.init_problem(dev, run)
        n = g.nsteps
        # --- operator-level path, host-driven loop like main.cpp:62-84
        with capi.Context(dev, run.time_stepping, run.epsilon_reset_negative) as ctx:
            ctx.upload_Q(Q0)
            ctx.prim_to_cons()
            dts = []
            for _ in range(n):
                dt, _inv = ctx.compute_dt()
                dts.append(dt)
                ctx.update(dt)
                ctx.cons_to_prim()
                ctx.check_negatives()
            Qo, Uo = ctx.download_Q(), ctx.download_U()
        dts = np.array(dts)
        ops_exact = (np.array_equal(dts, g.dts) and np.array_equal(O.domain(dev, Qo), g.QN)
                     and np.array_equal(O.domain(dev, Uo), g.UN))
        ops_l1 = rel_l1(O.domain(dev, Uo), g.UN)
        # --- fused path, device-resident dt
        with capi.Context(dev, run.time_stepping, run.epsilon_reset_negative) as ctx:
            ctx.upload_Q(Q0)
            ctx.prim_to_cons()
            ctx.compute_dt()
            ctx.run_steps(n)
            Qf, Uf = ctx.download_Q(), ctx.download_U()
            fdts = ctx.dt_history(n)
            t, _, steps = ctx.get_time()
        f_l1u = max(rel_l1(O.domain(dev, Uf)[f], g.UN[f]) for f in (0, 3))
        f_l1q = rel_l1(O.domain(dev, Qf), g.QN)
        f_dt = float(np.max(np.abs(fdts - g.dts) / g.dts)) if len(fdts) == n else float("nan")
        ok = ops_exact and f_l1u <= 1e-12 and f_l1q <= 1e-12 and f_dt <= 1e-13
        bad += not ok
        print(f"{name:20s} ops bit-exact={ops_exact} (L1 {ops_l1:.1e}) | fused relL1 U={f_l1u:.2e} Q={f_l1q:.2e} "
              f"dt={f_dt:.2e} steps={steps} {'OK' if ok else 'FAIL'}", flush=True)
    print("FAILED" if bad else "ALL OK", bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
