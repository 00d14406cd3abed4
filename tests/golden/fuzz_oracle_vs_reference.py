#!/usr/bin/env python
"""Differential fuzz of the CPU oracle against THE REFERENCE ITSELF (dev container only).

tests/golden/*.npz pin the oracle (oracle/fv2d_oracle.c) and the host Init / .ini mirrors on 22
hand-picked configurations.  This script widens the net: it draws random configurations over the
whole option space of the hot path (problem, Riemann solver, reconstruction, time integrator,
boundary types, gravity mode, well-balanced flux, conduction incl. its boundary modes, viscosity,
CFL, grid shape), runs oracle/_ref/fv2d_ref (the unmodified reference headers on Kokkos-OpenMP,
one thread) and the oracle on the same .ini, and demands BIT-IDENTICAL dt sequences and final
Q / U.  Nothing is written to tests/golden; the outcome of the last run is recorded in DESIGN.md §7.

    make -C oracle ref oracle && python tests/golden/fuzz_oracle_vs_reference.py [--n 300] [--seed 1]
"""
import argparse
import os
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
sys.path.insert(0, str(ROOT / "tests" / "golden"))
import oracle_lib as O  # noqa: E402
from fv2d_b200 import capi  # noqa: E402
from make_goldens import REF, SETTINGS, apply_overrides, read_dump  # noqa: E402

BASES = ["sod_x.ini", "sod_y.ini", "blast.ini", "rayleigh_taylor.ini", "diffusion.ini", "H84_chi1.5.ini", "C91.ini",
         "kelvin_helmholtz.ini", "gresho_vortex.ini"]


def draw(rng):
    base = BASES[rng.integers(len(BASES))]
    ov = {"mesh.Nx": int(rng.integers(4, 41)), "mesh.Ny": int(rng.integers(4, 41))}
    if rng.random() < 0.8:
        ov["solvers.riemann_solver"] = ["hll", "hllc", "fslp"][rng.integers(3)]
    if rng.random() < 0.8:
        ov["solvers.reconstruction"] = ["pcm", "pcm_wb", "plm"][rng.integers(3)]
    if rng.random() < 0.5:
        ov["solvers.time_stepping"] = ["euler", "RK2"][rng.integers(2)]
    if rng.random() < 0.5:
        ov["solvers.CFL"] = round(float(rng.uniform(0.05, 0.6)), 3)
    for ax in ("x", "y"):
        if rng.random() < 0.6:
            ov[f"run.boundaries_{ax}"] = ["absorbing", "reflecting", "periodic"][rng.integers(3)]
    if rng.random() < 0.5:
        mode = ["none", "constant", "analytical"][rng.integers(3)]
        ov["gravity.mode"] = mode
        ov["gravity.gx"] = round(float(rng.uniform(-1, 1)), 3)
        ov["gravity.gy"] = round(float(rng.uniform(-1, 1)), 3)
        if mode == "analytical":
            ov["hot_bubble.g0"] = round(float(rng.uniform(-1, 1)), 3)
    if rng.random() < 0.4:
        ov["physics.well_balanced_flux_at_y_bc"] = ["true", "false"][rng.integers(2)]
    if rng.random() < 0.4:
        ov["thermal_conduction.active"] = "true"
        ov["thermal_conduction.kappa"] = round(float(rng.uniform(0.0, 0.1)), 4)
        for side in ("ymin", "ymax"):
            ov[f"thermal_conduction.bc_{side}"] = ["none", "fixed_temperature", "fixed_gradient"][rng.integers(3)]
            ov[f"thermal_conduction.bc_{side}_value"] = round(float(rng.uniform(0.5, 5.0)), 3)
    if rng.random() < 0.4:
        ov["viscosity.active"] = "true"
        ov["viscosity.mu"] = round(float(rng.uniform(0.0, 0.05)), 4)
    # (drawn last, so that the configurations of earlier seeds keep everything else)
    if rng.random() < 0.25:
        ov["mesh.Nghosts"] = int(rng.integers(2, 5))
    return base, ov


def one(base, ov, steps, env):
    ini_text = apply_overrides((SETTINGS / base).read_text(), ov)
    with tempfile.TemporaryDirectory() as td:
        ini = Path(td) / "case.ini"
        ini.write_text(ini_text)
        dump = Path(td) / "dump.bin"
        out = subprocess.run([str(REF), str(ini), "--steps", str(steps), "--dump", str(dump)], env=env,
                             capture_output=True, text=True, cwd=td)
        if out.returncode != 0:
            return "ref-failed", out.stderr[-300:]
        d = read_dump(dump)
        try:
            dev, run = capi.params_from_ini(str(ini))
        except Exception as e:  # configuration the mirror rejects on purpose (Nghosts < 2, TCM_B02, ...)
            return "rejected", str(e)
        Q = capi.init_problem(dev, run)
        if not np.array_equal(O.domain(dev, Q), d["Q0"], equal_nan=True):
            return "MISMATCH init", ""
        U = O.prim_to_cons(dev, Q)
        n, t, dts, neg = O.run(dev, run.time_stepping, run.epsilon_reset_negative, run.tend, Q, U, steps)
        ok = (n == d["nsteps"] and np.array_equal(dts, d["dts"], equal_nan=True)
              and np.array_equal(O.domain(dev, Q), d["QN"], equal_nan=True)
              and np.array_equal(O.domain(dev, U), d["UN"], equal_nan=True))
        if not ok:
            return "MISMATCH run", ""
        return ("ok" if np.all(np.isfinite(d["UN"])) else "ok (non-finite state, identically so)"), ""


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=300)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    a = ap.parse_args()
    if not REF.exists():
        sys.exit(f"{REF} missing: run `make -C oracle ref` (needs /root/reference)")
    env = dict(os.environ, OMP_NUM_THREADS="1", OMP_PROC_BIND="false")
    rng = np.random.default_rng(a.seed)
    tally = {}
    for k in range(a.n):
        base, ov = draw(rng)
        res, msg = one(base, ov, a.steps, env)
        tally[res] = tally.get(res, 0) + 1
        if res.startswith("MISMATCH") or res == "ref-failed":
            print(f"[{k}] {res}: {base} {ov} {msg}", flush=True)
    print("fuzz_oracle_vs_reference:", a.n, "cases, seed", a.seed, "->", tally)
    return 1 if any(k.startswith("MISMATCH") for k in tally) else 0


if __name__ == "__main__":
    sys.exit(main())
