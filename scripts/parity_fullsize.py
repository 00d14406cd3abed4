#!/usr/bin/env python
"""Full-size parity of the fused CUDA path against the REFERENCE BINARY on the GPU box.

BASELINE.json's correctness bar — "checked against the reference's own Kokkos-OpenMP build on
identical .ini inputs: relative L1 <= 1e-12 on the conserved fields after 10 steps, dt sequence
<= 1e-13 relative, domain-integrated mass / energy drift equal to the reference" — evaluated
DIRECTLY at the sizes BASELINE.json names (blast 4096^2, Kelvin-Helmholtz 8192^2 PLM, C91 8192^2,
Rayleigh-Taylor 16384^2), not through the small fixtures:

  1. the host Init (bit-identical to the reference's, tests/test_init_and_oracle.py) builds Q0;
  2. oracle/_ref/fv2d_ref (the unmodified reference headers + Kokkos-OpenMP, all host cores) loads
     that Q0 (--load-q0: sidesteps the thread-count dependent RNG of C91), runs 10
     steps and dumps U_N + the dt sequence (--dump-lean);
  3. the fused path (fv2d_run_steps, dt resident on the device) runs the same 10 steps from the
     same Q0 on cuda:0;
  4. the two are compared.

The reference binary is test infrastructure (oracle/); nothing here is timed as a product number.

    python scripts/parity_fullsize.py [--cases blast_4096,kelvin_helmholtz_8192_plm,...] [--steps 10]
                                      [--out gpurun_out/parity_fullsize.json]
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import struct
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests" / "golden"))

REF = ROOT / "oracle" / "_ref" / "fv2d_ref"

BAR_L1, BAR_DT, BAR_SUM = 1e-12, 1e-13, 1e-13

CASES = {
    # name: (settings file, overrides)                                    BASELINE.json config
    "blast_4096": ("blast.ini", {"mesh.Nx": 4096, "mesh.Ny": 4096}),                               # C2
    "kelvin_helmholtz_8192_plm": ("kelvin_helmholtz.ini", {"mesh.Nx": 8192, "mesh.Ny": 8192,
                                                           "solvers.reconstruction": "plm"}),     # C3
    "rayleigh_taylor_16384": ("rayleigh_taylor.ini", {"mesh.Nx": 16384, "mesh.Ny": 16384}),        # C4
    "c91_8192": ("C91.ini", {"mesh.Nx": 8192, "mesh.Ny": 8192}),                                   # C5
}


def mem_available_gb() -> float:
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable:"):
                return int(line.split()[1]) / 1e6
    except OSError:
        pass
    return 0.0


def needs(name: str, scale=None):
    """(host RAM GB, scratch disk GB) one case needs: the reference holds 4 arrays + the dumped
    U_N; this process holds Q0 (+ its domain copy while writing), the reference's U_N and ours."""
    ini, ov = CASES[name]
    nx, ny = (scale or (ov["mesh.Nx"], ov["mesh.Ny"]))
    arr = 32.0 * (nx + 4) * (ny + 4) / 1e9
    return 9.0 * arr + 2.0, 2.2 * arr + 0.5


def read_lean(path):
    with open(path, "rb") as f:
        head = f.read(24)
        assert head[:8] == b"FV2DLEAN", head[:8]
        nx, ny, nsteps, nf = struct.unpack("<4i", head[8:24])
        (t,) = struct.unpack("<d", f.read(8))
        dts = np.frombuffer(f.read(8 * nsteps), "<f8").copy()
        UN = np.fromfile(f, "<f8", nf * ny * nx).reshape(nf, ny, nx)
        mass, energy = struct.unpack("<2d", f.read(16))
    return dict(Nx=nx, Ny=ny, nsteps=nsteps, t=t, dts=dts, UN=UN, mass=mass, energy=energy)


def run_case(name: str, steps: int = 10, scale=None, device: int = 0, threads: int | None = None) -> dict:
    """Runs one configuration through the reference binary and through the fused CUDA path and
    returns the comparison record.  `scale` = (Nx, Ny) overrides the size (development only)."""
    from make_goldens import apply_overrides

    from fv2d_b200 import capi

    ini_name, ov = CASES[name]
    ov = dict(ov)
    if scale:
        ov["mesh.Nx"], ov["mesh.Ny"] = scale
    rec = {"case": name, "steps": steps}
    with tempfile.TemporaryDirectory(prefix="fv2d_parity_") as td:
        td = Path(td)
        ini = td / f"{name}.ini"
        ini.write_text(apply_overrides((ROOT / "settings" / ini_name).read_text(), ov))
        dev, run = capi.params_from_ini(ini)
        rec.update(Nx=dev.Nx, Ny=dev.Ny)
        t0 = time.perf_counter()
        Q0 = capi.init_problem(dev, run)
        rec["host_init_s"] = round(time.perf_counter() - t0, 2)
        J, I = slice(dev.jbeg, dev.jend), slice(dev.ibeg, dev.iend)
        with open(td / "q0.bin", "wb") as f:
            for fld in range(4):
                np.ascontiguousarray(Q0[fld, J, I]).tofile(f)

        # ---- the reference itself
        ncores = threads or os.cpu_count() or 1
        env = dict(os.environ, OMP_NUM_THREADS=str(ncores), OMP_PROC_BIND="spread", OMP_PLACES="threads")
        t0 = time.perf_counter()
        out = subprocess.run([str(REF), str(ini), "--steps", str(steps), "--load-q0", str(td / "q0.bin"),
                              "--dump-lean", str(td / "ref.bin"), "--quiet"], env=env, capture_output=True, text=True,
                             cwd=td)
        rec["reference_s"] = round(time.perf_counter() - t0, 1)
        rec["reference_threads"] = ncores
        if out.returncode != 0:
            raise RuntimeError(f"{name}: fv2d_ref failed ({out.returncode})\n{out.stdout[-2000:]}\n{out.stderr[-2000:]}")
        (td / "q0.bin").unlink()
        ref = read_lean(td / "ref.bin")
        (td / "ref.bin").unlink()
        assert ref["nsteps"] == steps and ref["Nx"] == dev.Nx and ref["Ny"] == dev.Ny

        # ---- the fused CUDA path on the same Q0
        with capi.Context(dev, run.time_stepping, run.epsilon_reset_negative, device=device) as ctx:
            ctx.upload_Q(Q0)
            del Q0
            ctx.prim_to_cons()
            U0 = ctx.download_U()
            m0 = [float(np.sum(U0[f, J, I])) * dev.dx * dev.dy for f in (0, 3)]
            del U0
            ctx.compute_dt()
            ctx.run_steps(steps)
            dts = ctx.dt_history(steps)
            U = np.ascontiguousarray(ctx.download_U()[:, J, I])
            rec["negatives"] = ctx.negative_counts()

    Ur = ref["UN"]
    l1 = []
    scale_all = float(sum(np.sum(np.abs(Ur[f])) for f in range(4)))
    for f in range(4):
        num = float(np.sum(np.abs(U[f] - Ur[f])))
        den = float(np.sum(np.abs(Ur[f])))
        l1.append({"field": ("rho", "rho_u", "rho_v", "E")[f], "rel_l1": num / den if den > 0 else num,
                   "rel_l1_vs_state": num / scale_all})
    rec["rel_l1"] = l1
    rec["max_abs_diff"] = float(max(np.max(np.abs(U[f] - Ur[f])) for f in range(4)))
    rec["dt_max_rel_err"] = float(np.max(np.abs(dts - ref["dts"]) / ref["dts"]))
    rec["dt_first"], rec["dt_last"] = float(dts[0]), float(dts[-1])
    cell = dev.dx * dev.dy
    mine = [float(np.sum(U[f])) * cell for f in (0, 3)]
    theirs = [float(np.sum(Ur[f])) * cell for f in (0, 3)]
    rec["mass"] = {"initial": m0[0], "fused": mine[0], "reference": theirs[0], "drift_fused": mine[0] - m0[0],
                   "drift_reference": theirs[0] - m0[0], "rel_diff": abs(mine[0] - theirs[0]) / abs(theirs[0])}
    rec["energy"] = {"initial": m0[1], "fused": mine[1], "reference": theirs[1], "drift_fused": mine[1] - m0[1],
                     "drift_reference": theirs[1] - m0[1], "rel_diff": abs(mine[1] - theirs[1]) / abs(theirs[1])}
    # rho and E field by field; a momentum component that is ~0 everywhere (rho*u in
    # Rayleigh-Taylor) has no scale of its own and is held on the scale of the state vector
    ok_l1 = (l1[0]["rel_l1"] <= BAR_L1 and l1[3]["rel_l1"] <= BAR_L1 and
             all(min(l1[f]["rel_l1"], l1[f]["rel_l1_vs_state"]) <= BAR_L1 for f in (1, 2)))
    rec["pass"] = bool(ok_l1 and rec["dt_max_rel_err"] <= BAR_DT and rec["mass"]["rel_diff"] <= BAR_SUM and
                       rec["energy"]["rel_diff"] <= BAR_SUM and rec["negatives"] == [0, 0, 0])
    rec["bars"] = {"rel_l1": BAR_L1, "dt": BAR_DT, "mass_energy": BAR_SUM}
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default=",".join(CASES))
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--out", default=str(ROOT / "gpurun_out" / "parity_fullsize.json"))
    ap.add_argument("--scale", default="", help="NxxNy: run every case at this size instead (development)")
    args = ap.parse_args()
    if not REF.exists():
        sys.exit(f"{REF} missing: build it in the dev container with `make -C oracle ref`")
    scale = tuple(int(v) for v in args.scale.split("x")) if args.scale else None
    records = []
    for name in args.cases.split(","):
        ram, disk = needs(name, scale)
        free_ram, free_disk = mem_available_gb(), shutil.disk_usage(tempfile.gettempdir()).free / 1e9
        if free_ram < ram or free_disk < disk:
            records.append({"case": name, "skipped": f"needs {ram:.0f} GB RAM / {disk:.0f} GB scratch, box has "
                                                     f"{free_ram:.0f} / {free_disk:.0f}"})
            print(json.dumps(records[-1]), flush=True)
            continue
        rec = run_case(name, args.steps, scale)
        records.append(rec)
        print(json.dumps(rec), flush=True)
    Path(args.out).parent.mkdir(parents=True, exist_ok=True)
    Path(args.out).write_text(json.dumps(records, indent=1) + "\n")
    bad = [r["case"] for r in records if not r.get("pass", True)]
    if bad:
        sys.exit(f"parity FAILED for {bad}")


if __name__ == "__main__":
    main()
