#!/usr/bin/env python
"""Regenerates tests/golden/*.npz by RUNNING THE REFERENCE ITSELF.

Each fixture is produced by oracle/_ref/fv2d_ref — the unmodified reference headers
(mdelorme/fv2d @680ff34) compiled against the vendored Kokkos 4.1.00 OpenMP backend by
oracle/Makefile — on an .ini derived from settings/<name>.ini with a few keys replaced
(sizes scaled down so fixtures stay small).  The exact .ini text is stored inside the
fixture so the parity tests re-create the same input anywhere (the GPU box has neither
/root/reference nor needs fv2d_ref).

Run in the dev container:   make -C oracle ref && python tests/golden/make_goldens.py
OMP_NUM_THREADS=1 is used so that the C91/H84 initial perturbation (Kokkos random pool,
thread-count dependent: SURVEY.md Q11) is the single-stream one the host Init reproduces.
"""
import os
import re
import struct
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
REF = ROOT / "oracle" / "_ref" / "fv2d_ref"
SETTINGS = ROOT / "settings"
OUT = Path(__file__).resolve().parent

# name -> (base ini, {section.key: value}, steps)
CASES = {
    "sod_x": ("sod_x.ini", {}, 10),
    "sod_y": ("sod_y.ini", {}, 10),
    "blast_64": ("blast.ini", {"mesh.Nx": 64, "mesh.Ny": 64}, 10),
    "kh_plm_128x64": ("kelvin_helmholtz.ini", {"mesh.Nx": 128, "mesh.Ny": 64, "solvers.reconstruction": "plm"}, 10),
    "kh_pcm_hll_64x32": ("kelvin_helmholtz.ini", {"mesh.Nx": 64, "mesh.Ny": 32, "solvers.riemann_solver": "hll"}, 10),
    "kh_plm_hll_64x32": ("kelvin_helmholtz.ini", {"mesh.Nx": 64, "mesh.Ny": 32, "solvers.riemann_solver": "hll",
                                                 "solvers.reconstruction": "plm"}, 10),
    "rt_plm_32x96": ("rayleigh_taylor.ini", {"mesh.Nx": 32, "mesh.Ny": 96}, 10),
    "rt_fslp_32x96": ("rayleigh_taylor.ini", {"mesh.Nx": 32, "mesh.Ny": 96, "solvers.riemann_solver": "fslp"}, 10),
    "c91_64x32": ("C91.ini", {"mesh.Nx": 64, "mesh.Ny": 32}, 10),
    "c91_bctc_64x32": ("C91.ini", {"mesh.Nx": 64, "mesh.Ny": 32, "thermal_conduction.bc_ymin": "fixed_temperature",
                                   "thermal_conduction.bc_ymax": "fixed_gradient",
                                   "thermal_conduction.bc_ymin_value": 1.0,
                                   "thermal_conduction.bc_ymax_value": 10.0}, 10),
    "h84_80x20": ("H84_chi1.5.ini", {"mesh.Nx": 80, "mesh.Ny": 20}, 10),
    "diffusion_48": ("diffusion.ini", {"mesh.Nx": 48, "mesh.Ny": 48}, 10),
    "gresho_rk2_32": ("gresho_vortex.ini", {"mesh.Nx": 32, "mesh.Ny": 32}, 10),
    "blast_rk2_plm_48": ("blast.ini", {"mesh.Nx": 48, "mesh.Ny": 48, "solvers.reconstruction": "plm",
                                       "solvers.time_stepping": "RK2", "run.boundaries_x": "absorbing",
                                       "run.boundaries_y": "reflecting"}, 10),
    "sod_x_300wide": ("sod_x.ini", {"mesh.Nx": 300, "mesh.Ny": 8, "solvers.reconstruction": "plm"}, 10),
    # well-balanced y-boundary flux (Update.h:148-156) outside its C91 habitat: with PLM + gravity, and
    # with NO gravity at all (the reference applies it whatever the gravity mode, g = 0)
    "rt_wb_plm_32x96": ("rayleigh_taylor.ini", {"mesh.Nx": 32, "mesh.Ny": 96,
                                                "physics.well_balanced_flux_at_y_bc": "true"}, 10),
    "blast_wb_nograv_48": ("blast.ini", {"mesh.Nx": 48, "mesh.Ny": 48, "run.boundaries_y": "reflecting",
                                         "physics.well_balanced_flux_at_y_bc": "true"}, 10),
    "c91_wb_nograv_48x24": ("C91.ini", {"mesh.Nx": 48, "mesh.Ny": 24, "gravity.mode": "none"}, 10),
    # conduction + viscosity on top of PLM slopes, and with the HLL solver
    "c91_plm_64x32": ("C91.ini", {"mesh.Nx": 64, "mesh.Ny": 32, "solvers.reconstruction": "plm"}, 10),
    "c91_hll_48x24": ("C91.ini", {"mesh.Nx": 48, "mesh.Ny": 24, "solvers.riemann_solver": "hll"}, 10),
    "rt_fslp_pcm_32x96": ("rayleigh_taylor.ini", {"mesh.Nx": 32, "mesh.Ny": 96, "solvers.riemann_solver": "fslp",
                                                  "solvers.reconstruction": "pcm"}, 10),
    "c91_fslp_plm_48x24": ("C91.ini", {"mesh.Nx": 48, "mesh.Ny": 24, "solvers.riemann_solver": "fslp",
                                       "solvers.reconstruction": "plm"}, 10),
}


def apply_overrides(text: str, overrides: dict) -> str:
    """Replace/insert `key=value` inside `[section]` of an .ini text."""
    lines = text.splitlines()
    for sk, val in overrides.items():
        section, key = sk.split(".", 1)
        sec_start, sec_end, done = None, len(lines), False
        for n, line in enumerate(lines):
            s = line.strip()
            if s.startswith("["):
                if sec_start is not None and sec_end == len(lines):
                    sec_end = n
                if s[1:s.index("]")].lower() == section.lower() and sec_start is None:
                    sec_start = n
        if sec_start is None:
            lines += ["", f"[{section}]", f"{key}={val}"]
            continue
        for n in range(sec_start + 1, sec_end):
            m = re.match(r"\s*([^=:;#\s]+)\s*[=:]", lines[n])
            if m and m.group(1).lower() == key.lower():
                lines[n] = f"{key}={val}"
                done = True
        if not done:
            lines.insert(sec_start + 1, f"{key}={val}")
    return "\n".join(lines) + "\n"


def read_dump(path):
    raw = Path(path).read_bytes()
    assert raw[:8] == b"FV2DDUMP"
    nx, ny, nsteps, nf = struct.unpack_from("<4i", raw, 8)
    off = 24
    (t,) = struct.unpack_from("<d", raw, off)
    off += 8
    dts = np.frombuffer(raw, "<f8", nsteps, off)
    off += 8 * nsteps
    n = nf * ny * nx
    Q0 = np.frombuffer(raw, "<f8", n, off).reshape(nf, ny, nx)
    off += 8 * n
    QN = np.frombuffer(raw, "<f8", n, off).reshape(nf, ny, nx)
    off += 8 * n
    UN = np.frombuffer(raw, "<f8", n, off).reshape(nf, ny, nx)
    off += 8 * n
    mass, energy = struct.unpack_from("<2d", raw, off)
    return dict(Nx=nx, Ny=ny, nsteps=nsteps, t=t, dts=dts.copy(), Q0=Q0.copy(), QN=QN.copy(), UN=UN.copy(), mass=mass,
                energy=energy)


def main(names):
    if not REF.exists():
        sys.exit(f"{REF} missing: run `make -C oracle ref` (needs /root/reference)")
    env = dict(os.environ, OMP_NUM_THREADS="1", OMP_PROC_BIND="false")
    for name in names:
        base, ov, steps = CASES[name]
        ini_text = apply_overrides((SETTINGS / base).read_text(), ov)
        with tempfile.TemporaryDirectory() as td:
            ini = Path(td) / f"{name}.ini"
            ini.write_text(ini_text)
            dump = Path(td) / "dump.bin"
            out = subprocess.run([str(REF), str(ini), "--steps", str(steps), "--dump", str(dump)], env=env,
                                 capture_output=True, text=True, cwd=td)
            if out.returncode != 0:
                sys.exit(f"{name}: fv2d_ref failed\n{out.stdout}\n{out.stderr}")
            d = read_dump(dump)
        params_line = [l for l in out.stdout.splitlines() if l.startswith("params:")][0]
        np.savez_compressed(OUT / f"{name}.npz", ini_text=np.array(ini_text), base=np.array(base),
                            params_line=np.array(params_line), warnings=np.array(out.stderr), **d)
        print(f"{name}: {d['Nx']}x{d['Ny']} steps={d['nsteps']} dt0={d['dts'][0]:.17g} mass={d['mass']:.17g} "
              f"-> {(OUT / (name + '.npz')).stat().st_size} B")


if __name__ == "__main__":
    main(sys.argv[1:] or list(CASES))
