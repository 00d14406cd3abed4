"""Host logic: .ini reader / Params mirror (reference SimInfo.h:100-569, inih INIReader.h).
CPU-only; the expected values come from the reference's own printout stored in the goldens."""
import ctypes as C

import numpy as np
import pytest

from conftest import load_golden
from fv2d_b200 import capi


def test_params_match_reference_printout(golden, capfd):
    dev, run = capi.params_from_ini(golden.ini_path())
    err = capfd.readouterr().err
    ref = golden.ref_params()
    for k in ("dx", "dy", "gamma0", "CFL", "epsilon", "gx", "gy", "kappa", "mu"):
        assert getattr(dev, k) == ref[k], k  # bit-identical doubles
    assert run.tend == ref["tend"]
    assert (dev.Nx, dev.Ny) == (int(ref["Nx"]), int(ref["Ny"]))
    assert dev.Ntx == dev.Nx + 4 and dev.Nty == dev.Ny + 4 and dev.ibeg == 2 and dev.jend == dev.Ny + 2
    # checkValidityIni warnings are the reference's, line for line (SimInfo.h:501-527)
    ours = sorted(l for l in err.splitlines() if l.startswith("WARNING"))
    theirs = sorted(l for l in golden.warnings.splitlines() if l.startswith("WARNING"))
    assert ours == theirs


def test_real_values_are_float_rounded_Q1():
    g = load_golden("sod_x")
    dev, run = capi.params_from_ini(g.ini_path())
    assert dev.gamma0 == float(np.float32(1.666666667)) == 1.6666666269302368
    assert dev.CFL == float(np.float32(0.1)) == 0.10000000149011612
    assert run.save_freq == float(np.float32(0.01))
    assert dev.epsilon == float(np.float32(1.0e-6))  # defaults are narrowed too
    assert run.epsilon_reset_negative == float(np.float32(1.0e-8))


def test_misspelt_sections_fall_back_to_defaults_Q2():
    dev, _ = capi.params_from_ini(load_golden("kh_plm_128x64").ini_path())
    # uflow / z1 / z2 of the file are ignored: read from section "kelvin_helmholts"
    assert (dev.kh_uflow, dev.kh_y1, dev.kh_y2) == (1.0, 0.5, 1.5)
    assert dev.kh_P0 == 10.0 and dev.kh_rho_fac == 1.0
    dev, _ = capi.params_from_ini(load_golden("c91_64x32").ini_path())
    # bc_xmin/bc_xmax in the file, bc_ymin/bc_ymax in the reader -> no conduction BC
    assert dev.bctc_ymin == capi.BCTC_NONE and dev.bctc_ymax == capi.BCTC_NONE
    assert dev.reconstruction == capi.PCM_WB and dev.well_balanced_flux_at_y_bc == 1
    assert dev.thermal_conductivity_active == 1 and dev.viscosity_active == 1
    assert dev.kappa == float(np.float32(0.07)) and dev.mu == float(np.float32(0.0028))


def test_overrides_behave_like_file_lines(tmp_path):
    g = load_golden("blast_64")
    dev, run = capi.params_from_ini(g.ini_path(), {"mesh.Nx": 100, "solvers.CFL": 0.3, "solvers.time_stepping": "RK2",
                                                     "Run.Boundaries_X": "absorbing"})
    assert dev.Nx == 100 and dev.Ntx == 104 and dev.iend == 102
    assert dev.dx == (dev.xmax - dev.xmin) / 100
    assert dev.CFL == float(np.float32(0.3))
    assert run.time_stepping == capi.TS_RK2 and dev.boundary_x == capi.BC_ABSORBING


def _write(tmp_path, text):
    p = tmp_path / "t.ini"
    p.write_text(text)
    return p


def test_ini_syntax_quirks(tmp_path):
    text = ("﻿[MESH]\nNX = 0x20 ; hex and an inline comment\nny: 12\n"
            "# comment line\n; another\n[Solvers]\nCFL=0.25;not-a-comment\nriemann_solver = hll\n"
            "[run]\ntend=2\n  \n[physics]\nproblem = sod_x\nwell_balanced_flux_at_y_bc = YES\n")
    dev, run = capi.params_from_ini(_write(tmp_path, text))
    assert dev.Nx == 32 and dev.Ny == 12          # strtol(..., 0); ':' separator; case-insensitive names
    assert dev.CFL == float(np.float32(0.25))      # ';' without preceding blank is part of the value; strtof stops at it
    assert dev.riemann_solver == capi.HLL and dev.well_balanced_flux_at_y_bc == 1
    assert run.tend == 2.0 and run.problem == b"sod_x"
    # defaults (SimInfo.h:360-458): reflecting, pcm, hllc default only when the key is absent
    assert dev.boundary_x == capi.BC_REFLECTING and dev.reconstruction == capi.PCM and dev.Ng == 2


def test_multiline_value_and_duplicate_key(tmp_path):
    # a repeated key appends "\n<value>" (INIReader.h:450-459); strtof then reads the first number
    dev, _ = capi.params_from_ini(_write(tmp_path, "[solvers]\nCFL=0.5\nCFL=0.7\n[physics]\nproblem=blast\n"))
    assert dev.CFL == 0.5
    # ... and makes an enum string invalid -> runtime_error in the reference, error code here
    with pytest.raises(capi.Fv2dError) as e:
        capi.params_from_ini(_write(tmp_path, "[solvers]\nriemann_solver=hll\nriemann_solver=hllc\n"))
    assert e.value.code == 4 and "bad parameter for riemann_solver" in str(e.value)


def test_error_conventions(tmp_path):
    with pytest.raises(capi.Fv2dError) as e:
        capi.params_from_ini(tmp_path / "missing.ini")
    assert e.value.code == 3
    with pytest.raises(capi.Fv2dError) as e:  # bad enum string (SimInfo.h:214-220)
        capi.params_from_ini(_write(tmp_path, "[run]\nboundaries_x=open\n"))
    assert e.value.code == 4 and "allowed values" in str(e.value)
    dev, run = capi.params_from_ini(_write(tmp_path, "[physics]\nproblem=warp_drive\n[mesh]\nNx=4\nNy=4\n"))
    with pytest.raises(capi.Fv2dError) as e:  # unknown problem (Init.h:303-304)
        capi.init_problem(dev, run)
    assert "unknown problem warp_drive" in str(e.value)


def test_effective_config_dump(tmp_path):
    g = load_golden("rt_plm_32x96")
    out = tmp_path / "last.ini"
    capi.params_dump_ini(g.ini_path(), out)
    text = out.read_text()
    assert text.startswith("; Parameters used for the problem: rayleigh-taylor")
    assert "[gravity]" in text and "[thermal_conduction]" not in text  # sections absent from the file are skipped
    line = [l for l in text.splitlines() if l.startswith("gy ")][0]
    assert "-1.000000014901e-01" in line and "default" not in line       # float-rounded, 12 digits, from file
    assert any(l.rstrip().endswith("; default") for l in text.splitlines() if l.startswith("fslp_k") or l.startswith("gx "))
    # the dump is itself a valid .ini that reproduces the same parameters
    dev1, _ = capi.params_from_ini(g.ini_path())
    dev2, _ = capi.params_from_ini(out)
    assert bytes(dev1) == bytes(dev2)
