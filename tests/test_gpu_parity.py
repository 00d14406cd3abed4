"""Parity of the CUDA path (through the C ABI) with the oracle / the reference's goldens.

Tolerances are BASELINE.json's: conserved fields relative L1 <= 1e-12 after 10 steps, dt
sequence <= 1e-13 relative.  The operator-level path is held to a stricter bar: it must be
BIT-IDENTICAL to the reference (it is compiled without FMA contraction and uses IEEE
division / square root in the reference's order of operations).
"""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

import oracle_lib as O
from conftest import load_golden, rel_l1
from fv2d_b200 import capi

pytestmark = pytest.mark.gpu

TOL_L1 = 1e-12   # BASELINE.json: relative L1 of conserved fields after 10 steps
TOL_DT = 1e-13   # BASELINE.json: dt sequence, relative
ROOT = Path(__file__).resolve().parents[1]


def _setup(golden):
    dev, run = capi.params_from_ini(golden.ini_path())
    return dev, run, capi.init_problem(dev, run)


def test_operator_path_is_bit_identical_to_reference(golden):
    """main.cpp:62-84 driven operator by operator through the C ABI."""
    dev, run, Q0 = _setup(golden)
    with capi.Context(dev, run.time_stepping, run.epsilon_reset_negative) as ctx:
        ctx.upload_Q(Q0)
        ctx.prim_to_cons()
        dts = []
        for _ in range(golden.nsteps):
            dt, _ = ctx.compute_dt()
            dts.append(dt)
            ctx.update(dt)
            ctx.cons_to_prim()
            assert ctx.check_negatives() == [0, 0, 0]
        Q, U = ctx.download_Q(), ctx.download_U()
    assert np.array_equal(np.array(dts), golden.dts)
    assert np.array_equal(O.domain(dev, U), golden.UN)
    assert np.array_equal(O.domain(dev, Q), golden.QN)


def test_fused_path_matches_reference(golden):
    """The hot path: fused sweep kernel per RK stage, dt resident on the device."""
    dev, run, Q0 = _setup(golden)
    with capi.Context(dev, run.time_stepping, run.epsilon_reset_negative) as ctx:
        ctx.upload_Q(Q0)
        ctx.prim_to_cons()
        ctx.compute_dt()
        ctx.run_steps(golden.nsteps)
        Q, U = ctx.download_Q(), ctx.download_U()
        dts = ctx.dt_history(golden.nsteps)
        t, _, steps = ctx.get_time()
        mass, energy = ctx.mass_energy()
        assert ctx.negative_counts() == [0, 0, 0]
    assert steps == golden.nsteps
    assert np.max(np.abs(dts - golden.dts) / golden.dts) <= TOL_DT
    assert abs(t - golden.t) <= TOL_DT * golden.t
    Ud, Qd = O.domain(dev, U), O.domain(dev, Q)
    assert rel_l1(Ud, golden.UN) <= TOL_L1 and rel_l1(Qd, golden.QN) <= TOL_L1
    for f in range(4):  # per field too (momenta can be tiny: bound by the field's own scale or rho's)
        scale = max(np.sum(np.abs(golden.UN[f])), 1e-3 * np.sum(np.abs(golden.UN[0])))
        assert np.sum(np.abs(Ud[f] - golden.UN[f])) / scale <= TOL_L1, f
    # domain-integrated mass / energy drift equal to the reference's
    assert abs(mass - golden.mass) <= TOL_L1 * abs(golden.mass)
    assert abs(energy - golden.energy) <= TOL_L1 * abs(golden.energy)


def test_fused_step_with_host_dt_equals_device_dt():
    g = load_golden("kh_plm_128x64")
    dev, run, Q0 = _setup(g)
    out = []
    for mode in ("device", "host"):
        with capi.Context(dev, run.time_stepping, run.epsilon_reset_negative) as ctx:
            ctx.upload_Q(Q0)
            ctx.prim_to_cons()
            ctx.compute_dt()
            if mode == "device":
                ctx.run_steps(5)
            else:
                for k in range(5):
                    _, dt, _ = ctx.get_time()
                    ctx.step(dt)
            out.append((ctx.download_U(), ctx.dt_history(5)))
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])


@pytest.mark.parametrize("name", ["rt_plm_32x96", "c91_bctc_64x32", "kh_plm_hll_64x32", "rt_fslp_32x96"])
def test_individual_operators_bit_identical_to_oracle(name):
    g = load_golden(name)
    dev, run, Q0 = _setup(g)
    dt = float(g.dts[0])
    with capi.Context(dev, run.time_stepping, run.epsilon_reset_negative) as ctx:
        # primToCons / consToPrim over range_tot
        ctx.upload_Q(Q0)
        ctx.prim_to_cons()
        U = ctx.download_U()
        assert np.array_equal(U, O.prim_to_cons(dev, Q0))
        ctx.cons_to_prim()
        assert np.array_equal(ctx.download_Q(), O.cons_to_prim(dev, U))
        # fillBoundaries on scrambled ghosts
        Qs = Q0.copy()
        rng = np.random.default_rng(1)
        mask = np.ones(Qs.shape[1:], bool)
        mask[dev.jbeg:dev.jend, dev.ibeg:dev.iend] = False
        Qs[:, mask] = rng.normal(size=(4, int(mask.sum())))
        ctx.upload_Q(Qs)
        ctx.fill_boundaries()
        Qo = Qs.copy()
        O.fill_boundaries(dev, Qo)
        assert np.array_equal(ctx.download_Q(), Qo)
        # computeDt
        dtg, inv = ctx.compute_dt()
        dto, invo = O.compute_dt(dev, Qo)
        assert dtg == dto and list(inv) == list(invo)
        # slopes + fluxes + source terms, one by one, on the same Q
        ctx.upload_U(U)
        Uo = U.copy()
        sx, sy = np.zeros_like(U), np.zeros_like(U)
        if dev.reconstruction == capi.PLM:
            ctx.compute_slopes()
            sx, sy = O.compute_slopes(dev, Qo)
        ctx.compute_fluxes_and_update(dt)
        O.compute_fluxes_and_update(dev, Qo, sx, sy, Uo, dt)
        assert np.array_equal(ctx.download_U(), Uo)
        if dev.thermal_conductivity_active:
            ctx.apply_thermal_conduction(dt)
            O.apply_thermal_conduction(dev, Qo, Uo, dt)
            assert np.array_equal(ctx.download_U(), Uo)
        if dev.viscosity_active:
            ctx.apply_viscosity(dt)
            O.apply_viscosity(dev, Qo, Uo, dt)
            assert np.array_equal(ctx.download_U(), Uo)


def test_check_negatives_on_device():
    g = load_golden("sod_x")
    dev, run, Q0 = _setup(g)
    Q = Q0.copy()
    Q[0, dev.jbeg + 1, dev.ibeg + 3] = -1.0
    Q[3, dev.jbeg + 2, dev.ibeg + 5] = -2.0
    Q[3, dev.jbeg + 2, dev.ibeg + 6] = -2.5
    Q[1, dev.jbeg, dev.ibeg] = np.nan
    Q[0, 0, 0] = -5.0  # ghost cell: not in range_dom
    Qo = Q.copy()
    want = O.check_negatives(dev, run.epsilon_reset_negative, Qo)
    with capi.Context(dev, run.time_stepping, run.epsilon_reset_negative) as ctx:
        ctx.upload_Q(Q)
        got = ctx.check_negatives()
        Qg = ctx.download_Q()
    assert got == want == [1, 2, 1]
    assert np.array_equal(Qg, Qo, equal_nan=True)


def test_fused_negative_reset_and_counters():
    """Steps that drive cells negative (violent expansion, dt 12x the CFL limit): the fused
    epilogue must reset Q (not U) and count exactly like consToPrim + checkNegatives
    (SimInfo.h:602-646)."""
    g = load_golden("blast_64")
    dev, run, Q0 = _setup(g)
    Q0 = Q0.copy()
    mid = dev.ibeg + dev.Nx // 2
    Q0[1, :, mid:], Q0[1, :, :mid], Q0[3] = 4.0, -4.0, 0.05
    O.fill_boundaries(dev, Q0)
    eps = run.epsilon_reset_negative
    dt = 12.0 * O.compute_dt(dev, Q0)[0]
    Qo, Uo, neg_o = Q0.copy(), O.prim_to_cons(dev, Q0), [0, 0, 0]
    for _ in range(3):
        O.update(dev, run.time_stepping, Qo, Uo, dt)
        Qo = O.cons_to_prim(dev, Uo)
        neg_o = [x + y for x, y in zip(neg_o, O.check_negatives(dev, eps, Qo))]
    with capi.Context(dev, run.time_stepping, eps) as ctx:
        ctx.upload_Q(Q0)
        ctx.prim_to_cons()
        for _ in range(3):
            ctx.step(dt)
        neg = ctx.negative_counts()
        Q, U = ctx.download_Q(), ctx.download_U()
    assert neg_o[0] > 0 and neg_o[1] > 0, "test setup should trigger resets"
    assert neg == neg_o
    assert np.array_equal(O.domain(dev, Q) == eps, O.domain(dev, Qo) == eps)
    assert rel_l1(O.domain(dev, U), O.domain(dev, Uo)) <= 1e-11  # U is not reset (SimInfo.h:614-627)


def test_run_until_replays_the_reference_loop_condition():
    g = load_golden("sod_x")
    dev, run, Q0 = _setup(g)
    tend = float(np.sum(g.dts[:4])) + 0.25 * float(g.dts[4])  # inside step 5 -> 5 steps are taken (Q6: no clipping)
    with capi.Context(dev, run.time_stepping, run.epsilon_reset_negative) as ctx:
        ctx.upload_Q(Q0)
        ctx.prim_to_cons()
        ctx.compute_dt()
        n = ctx.run_until(tend, 1000)
        t, _, steps = ctx.get_time()
    assert n == steps == 5 and t > tend
    assert abs(t - float(np.sum(g.dts[:5]))) <= 1e-15


def test_advance_host_equals_resident_run():
    g = load_golden("rt_plm_32x96")
    dev, run, Q0 = _setup(g)
    with capi.Context(dev, run.time_stepping, run.epsilon_reset_negative) as ctx:
        Qout = np.empty_like(Q0)
        dts = np.zeros(g.nsteps)
        ctx.advance_host(Q0, Qout, g.nsteps, dts)
    assert np.max(np.abs(dts - g.dts) / g.dts) <= TOL_DT
    assert rel_l1(O.domain(dev, Qout), g.QN) <= TOL_L1


@pytest.mark.parametrize("name", ["kh_plm_128x64", "c91_plm_64x32", "blast_rk2_plm_48"])
def test_cpp_host_driver_on_y_slabs_is_bitwise_the_single_gpu_run(tmp_path, name):
    """fv2d_b200_main <ini> --gpus N: the C++17 driver on N y-slabs (SlabSet in host/Operators.h, one
    process, one context per slab; the slabs share devices when the box has fewer GPUs than slabs).
    Same log lines, and the snapshot files are byte for byte the single-GPU run's."""
    exe = ROOT / "fv2d_b200" / "fv2d_b200_main"
    if not exe.exists():
        subprocess.run(["make", "driver"], cwd=ROOT, check=True, capture_output=True)
    g = load_golden(name)
    ndev = capi.device_count()
    runs = {}
    for n in (1, 2, 3):
        d = tmp_path / f"n{n}"
        d.mkdir()
        cmd = [str(exe), g.ini_path(), "--max-steps", "10"]
        if n > 1:
            cmd += ["--gpus", str(n)] + (["--share-devices"] if ndev < n else [])
        r = subprocess.run(cmd, cwd=d, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        files = {p.name: p.read_bytes() for p in sorted(d.iterdir()) if p.suffix in (".h5", ".xmf")}
        assert any(k.endswith(".h5") for k in files)
        runs[n] = ([l for l in r.stdout.splitlines() if l.startswith(("Computing dts", " - Saving", "Time at end", "-->"))], files)
    for n in (2, 3):
        assert runs[n][0] == runs[1][0], (n, runs[n][0], runs[1][0])
        # run.h5 / run.xmf (or the per-snapshot files of run.multiple_outputs): the same bytes
        assert runs[n][1].keys() == runs[1][1].keys()
        for k in runs[1][1]:
            assert runs[n][1][k] == runs[1][1][k], (n, k)


def test_cpp_host_driver_runs_the_reference_loop(tmp_path):
    """fv2d_b200_main <ini>: the C++17 mirror of main.cpp, fused and --unfused."""
    exe = ROOT / "fv2d_b200" / "fv2d_b200_main"
    if not exe.exists():
        subprocess.run(["make", "driver"], cwd=ROOT, check=True, capture_output=True)
    g = load_golden("sod_x")
    outs = {}
    for mode in ([], ["--unfused"]):
        r = subprocess.run([str(exe), g.ini_path(), "--max-steps", "10"] + mode, cwd=tmp_path, capture_output=True,
                           text=True)
        assert r.returncode == 0, r.stderr
        assert "Computing dts at (t=0) : dt_hyp=0.00968246" in r.stdout
        outs[bool(mode)] = r.stdout
        import h5mini

        f = h5mini.File(tmp_path / "run.h5")
        groups = [k for k in f.keys() if k.startswith("ite_")]
        assert groups == ["ite_0000", "ite_0001"] and f.keys()[-2:] == ["x", "y"]
        assert int(f.attrs["Nx"]) == 64 and int(f.attrs["Ny"]) == 16 and f.attrs["problem"] == "sod_x"
        last = f["ite_0001"]
        t = float(last.attrs["time"])
        assert abs(t - g.t) <= 1e-13 * g.t and int(last.attrs["iteration"]) == 1
        rho = last["rho"].read().reshape(16, 64)
        assert rel_l1(rho, g.QN[0]) <= TOL_L1
        assert np.array_equal(f["ite_0000/rho"].read().reshape(16, 64), g.Q0[0])
        assert "ite_0001/prs" in (tmp_path / "run.xmf").read_text()
        (tmp_path / "run.h5").unlink()
    assert (tmp_path / "last.ini").read_text().startswith("; Parameters used for the problem: sod_x")


def test_cpp_host_driver_restart_from_run_h5(tmp_path):
    """Restart (main.cpp:47-55, IOManager.h:284-398): 10 steps, snapshot, restart for 10 more ==
    20 steps in one go (the restart file stores Q only; U is recomputed, main.cpp:58)."""
    import h5mini

    exe = ROOT / "fv2d_b200" / "fv2d_b200_main"
    if not exe.exists():
        subprocess.run(["make", "driver"], cwd=ROOT, check=True, capture_output=True)
    g = load_golden("sod_x")
    ini = Path(g.ini_path()).read_text().replace("save_freq=0.01", "save_freq=1.0")  # no intermediate snapshot
    assert "save_freq=1.0" in ini
    a, b = tmp_path / "a", tmp_path / "b"
    a.mkdir(), b.mkdir()
    ini1 = tmp_path / "sod.ini"
    ini1.write_text(ini)
    r = subprocess.run([str(exe), str(ini1), "--max-steps", "20", "--quiet"], cwd=a, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe), str(ini1), "--max-steps", "10", "--quiet"], cwd=b, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    (b / "first.h5").write_bytes((b / "run.h5").read_bytes())
    ini2 = b / "restart.ini"
    ini2.write_text(ini.replace("[run]", "[run]\nrestart_file=first.h5"))
    r = subprocess.run([str(exe), str(ini2), "--max-steps", "10", "--quiet"], cwd=b, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "Restart at iteration 1" in r.stdout
    fa, fb = h5mini.File(a / "run.h5"), h5mini.File(b / "run.h5")
    # restarted run: truncated file re-saved the loaded state as ite_0001, then wrote ite_0002 at the end
    assert [k for k in fb.keys() if k.startswith("ite_")] == ["ite_0001", "ite_0002"]
    ta, tb = float(fa["ite_0001"].attrs["time"]), float(fb["ite_0002"].attrs["time"])
    assert abs(ta - tb) <= 1e-13 * ta
    for fld in ("rho", "u", "v", "prs"):
        assert rel_l1(fb["ite_0002/" + fld].read(), fa["ite_0001/" + fld].read()) <= TOL_L1
