"""The streamed host path (fv2d_advance_host_stream): a state that lives on the host, moved in row
blocks with upload / sweep / download overlapped, the caller's dt checked on the way.

What is held here:
  * a chain of streamed calls reproduces the reference's goldens within BASELINE.json's bars;
  * the hint changes how long a call takes, never what it returns: with a good hint, without a
    hint and with a wrong hint the outputs (state, dt used, next dt) are the same BITS;
  * from the second call of a chain on the step really is streamed (the hint is accepted);
  * the context's clock, step counter and dt history follow the chain.
FV2D_STREAM_ROWS=16 makes the small fixtures move in many row blocks (default: 256 rows).
"""
import os

import numpy as np
import pytest

import oracle_lib as O
from conftest import GOLDEN_NAMES, load_golden, rel_l1
from fv2d_b200 import capi

pytestmark = pytest.mark.gpu

TOL_L1 = 1e-12  # BASELINE.json: relative L1 of conserved fields after 10 steps
TOL_DT = 1e-13  # BASELINE.json: dt sequence, relative


@pytest.fixture(autouse=True)
def small_blocks(monkeypatch):
    monkeypatch.setenv("FV2D_STREAM_ROWS", "16")


def _chain(dev, run, Q0, nsteps, hints="good"):
    """nsteps calls, each fed the previous call's output; returns (Q, dts used, next dts, streamed flags, ctx info)."""
    a, b = Q0.copy(), np.empty_like(Q0)
    used, nxt, flags = [], [], []
    with capi.Context(dev, run.time_stepping, run.epsilon_reset_negative) as ctx:
        hint = 0.0
        for k in range(nsteps):
            if hints == "none":
                hint = 0.0
            elif hints == "wrong" and k > 0:
                hint = hint * (1.0 + 2.0 ** -40)
            du, dn, st = ctx.advance_host_stream(a, b, hint)
            used.append(du), nxt.append(dn), flags.append(st)
            hint = dn
            a, b = b, a
        t, _, steps = ctx.get_time()
        hist = ctx.dt_history(nsteps)
        U = ctx.download_U()
        neg = ctx.negative_counts()
    return a, np.array(used), np.array(nxt), flags, (t, steps, hist, U, neg)


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_streamed_chain_matches_reference(name):
    g = load_golden(name)
    dev, run = capi.params_from_ini(g.ini_path())
    Q0 = capi.init_problem(dev, run)
    Q, used, nxt, flags, (t, steps, hist, U, neg) = _chain(dev, run, Q0, g.nsteps)
    assert steps == g.nsteps and neg == [0, 0, 0]
    assert np.array_equal(hist, used)
    assert np.array_equal(used[1:], nxt[:-1])  # the dt a step announces is the dt the next one takes
    assert np.max(np.abs(used - g.dts) / g.dts) <= TOL_DT
    assert abs(t - g.t) <= TOL_DT * g.t
    Qd, Ud = O.domain(dev, Q), O.domain(dev, U)
    assert rel_l1(Qd, g.QN) <= TOL_L1 and rel_l1(Ud, g.UN) <= TOL_L1
    # speculation needs a single forward-Euler sweep per step
    if run.time_stepping == 0:
        assert flags[0] is False and all(flags[1:]), flags
    else:
        assert not any(flags)


@pytest.mark.parametrize("name", ["kh_plm_128x64", "c91_plm_64x32", "c91_bctc_64x32", "rt_wb_plm_32x96", "rt_fslp_32x96", "blast_64", "gresho_rk2_32"])
def test_hint_never_changes_the_result(name):
    g = load_golden(name)
    dev, run = capi.params_from_ini(g.ini_path())
    Q0 = capi.init_problem(dev, run)
    n = min(g.nsteps, 5)
    ref = _chain(dev, run, Q0, n, "none")
    assert not any(ref[3])
    for mode in ("good", "wrong"):
        got = _chain(dev, run, Q0, n, mode)
        assert np.array_equal(got[0], ref[0]), mode       # the whole host array, ghost cells included
        assert np.array_equal(got[1], ref[1]) and np.array_equal(got[2], ref[2]), mode
        assert got[4][0] == ref[4][0] and got[4][1] == ref[4][1], mode
        assert np.array_equal(got[4][2], ref[4][2]) and np.array_equal(got[4][3], ref[4][3]), mode
        if mode == "wrong":
            assert not any(got[3])  # every wrong hint was caught


def test_streamed_equals_resident_run_given_the_same_first_dt():
    """The host round trip is lossless.  A chain hands only Q from call to call, so every call starts
    from U = primToCons(Q); a resident run that does the same between its steps (and takes the same
    first dt) stays bitwise together with the chain: same dt source (the sweep's CFL maximum), same
    sweep, whether the rows arrive from HBM or in blocks from the host."""
    g = load_golden("kh_plm_128x64")
    dev, run = capi.params_from_ini(g.ini_path())
    Q0 = capi.init_problem(dev, run)
    n = 6
    a, b = Q0.copy(), np.empty_like(Q0)
    with capi.Context(dev, run.time_stepping, run.epsilon_reset_negative) as ctx:
        dt0, dn, _ = ctx.advance_host_stream(a, b, 0.0)
    with capi.Context(dev, run.time_stepping, run.epsilon_reset_negative) as ctx:
        ctx.upload_Q(Q0)
        ctx.prim_to_cons()
        ctx.step(dt0)
        for _ in range(n - 1):
            ctx.prim_to_cons()
            ctx.run_steps(1)
        Ures, Qres, hres = ctx.download_U(), ctx.download_Q(), ctx.dt_history(n)
    Q, used, _, flags, (_, _, hist, U, _) = _chain(dev, run, Q0, n)
    assert np.array_equal(hist, hres) and np.array_equal(U, Ures) and np.array_equal(Q, Qres)


def test_block_size_does_not_matter(monkeypatch):
    g = load_golden("kh_plm_128x64")
    dev, run = capi.params_from_ini(g.ini_path())
    Q0 = capi.init_problem(dev, run)
    outs = []
    for rows in ("16", "24", "64"):
        monkeypatch.setenv("FV2D_STREAM_ROWS", rows)
        outs.append(_chain(dev, run, Q0, 4))
    for o in outs[1:]:
        assert np.array_equal(o[0], outs[0][0]) and np.array_equal(o[1], outs[0][1]) and o[3] == outs[0][3]


def test_negative_counters_survive_a_rejected_speculation():
    """A state with negative pressures: the speculative sweeps count their resets, the rejected
    step must not leave them in the cumulative counters twice."""
    g = load_golden("blast_64")
    dev, run = capi.params_from_ini(g.ini_path())
    Q0 = capi.init_problem(dev, run)
    out = []
    for hint_scale in (0.0, 1.0 + 2.0 ** -30):
        a, b = Q0.copy(), np.empty_like(Q0)
        with capi.Context(dev, run.time_stepping, 1e-5) as ctx:
            _, dn, _ = ctx.advance_host_stream(a, b, 0.0)
            # poke negative pressures into the state between two calls (domain cells)
            b[3, dev.Ng + 5: dev.Ng + 9, dev.Ng + 7: dev.Ng + 11] = -1.0
            du, dn2, st = ctx.advance_host_stream(b, a, dn * hint_scale)
            assert st is False  # the poked state has another CFL maximum (hint 0: no speculation at all)
            out.append((a.copy(), du, dn2, ctx.negative_counts()))
    assert np.array_equal(out[0][0], out[1][0], equal_nan=True) and out[0][1:3] == out[1][1:3]
    assert out[0][3] == out[1][3] and sum(out[0][3]) > 0


@pytest.mark.parametrize("bx", ["absorbing", "reflecting", "periodic"])
@pytest.mark.parametrize("by", ["absorbing", "reflecting", "periodic"])
@pytest.mark.parametrize("ng", [2, 3])
def test_every_boundary_combination_and_ghost_depth(bx, by, ng):
    """The streamed path fills the ghost cells of the rows it uploads itself (low y-ghost rows with the
    first block - a periodic boundary needs the LAST domain rows for them - high ones with the last):
    all nine boundary combinations, Nghosts 2 and 3, on a grid of several row blocks, against the
    serial route of the same entry point (bitwise) and against the resident fused run (tolerance)."""
    g = load_golden("blast_64")
    ov = {"run.boundaries_x": bx, "run.boundaries_y": by, "mesh.Nghosts": ng, "mesh.Nx": 300, "mesh.Ny": 70}
    dev, run = capi.params_from_ini(g.ini_path(), ov)
    Q0 = capi.init_problem(dev, run)
    n = 4
    ref = _chain(dev, run, Q0, n, "none")
    got = _chain(dev, run, Q0, n, "good")
    assert got[3] == [False] + [True] * (n - 1)
    assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1]) and np.array_equal(got[4][3], ref[4][3])
    with capi.Context(dev, run.time_stepping, run.epsilon_reset_negative) as ctx:
        ctx.upload_Q(Q0)
        ctx.prim_to_cons()
        ctx.compute_dt()
        ctx.run_steps(n)
        Qres, hres = ctx.download_Q(), ctx.dt_history(n)
    assert np.max(np.abs(got[1] - hres) / hres) <= TOL_DT
    assert rel_l1(got[0], Qres) <= TOL_L1  # whole arrays, ghost cells included


def test_in_place_call_takes_the_serial_route_and_gives_the_same_bits():
    g = load_golden("kh_plm_128x64")
    dev, run = capi.params_from_ini(g.ini_path())
    Q0 = capi.init_problem(dev, run)
    ref = _chain(dev, run, Q0, 3)
    a = Q0.copy()
    with capi.Context(dev, run.time_stepping, run.epsilon_reset_negative) as ctx:
        hint = 0.0
        for _ in range(3):
            _, hint, st = ctx.advance_host_stream(a, a, hint)
            assert st is False
    assert np.array_equal(a, ref[0])
