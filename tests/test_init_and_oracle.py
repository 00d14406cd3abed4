"""The oracle (oracle/fv2d_oracle.c) and the host Init mirror, pinned against fixtures
dumped by the reference itself (tests/golden/make_goldens.py).  CPU-only."""
import numpy as np
import pytest

import oracle_lib as O
from conftest import load_golden
from fv2d_b200 import capi


def test_init_bit_identical_to_reference(golden):
    dev, run = capi.params_from_ini(golden.ini_path())
    Q = capi.init_problem(dev, run)
    assert np.array_equal(O.domain(dev, Q), golden.Q0)
    # ghosts: fillBoundaries applied (Init.h:356-357) == the oracle's fill on the same domain
    Q2 = Q.copy()
    Q2[:, : dev.jbeg, :] = -7.0
    Q2[:, :, dev.iend:] = -7.0
    Q2[:, dev.jend:, :] = -7.0
    Q2[:, :, : dev.ibeg] = -7.0
    O.fill_boundaries(dev, Q2)
    assert np.array_equal(Q, Q2)


def test_oracle_bit_identical_to_reference(golden):
    dev, run = capi.params_from_ini(golden.ini_path())
    Q = capi.init_problem(dev, run)
    U = O.prim_to_cons(dev, Q)
    n, t, dts, neg = O.run(dev, run.time_stepping, run.epsilon_reset_negative, run.tend, Q, U, golden.nsteps)
    assert n == golden.nsteps
    assert np.array_equal(dts, golden.dts)
    assert t == golden.t
    assert np.array_equal(O.domain(dev, Q), golden.QN)
    assert np.array_equal(O.domain(dev, U), golden.UN)
    assert neg == [0, 0, 0]
    # the reference's only conservation diagnostic (plot_energy_evolution.py:28-47)
    UN = O.domain(dev, U)
    mass = float(np.sum(UN[0].ravel() * dev.dx * dev.dy))
    assert abs(mass - golden.mass) <= 1e-13 * abs(golden.mass)


def test_baseline_md_anchors():
    """the anchors the survey recorded from the reference build."""
    g = load_golden("sod_x")
    assert list(g.dts[:3]) == [0.00096824586252223176, 0.0009498801116487283, 0.00088692715039051105]


def test_baseline_md_anchors_kh_256():
    """BASELINE.md §6, second row of anchors: Kelvin-Helmholtz 256^2 PLM, dt0 / dt9 and the
    domain sums after 10 steps, as the survey recorded them from the reference build.  The oracle is
    bit-identical to the reference on every fixture, so it must land on the same digits (the sums are
    accumulated sequentially like the survey's driver did; a last-digit difference of the summation
    order is allowed for)."""
    dev, run = capi.params_from_ini(load_golden("kh_plm_128x64").ini_path(), {"mesh.Nx": 256, "mesh.Ny": 256})
    Q = capi.init_problem(dev, run)
    U = O.prim_to_cons(dev, Q)
    n, t, dts, neg = O.run(dev, run.time_stepping, run.epsilon_reset_negative, run.tend, Q, U, 10)
    assert n == 10 and neg == [0, 0, 0]
    assert dts[0] == 0.00094323644991366196 and dts[9] == 0.00094313650081637775
    UN = O.domain(dev, U)
    mass, energy = 0.0, 0.0
    for v in UN[0].ravel():
        mass += v * dev.dx * dev.dy
    for v in UN[3].ravel():
        energy += v * dev.dx * dev.dy
    assert abs(mass - 11.999999999586402) <= 2e-15 * 12.0
    assert abs(energy - 125.40006088566672) <= 2e-15 * 125.4


def test_sod_x_y_symmetry():
    """sod_y is sod_x transposed with u<->v (SURVEY.md §4 known-answer check)."""
    gx, gy = load_golden("sod_x"), load_golden("sod_y")
    assert np.array_equal(gx.dts, gy.dts)
    assert np.allclose(gx.QN[0], gy.QN[0].T, rtol=0, atol=1e-15)
    assert np.allclose(gx.QN[1], gy.QN[2].T, rtol=0, atol=1e-15)
    assert np.allclose(gx.QN[3], gy.QN[3].T, rtol=0, atol=1e-15)


def _uniform(dev, rho=1.3, u=0.4, v=-0.2, p=2.0):
    Q = np.zeros(dev.shape())
    Q[0], Q[1], Q[2], Q[3] = rho, u, v, p
    return Q


@pytest.mark.parametrize("solver", [capi.HLL, capi.HLLC, capi.FSLP])
@pytest.mark.parametrize("recon", [capi.PCM, capi.PLM])
def test_uniform_state_is_preserved(solver, recon):
    dev, run = capi.params_from_ini(load_golden("blast_64").ini_path())
    dev.riemann_solver, dev.reconstruction = solver, recon
    Q = _uniform(dev)
    U = O.prim_to_cons(dev, Q)
    U0 = U.copy()
    O.run(dev, 0, 1e-8, 1e9, Q, U, 5)
    assert np.allclose(O.domain(dev, U), O.domain(dev, U0), rtol=1e-14, atol=0)


def test_periodic_blast_conserves_mass_and_energy():
    g = load_golden("blast_64")
    dev, run = capi.params_from_ini(g.ini_path())
    Q = capi.init_problem(dev, run)
    U = O.prim_to_cons(dev, Q)
    m0 = O.domain(dev, U).sum(axis=(1, 2))
    O.run(dev, 0, 1e-8, 1e9, Q, U, 20)
    m1 = O.domain(dev, U).sum(axis=(1, 2))
    assert abs(m1[0] - m0[0]) <= 1e-13 * abs(m0[0])
    assert abs(m1[3] - m0[3]) <= 1e-13 * abs(m0[3])


def test_riemann_solvers_consistency():
    """F(q, q) is the physical flux for every solver (a known-answer test the domain offers)."""
    dev, _ = capi.params_from_ini(load_golden("sod_x").ini_path())
    q = np.array([1.1, 0.3, -0.4, 0.9])
    E = 0.5 * q[0] * (q[1] ** 2 + q[2] ** 2) + q[3] / (dev.gamma0 - 1.0)
    phys = np.array([q[0] * q[1], q[0] * q[1] ** 2 + q[3], q[0] * q[1] * q[2], (E + q[3]) * q[1]])
    for solver in (capi.HLL, capi.HLLC, capi.FSLP):
        f, pout = O.riemann(dev, solver, q, q)
        assert np.allclose(f, phys, rtol=1e-14), solver
        assert abs(pout - q[3]) <= 1e-15
    # supersonic to the right: upwind = left flux exactly (HLL / HLLC)
    qL, qR = np.array([1.0, 5.0, 0.1, 1.0]), np.array([0.5, 5.5, -0.1, 0.8])
    EL = 0.5 * qL[0] * (qL[1] ** 2 + qL[2] ** 2) + qL[3] / (dev.gamma0 - 1.0)
    physL = np.array([qL[0] * qL[1], qL[0] * qL[1] ** 2 + qL[3], qL[0] * qL[1] * qL[2], (EL + qL[3]) * qL[1]])
    for solver in (capi.HLL, capi.HLLC):
        f, pout = O.riemann(dev, solver, qL, qR)
        assert np.allclose(f, physL, rtol=1e-15) and pout == qL[3]


def test_check_negatives_resets_and_counts():
    dev, _ = capi.params_from_ini(load_golden("sod_x").ini_path())
    Q = _uniform(dev)
    Q[0, dev.jbeg + 1, dev.ibeg + 3] = -1.0
    Q[3, dev.jbeg + 2, dev.ibeg + 5] = -2.0
    Q[1, dev.jbeg, dev.ibeg] = np.nan
    Q[0, 0, 0] = -5.0  # ghost: outside range_dom, untouched
    c = O.check_negatives(dev, 1e-8, Q)
    assert c == [1, 1, 1]
    assert Q[0, dev.jbeg + 1, dev.ibeg + 3] == 1e-8 and Q[3, dev.jbeg + 2, dev.ibeg + 5] == 1e-8 and Q[0, 0, 0] == -5.0
