"""ctypes access to the CPU oracle (oracle/libfv2d_oracle.so).  TEST INFRASTRUCTURE ONLY:
imported by tests/, bench.py's cpu_baseline / --impl reference legs and
__graft_entry__.smoke(); never by the product package."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
ORACLE_DIR = ROOT / "oracle"
ORACLE_SO = ORACLE_DIR / "libfv2d_oracle.so"
REF_BIN = ORACLE_DIR / "_ref" / "fv2d_ref"

import sys

sys.path.insert(0, str(ROOT))
from fv2d_b200.capi import DeviceParams  # noqa: E402  (POD layout shared through include/fv2d_params.h)

_dp = C.POINTER(C.c_double)
_lib = None


def build():
    subprocess.run(["make", "-C", str(ORACLE_DIR), "oracle"], check=True, capture_output=True)


def lib():
    global _lib
    if _lib is None:
        if not ORACLE_SO.exists():
            build()
        L = C.CDLL(str(ORACLE_SO))
        P = C.POINTER(DeviceParams)
        L.fv2d_oracle_compute_dt.restype = C.c_double
        L.fv2d_oracle_compute_dt.argtypes = [P, _dp, _dp]
        L.fv2d_oracle_get_gravity.restype = C.c_double
        L.fv2d_oracle_get_gravity.argtypes = [P, C.c_int, C.c_int, C.c_int]
        L.fv2d_oracle_run.restype = C.c_long
        L.fv2d_oracle_run.argtypes = [P, C.c_int, C.c_double, C.c_double, _dp, _dp, C.c_long, _dp, _dp,
                                      C.POINTER(C.c_uint64)]
        L.fv2d_oracle_riemann.restype = None
        L.fv2d_oracle_riemann.argtypes = [P, C.c_int, _dp, _dp, C.c_double, _dp, _dp]
        for name, args in {
            "fv2d_oracle_cons_to_prim": [P, _dp, _dp],
            "fv2d_oracle_prim_to_cons": [P, _dp, _dp],
            "fv2d_oracle_check_negatives": [P, C.c_double, _dp, C.POINTER(C.c_uint64)],
            "fv2d_oracle_fill_boundaries": [P, _dp],
            "fv2d_oracle_compute_slopes": [P, _dp, _dp, _dp],
            "fv2d_oracle_compute_fluxes_and_update": [P, _dp, _dp, _dp, _dp, C.c_double],
            "fv2d_oracle_apply_viscosity": [P, _dp, _dp, C.c_double],
        }.items():
            getattr(L, name).argtypes = args
            getattr(L, name).restype = None
        for name, args in {
            "fv2d_oracle_apply_thermal_conduction": [P, _dp, _dp, C.c_double],
            "fv2d_oracle_euler_step": [P, _dp, _dp, C.c_double],
            "fv2d_oracle_update": [P, C.c_int, _dp, _dp, C.c_double],
        }.items():
            getattr(L, name).argtypes = args
            getattr(L, name).restype = C.c_int
        _lib = L
    return _lib


def _p(a):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_dp)


def prim_to_cons(dev, Q):
    U = np.zeros_like(Q)
    lib().fv2d_oracle_prim_to_cons(C.byref(dev), _p(Q), _p(U))
    return U


def cons_to_prim(dev, U):
    Q = np.zeros_like(U)
    lib().fv2d_oracle_cons_to_prim(C.byref(dev), _p(U), _p(Q))
    return Q


def fill_boundaries(dev, Q):
    lib().fv2d_oracle_fill_boundaries(C.byref(dev), _p(Q))


def compute_dt(dev, Q):
    inv = np.zeros(3)
    dt = lib().fv2d_oracle_compute_dt(C.byref(dev), _p(Q), _p(inv))
    return dt, inv


def check_negatives(dev, eps, Q):
    c = (C.c_uint64 * 3)()
    lib().fv2d_oracle_check_negatives(C.byref(dev), eps, _p(Q), c)
    return [int(v) for v in c]


def compute_slopes(dev, Q):
    sx, sy = np.zeros_like(Q), np.zeros_like(Q)
    lib().fv2d_oracle_compute_slopes(C.byref(dev), _p(Q), _p(sx), _p(sy))
    return sx, sy


def compute_fluxes_and_update(dev, Q, sx, sy, U, dt):
    lib().fv2d_oracle_compute_fluxes_and_update(C.byref(dev), _p(Q), _p(sx), _p(sy), _p(U), dt)


def apply_thermal_conduction(dev, Q, U, dt):
    return lib().fv2d_oracle_apply_thermal_conduction(C.byref(dev), _p(Q), _p(U), dt)


def apply_viscosity(dev, Q, U, dt):
    lib().fv2d_oracle_apply_viscosity(C.byref(dev), _p(Q), _p(U), dt)


def update(dev, time_stepping, Q, U, dt):
    return lib().fv2d_oracle_update(C.byref(dev), time_stepping, _p(Q), _p(U), dt)


def riemann(dev, solver, qL, qR, gdx=0.0):
    qL = np.ascontiguousarray(qL, dtype=np.float64)
    qR = np.ascontiguousarray(qR, dtype=np.float64)
    flux = np.zeros(4)
    pout = C.c_double()
    lib().fv2d_oracle_riemann(C.byref(dev), solver, _p(qL), _p(qR), gdx, _p(flux), C.cast(C.byref(pout), _dp))
    return flux, pout.value


def run(dev, time_stepping, eps_reset, tend, Q, U, max_steps, t0=0.0):
    """main.cpp:62-84 without IO; Q and U are advanced in place.  Returns (steps, t, dts, neg_counts)."""
    t = C.c_double(t0)
    dts = np.zeros(max(max_steps, 1))
    neg = (C.c_uint64 * 3)()
    n = lib().fv2d_oracle_run(C.byref(dev), time_stepping, eps_reset, tend, _p(Q), _p(U), max_steps,
                              C.cast(C.byref(t), _dp), _p(dts), neg)
    if n < 0:
        raise RuntimeError("oracle: unsupported configuration")
    return n, t.value, dts[:n].copy(), [int(v) for v in neg]


def domain(dev, A):
    """[f][Nty][Ntx] -> domain-only [f][Ny][Nx]"""
    return A[:, dev.jbeg:dev.jend, dev.ibeg:dev.iend]
