"""ctypes binding of libfv2d_b200.so — a 1:1 view of include/fv2d_b200.h.

This is plumbing for tests, bench.py and the Python drivers; all computation happens in
the CUDA library.  There is no CPU fallback: if the shared library is missing the import
of :func:`lib` raises, and on a machine without an sm_100 device every compute call fails
with ``Fv2dError`` (FV2D_ERR_CUDA).
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("FV2D_B200_LIB", _HERE / "libfv2d_b200.so"))

FV2D_DT_HISTORY = 4096
FV2D_IPC_HANDLE_BYTES = 512

# enums of include/fv2d_params.h
HLL, HLLC, FSLP = 0, 1, 2
BC_ABSORBING, BC_REFLECTING, BC_PERIODIC = 0, 1, 2
TS_EULER, TS_RK2 = 0, 1
PCM, PCM_WB, PLM = 0, 1, 2
TCM_CONSTANT, TCM_B02 = 0, 1
BCTC_NONE, BCTC_FIXED_TEMPERATURE, BCTC_FIXED_GRADIENT = 0, 1, 2
GRAV_NONE, GRAV_CONSTANT, GRAV_ANALYTICAL = 0, 1, 2
IR, IU, IV, IP, IE = 0, 1, 2, 3, 3


class DeviceParams(C.Structure):
    """fv2d_device_params (mirrors the reference's DeviceParams, SimInfo.h:266-354)."""

    _fields_ = [
        ("gamma0", C.c_double),
        ("gravity_mode", C.c_int32),
        ("analytical_gravity_mode", C.c_int32),
        ("gx", C.c_double),
        ("gy", C.c_double),
        ("well_balanced_flux_at_y_bc", C.c_int32),
        ("well_balanced", C.c_int32),
        ("fslp_K", C.c_double),
        ("thermal_conductivity_active", C.c_int32),
        ("thermal_conductivity_mode", C.c_int32),
        ("kappa", C.c_double),
        ("bctc_ymin", C.c_int32),
        ("bctc_ymax", C.c_int32),
        ("bctc_ymin_value", C.c_double),
        ("bctc_ymax_value", C.c_double),
        ("viscosity_active", C.c_int32),
        ("viscosity_mode", C.c_int32),
        ("mu", C.c_double),
        ("m1", C.c_double),
        ("theta1", C.c_double),
        ("m2", C.c_double),
        ("theta2", C.c_double),
        ("h84_pert", C.c_double),
        ("c91_pert", C.c_double),
        ("b02_ymid", C.c_double),
        ("b02_kappa1", C.c_double),
        ("b02_kappa2", C.c_double),
        ("b02_thickness", C.c_double),
        ("hot_bubble_g0", C.c_double),
        ("kh_y1", C.c_double),
        ("kh_y2", C.c_double),
        ("kh_a", C.c_double),
        ("kh_sigma", C.c_double),
        ("kh_rho_fac", C.c_double),
        ("kh_uflow", C.c_double),
        ("kh_amp", C.c_double),
        ("kh_P0", C.c_double),
        ("gresho_density", C.c_double),
        ("gresho_Mach", C.c_double),
        ("boundary_x", C.c_int32),
        ("boundary_y", C.c_int32),
        ("reconstruction", C.c_int32),
        ("riemann_solver", C.c_int32),
        ("CFL", C.c_double),
        ("Nx", C.c_int32),
        ("Ny", C.c_int32),
        ("Ng", C.c_int32),
        ("Ntx", C.c_int32),
        ("Nty", C.c_int32),
        ("ibeg", C.c_int32),
        ("iend", C.c_int32),
        ("jbeg", C.c_int32),
        ("jend", C.c_int32),
        ("pad0_", C.c_int32),
        ("xmin", C.c_double),
        ("xmax", C.c_double),
        ("ymin", C.c_double),
        ("ymax", C.c_double),
        ("dx", C.c_double),
        ("dy", C.c_double),
        ("epsilon", C.c_double),
    ]

    def copy(self) -> "DeviceParams":
        out = DeviceParams()
        C.memmove(C.byref(out), C.byref(self), C.sizeof(DeviceParams))
        return out

    def shape(self):
        return (4, self.Nty, self.Ntx)


class RunParams(C.Structure):
    """fv2d_run_params (host-only members of the reference's Params, SimInfo.h:463-492)."""

    _fields_ = [
        ("save_freq", C.c_double),
        ("tend", C.c_double),
        ("epsilon_reset_negative", C.c_double),
        ("time_stepping", C.c_int32),
        ("multiple_outputs", C.c_int32),
        ("seed", C.c_int32),
        ("log_frequency", C.c_int32),
        ("problem", C.c_char * 64),
        ("filename_out", C.c_char * 256),
        ("output_path", C.c_char * 256),
        ("restart_file", C.c_char * 256),
    ]


class Fv2dError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"fv2d error {code}: {msg}")
        self.code = code


_lib = None
_dp = C.POINTER(C.c_double)
_ctxp = C.c_void_p


def lib() -> C.CDLL:
    """Load libfv2d_b200.so (built in-tree by `make lib`); raises if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(f"{LIB_PATH} not found: build it with `make lib` (there is no CPU fallback)")
    L = C.CDLL(str(LIB_PATH))
    L.fv2d_last_error.restype = C.c_char_p
    L.fv2d_abi_version.restype = C.c_int
    sig = {
        "fv2d_params_from_ini": [C.c_char_p, C.c_char_p, C.POINTER(DeviceParams), C.POINTER(RunParams)],
        "fv2d_params_dump_ini": [C.c_char_p, C.c_char_p, C.c_char_p],
        "fv2d_init_problem": [C.POINTER(DeviceParams), C.POINTER(RunParams), _dp],
        "fv2d_init_problem_rows": [C.POINTER(DeviceParams), C.POINTER(RunParams), C.c_int, C.c_int, _dp],
        "fv2d_io_save_solution": [C.POINTER(DeviceParams), C.POINTER(RunParams), _dp, C.c_int, C.c_double,
                                  C.POINTER(C.c_int)],
        "fv2d_io_load_snapshot": [C.POINTER(DeviceParams), C.POINTER(RunParams), _dp, _dp, C.POINTER(C.c_int),
                                  C.POINTER(C.c_int)],
        "fv2d_ctx_create": [C.POINTER(DeviceParams), C.c_int, C.c_double, C.c_int, C.POINTER(_ctxp)],
        "fv2d_ctx_create_slab": [C.POINTER(DeviceParams), C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, C.POINTER(_ctxp)],
        "fv2d_ctx_set_stream": [_ctxp, C.c_void_p],
        "fv2d_sync": [_ctxp],
        "fv2d_ctx_geometry": [_ctxp, C.POINTER(C.c_int64)],
        "fv2d_upload_Q": [_ctxp, _dp],
        "fv2d_upload_U": [_ctxp, _dp],
        "fv2d_download_Q": [_ctxp, _dp],
        "fv2d_download_U": [_ctxp, _dp],
        "fv2d_prim_to_cons": [_ctxp],
        "fv2d_cons_to_prim": [_ctxp],
        "fv2d_check_negatives": [_ctxp, C.POINTER(C.c_uint64)],
        "fv2d_fill_boundaries": [_ctxp],
        "fv2d_compute_dt": [_ctxp, _dp, _dp],
        "fv2d_compute_slopes": [_ctxp],
        "fv2d_compute_fluxes_and_update": [_ctxp, C.c_double],
        "fv2d_apply_thermal_conduction": [_ctxp, C.c_double],
        "fv2d_apply_viscosity": [_ctxp, C.c_double],
        "fv2d_euler_step": [_ctxp, C.c_double],
        "fv2d_update": [_ctxp, C.c_double],
        "fv2d_step": [_ctxp, C.c_double],
        "fv2d_step_device_dt": [_ctxp],
        "fv2d_run_steps": [_ctxp, C.c_int64],
        "fv2d_run_until": [_ctxp, C.c_double, C.c_int64, C.POINTER(C.c_int64)],
        "fv2d_get_time": [_ctxp, _dp, _dp, C.POINTER(C.c_int64)],
        "fv2d_set_time": [_ctxp, C.c_double],
        "fv2d_get_dt_history": [_ctxp, _dp, C.c_int64, C.POINTER(C.c_int64)],
        "fv2d_get_negative_counts": [_ctxp, C.POINTER(C.c_uint64), C.c_int],
        "fv2d_integrate_mass_energy": [_ctxp, _dp, _dp],
        "fv2d_advance_host": [_ctxp, _dp, _dp, C.c_int64, _dp],
        "fv2d_advance_host_stream": [_ctxp, _dp, _dp, C.c_double, _dp, _dp, C.POINTER(C.c_int)],
        "fv2d_profile_enable": [_ctxp, C.c_int],
        "fv2d_profile_read": [_ctxp, _dp, C.POINTER(C.c_int64), C.POINTER(C.c_int64)],
        "fv2d_halo_export": [_ctxp, C.c_void_p],
        "fv2d_halo_connect": [_ctxp, C.c_char_p, C.c_int],
        "fv2d_get_inv_dt": [_ctxp, _dp],
        "fv2d_debug_math_probe": [C.c_int, C.c_int64, _dp, _dp, _dp, _dp],
        "fv2d_state_hash": [_ctxp, C.POINTER(C.c_uint64)],
        "fv2d_debug_fp64_peak": [C.c_int, _dp],
        "fv2d_debug_sweep_timing": [_ctxp, C.POINTER(C.c_int64), C.c_int],
        "fv2d_debug_sync_wait": [_ctxp, _dp, _dp, _dp, C.c_int],
        "fv2d_device_count": [C.POINTER(C.c_int)],
        "fv2d_debug_stream_blocks": [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int32), C.c_int, C.POINTER(C.c_int)],
        "fv2d_debug_schedule": [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int32), C.c_int, C.POINTER(C.c_int)],
    }
    for name, argtypes in sig.items():
        fn = getattr(L, name)
        fn.argtypes = argtypes
        fn.restype = C.c_int
    L.fv2d_ctx_destroy.argtypes = [_ctxp]
    L.fv2d_ctx_destroy.restype = None
    _lib = L
    return L


def _check(rc: int):
    if rc != 0:
        raise Fv2dError(rc, lib().fv2d_last_error().decode(errors="replace"))


def _ptr(a: np.ndarray):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_dp)


def format_overrides(overrides) -> bytes | None:
    """{"mesh.Nx": 128, ...} -> b"mesh.Nx=128;..." (the `overrides` argument of the C ABI)."""
    if not overrides:
        return None
    if isinstance(overrides, (bytes, str)):
        return overrides.encode() if isinstance(overrides, str) else overrides
    return ";".join(f"{k}={v}" for k, v in overrides.items()).encode()


def params_from_ini(path, overrides=None) -> tuple[DeviceParams, RunParams]:
    """readInifile(filename) of the reference (SimInfo.h:529-568)."""
    dev, run = DeviceParams(), RunParams()
    _check(lib().fv2d_params_from_ini(os.fsencode(str(path)), format_overrides(overrides), C.byref(dev), C.byref(run)))
    return dev, run


def params_dump_ini(path, out_path, overrides=None) -> None:
    _check(lib().fv2d_params_dump_ini(os.fsencode(str(path)), format_overrides(overrides), os.fsencode(str(out_path))))


def init_problem(dev: DeviceParams, run: RunParams) -> np.ndarray:
    """InitFunctor(params).init(Q) of the reference (Init.h:310-358); returns Q[f][j][i]."""
    Q = np.zeros(dev.shape(), dtype=np.float64)
    _check(lib().fv2d_init_problem(C.byref(dev), C.byref(run), _ptr(Q)))
    return Q


def init_problem_rows(dev: DeviceParams, run: RunParams, j_first: int, nrows: int) -> np.ndarray:
    """Rows [j_first, j_first+nrows) of init_problem(dev, run), ghosts included."""
    Q = np.zeros((4, nrows, dev.Ntx), dtype=np.float64)
    _check(lib().fv2d_init_problem_rows(C.byref(dev), C.byref(run), j_first, nrows, _ptr(Q)))
    return Q


def io_save_solution(dev: DeviceParams, run: RunParams, Q: np.ndarray, iteration: int, t: float,
                     force_file_truncation: bool = False) -> bool:
    """IOManager::saveSolution on a host array (IOManager.h:99-282); returns the updated
    force_file_truncation flag."""
    assert Q.shape == dev.shape()
    flag = C.c_int(1 if force_file_truncation else 0)
    _check(lib().fv2d_io_save_solution(C.byref(dev), C.byref(run), _ptr(Q), iteration, t, C.byref(flag)))
    return bool(flag.value)


def io_load_snapshot(dev: DeviceParams, run: RunParams, force_file_truncation: bool = False):
    """IOManager::loadSnapshot (IOManager.h:284-398) -> (Q with ghosts, time, iteration,
    force_file_truncation)."""
    Q = np.zeros(dev.shape(), dtype=np.float64)
    t, it, flag = C.c_double(0.0), C.c_int(0), C.c_int(1 if force_file_truncation else 0)
    _check(lib().fv2d_io_load_snapshot(C.byref(dev), C.byref(run), _ptr(Q), C.byref(t), C.byref(it), C.byref(flag)))
    return Q, t.value, it.value, bool(flag.value)


def math_probe(a: np.ndarray, b: np.ndarray, device: int = 0):
    """(1/a, sqrt(a/b)) as the fused sweep's division-free primitives compute them."""
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    assert a.shape == b.shape and a.ndim == 1
    r, c = np.empty_like(a), np.empty_like(a)
    _check(lib().fv2d_debug_math_probe(device, a.size, _ptr(a), _ptr(b), _ptr(r), _ptr(c)))
    return r, c


def schedule_runs(Nx: int, Ny_local: int, num_sms: int = 148, neighbour_lo: bool = False, neighbour_hi: bool = False):
    """Row runs [(first, last), ...] of the persistent sweep's work table (host logic, no GPU needed)."""
    buf = (C.c_int32 * (2 * 65536))()
    n = C.c_int()
    _check(lib().fv2d_debug_schedule(Nx, Ny_local, num_sms, int(neighbour_lo), int(neighbour_hi), buf, 65536, C.byref(n)))
    return [(buf[2 * k], buf[2 * k + 1]) for k in range(n.value)]


def stream_blocks(Ny: int, Ng: int = 2, block_rows: int = 0):
    """Row blocks [(up0, up1, sw0, sw1), ...] of the streamed host path (host logic, no GPU needed)."""
    buf = (C.c_int32 * (4 * 65536))()
    n = C.c_int()
    _check(lib().fv2d_debug_stream_blocks(Ny, Ng, block_rows, buf, 65536, C.byref(n)))
    return [tuple(buf[4 * k + i] for i in range(4)) for k in range(n.value)]


def device_count() -> int:
    n = C.c_int()
    _check(lib().fv2d_device_count(C.byref(n)))
    return n.value


def fp64_peak(device: int = 0) -> float:
    """Measured fp64 peak of the device in thread-level DFMA instructions per second."""
    r = C.c_double()
    _check(lib().fv2d_debug_fp64_peak(device, C.byref(r)))
    return r.value


class Context:
    """Owner of one fv2d_ctx: device-resident Q/U of one (slab of a) grid."""

    def __init__(self, dev: DeviceParams, time_stepping: int = TS_EULER, eps_reset_negative: float = 1e-8,
                 device: int = 0, rank: int = 0, nranks: int = 1):
        self.dev = dev.copy()
        self._h = _ctxp()
        _check(lib().fv2d_ctx_create_slab(C.byref(self.dev), time_stepping, eps_reset_negative, device, rank, nranks,
                                          C.byref(self._h)))
        g = (C.c_int64 * 6)()
        _check(lib().fv2d_ctx_geometry(self._h, g))
        self.Ntx, self.Nty, self.Ny, self.j_offset, self.pitch, self.lead = (int(v) for v in g)

    def close(self):
        if self._h:
            lib().fv2d_ctx_destroy(self._h)
            self._h = _ctxp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @property
    def handle(self):
        return self._h

    def local_shape(self):
        return (4, self.Nty, self.Ntx)

    # transfers
    def upload_Q(self, Q: np.ndarray):
        assert Q.shape == self.local_shape(), (Q.shape, self.local_shape())
        _check(lib().fv2d_upload_Q(self._h, _ptr(Q)))

    def upload_U(self, U: np.ndarray):
        assert U.shape == self.local_shape()
        _check(lib().fv2d_upload_U(self._h, _ptr(U)))

    def download_Q(self) -> np.ndarray:
        Q = np.empty(self.local_shape(), dtype=np.float64)
        _check(lib().fv2d_download_Q(self._h, _ptr(Q)))
        return Q

    def download_U(self) -> np.ndarray:
        U = np.empty(self.local_shape(), dtype=np.float64)
        _check(lib().fv2d_download_U(self._h, _ptr(U)))
        return U

    def set_stream(self, cuda_stream: int):
        _check(lib().fv2d_ctx_set_stream(self._h, C.c_void_p(cuda_stream)))

    def sync(self):
        _check(lib().fv2d_sync(self._h))

    # operator-level API
    def prim_to_cons(self):
        _check(lib().fv2d_prim_to_cons(self._h))

    def cons_to_prim(self):
        _check(lib().fv2d_cons_to_prim(self._h))

    def check_negatives(self):
        c = (C.c_uint64 * 3)()
        _check(lib().fv2d_check_negatives(self._h, c))
        return [int(v) for v in c]

    def fill_boundaries(self):
        _check(lib().fv2d_fill_boundaries(self._h))

    def compute_dt(self):
        dt = C.c_double()
        inv = (C.c_double * 3)()
        _check(lib().fv2d_compute_dt(self._h, C.byref(dt), inv))
        return dt.value, [float(v) for v in inv]

    def compute_slopes(self):
        _check(lib().fv2d_compute_slopes(self._h))

    def compute_fluxes_and_update(self, dt: float):
        _check(lib().fv2d_compute_fluxes_and_update(self._h, dt))

    def apply_thermal_conduction(self, dt: float):
        _check(lib().fv2d_apply_thermal_conduction(self._h, dt))

    def apply_viscosity(self, dt: float):
        _check(lib().fv2d_apply_viscosity(self._h, dt))

    def euler_step(self, dt: float):
        _check(lib().fv2d_euler_step(self._h, dt))

    def update(self, dt: float):
        _check(lib().fv2d_update(self._h, dt))

    # fused hot path
    def step(self, dt: float):
        _check(lib().fv2d_step(self._h, dt))

    def step_device_dt(self):
        _check(lib().fv2d_step_device_dt(self._h))

    def run_steps(self, n: int):
        _check(lib().fv2d_run_steps(self._h, n))

    def run_until(self, tend: float, max_steps: int) -> int:
        n = C.c_int64()
        _check(lib().fv2d_run_until(self._h, tend, max_steps, C.byref(n)))
        return n.value

    def get_time(self):
        t, dt, n = C.c_double(), C.c_double(), C.c_int64()
        _check(lib().fv2d_get_time(self._h, C.byref(t), C.byref(dt), C.byref(n)))
        return t.value, dt.value, n.value

    def set_time(self, t: float):
        _check(lib().fv2d_set_time(self._h, t))

    def dt_history(self, n: int) -> np.ndarray:
        out = np.zeros(max(n, 1), dtype=np.float64)
        got = C.c_int64()
        _check(lib().fv2d_get_dt_history(self._h, _ptr(out), n, C.byref(got)))
        return out[: got.value].copy()

    def negative_counts(self, reset: bool = False):
        c = (C.c_uint64 * 3)()
        _check(lib().fv2d_get_negative_counts(self._h, c, int(reset)))
        return [int(v) for v in c]

    def mass_energy(self):
        m, e = C.c_double(), C.c_double()
        _check(lib().fv2d_integrate_mass_energy(self._h, C.byref(m), C.byref(e)))
        return m.value, e.value

    def state_hash(self) -> int:
        """Decomposition-independent 64-bit hash of the local slab's conserved state (slab hashes
        add modulo 2**64 to the hash of the whole grid)."""
        h = C.c_uint64()
        _check(lib().fv2d_state_hash(self._h, C.byref(h)))
        return int(h.value)

    def sweep_timing(self, n_ctas: int) -> np.ndarray:
        """Development hook (-DFV2D_TIMING builds): [n_ctas, 4] cycles outside / inside the row loops,
        items, total of the last sweep."""
        out = np.zeros((n_ctas, 4), dtype=np.int64)
        _check(lib().fv2d_debug_sweep_timing(self._h, out.ctypes.data_as(C.POINTER(C.c_int64)), 4 * n_ctas))
        return out

    def sync_wait(self, reset: bool = False):
        """(wait for the other ranks' CFL mails in the last sweep, accumulated wait since the last reset,
        busy time of the last sweep), microseconds."""
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        _check(lib().fv2d_debug_sync_wait(self._h, C.byref(a), C.byref(b), C.byref(c), int(reset)))
        return a.value, b.value, c.value

    def halo_export(self) -> bytes:
        buf = C.create_string_buffer(FV2D_IPC_HANDLE_BYTES)
        _check(lib().fv2d_halo_export(self._h, buf))
        return buf.raw

    def halo_connect(self, handles: bytes, nranks: int):
        assert len(handles) == nranks * FV2D_IPC_HANDLE_BYTES
        _check(lib().fv2d_halo_connect(self._h, handles, nranks))

    def inv_dt(self):
        inv = (C.c_double * 3)()
        _check(lib().fv2d_get_inv_dt(self._h, inv))
        return [float(v) for v in inv]

    def profile_enable(self, on: bool = True):
        _check(lib().fv2d_profile_enable(self._h, int(on)))

    def profile_read(self):
        ms, ns, nt = C.c_double(), C.c_int64(), C.c_int64()
        _check(lib().fv2d_profile_read(self._h, C.byref(ms), C.byref(ns), C.byref(nt)))
        return ms.value, ns.value, nt.value

    def advance_host(self, Q_in: np.ndarray, Q_out: np.ndarray, nsteps: int, dts: np.ndarray | None = None):
        _check(lib().fv2d_advance_host(self._h, _ptr(Q_in), _ptr(Q_out), nsteps, _ptr(dts) if dts is not None else None))

    def advance_host_stream(self, Q_in: np.ndarray, Q_out: np.ndarray, dt_hint: float = 0.0):
        """One step on a host-resident state with overlapped transfers (fv2d_advance_host_stream).
        Returns (dt_used, dt_next, streamed): pass dt_next as the next call's dt_hint."""
        used, nxt, st = C.c_double(0.0), C.c_double(0.0), C.c_int(0)
        _check(lib().fv2d_advance_host_stream(self._h, _ptr(Q_in), _ptr(Q_out), float(dt_hint), C.byref(used), C.byref(nxt),
                                              C.byref(st)))
        return used.value, nxt.value, bool(st.value)


def exported_symbols_in_header() -> list[str]:
    """Every function name declared in include/fv2d_b200.h (used by the symbol test)."""
    import re

    text = (_HERE.parent / "include" / "fv2d_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fv2d_[A-Za-z0-9_]+)\s*\(", text)))
