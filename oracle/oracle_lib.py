"""TEST INFRASTRUCTURE ONLY: ctypes loader of the plain-C CPU restatement (oracle/muscl_oracle.c,
oracle/mhd_oracle.c) and a runner for the reference binary (oracle/_ref/euler_kokkos_ref).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "libmuscl_oracle.so")
REF_BIN = os.path.join(_HERE, "_ref", "euler_kokkos_ref")


def build(force: bool = False) -> str:
    if force or not os.path.exists(LIB_PATH) or any(
        os.path.getmtime(os.path.join(_HERE, s)) > os.path.getmtime(LIB_PATH)
        for s in ("muscl_oracle.c", "mhd_oracle.c", "muscl_oracle.h") if os.path.exists(os.path.join(_HERE, s))
    ):
        subprocess.run(["make", "-C", _HERE, "-s", "-B"], check=True)
    return LIB_PATH


class OrcParams(C.Structure):
    """struct orc_params (oracle/muscl_oracle.h) -- same field order as ek_params."""

    _fields_ = [
        ("dim", C.c_int32), ("mhd", C.c_int32), ("nbvar", C.c_int32), ("ghost", C.c_int32),
        ("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32),
        ("riemann", C.c_int32), ("impl_version", C.c_int32),
        ("bc", C.c_int32 * 6),
        ("gravity_enabled", C.c_int32), ("gravity_hancock", C.c_int32), ("reserved0", C.c_int32),
        ("dx", C.c_double), ("dy", C.c_double), ("dz", C.c_double),
        ("gamma0", C.c_double), ("gamma6", C.c_double), ("smallr", C.c_double), ("smallc", C.c_double),
        ("smallp", C.c_double), ("smallpp", C.c_double), ("cfl", C.c_double), ("slope_type", C.c_double),
        ("reserved1", C.c_double * 4),
    ]


def params_from(p) -> OrcParams:
    """Copy any ctypes struct with the same layout (e.g. capi.EkParams) into an OrcParams."""
    o = OrcParams()
    assert C.sizeof(o) == C.sizeof(p)
    C.memmove(C.byref(o), C.byref(p), C.sizeof(o))
    return o


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB_PATH)
        dp = C.POINTER(C.c_double)
        pp = C.POINTER(OrcParams)
        _lib.orc_make_boundaries.argtypes = [pp, dp]
        _lib.orc_make_boundaries.restype = None
        _lib.orc_hydro_invdt.argtypes = [pp, dp, dp]
        _lib.orc_hydro_invdt.restype = C.c_double
        _lib.orc_hydro_prims.argtypes = [pp, dp, dp]
        _lib.orc_hydro_prims.restype = None
        _lib.orc_hydro_step.argtypes = [pp, dp, dp, C.c_double, dp]
        _lib.orc_hydro_step.restype = None
        _lib.orc_riemann_hydro.argtypes = [pp, dp, dp, dp]
        _lib.orc_riemann_hydro.restype = None
        if hasattr(_lib, "orc_mhd_step"):
            _lib.orc_mhd_prims.argtypes = [pp, dp, dp]
            _lib.orc_mhd_prims.restype = None
            _lib.orc_mhd_invdt.argtypes = [pp, dp]
            _lib.orc_mhd_invdt.restype = C.c_double
            _lib.orc_mhd_step.argtypes = [pp, dp, dp, C.c_double, C.c_double, C.c_double, dp]
            _lib.orc_mhd_step.restype = None
            _lib.orc_riemann_mhd.argtypes = [pp, dp, dp, dp]
            _lib.orc_riemann_mhd.restype = None
    return _lib


def _dp(a):
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_double))


def make_boundaries(p, U):
    lib().orc_make_boundaries(C.byref(p), _dp(U))


def invdt(p, U, grav=None) -> float:
    """max inverse time step of the current state (ghosts must be filled for MHD)."""
    if p.mhd:
        Q = np.zeros_like(U)
        lib().orc_mhd_prims(C.byref(p), _dp(U), _dp(Q))
        return lib().orc_mhd_invdt(C.byref(p), _dp(Q))
    return lib().orc_hydro_invdt(C.byref(p), _dp(U), _dp(grav))


def step(p, Uin, dt, grav=None):
    """One godunov_unsplit_impl with a given dt.  Uin's ghosts are filled in place. Returns Uout."""
    Uout = np.zeros_like(Uin)
    if p.mhd:
        used = C.c_double()
        lib().orc_mhd_step(C.byref(p), _dp(Uin), _dp(Uout), float(dt), 0.0, 1e300, C.byref(used))
    else:
        lib().orc_hydro_step(C.byref(p), _dp(Uin), _dp(Uout), float(dt), _dp(grav))
    return Uout


def run(p, U, nsteps, t=0.0, t_end=1e300, grav=None):
    """The reference's time loop (main.cpp:126-131 + next_iteration_impl) on the oracle.
    Returns (U, t, list of dt)."""
    U = np.ascontiguousarray(U.copy())
    dts = []
    for _ in range(nsteps):
        if t >= t_end - 1e-14:
            break
        if p.mhd:
            make_boundaries(p, U)  # MHD computes dt after the ghost fill (SolverMHDMuscl.cpp:468-484)
        dt = p.cfl / invdt(p, U, grav)
        if t + dt > t_end:
            dt = t_end - t
        U = step(p, U, dt, grav)
        t += dt
        dts.append(dt)
    return U, t, dts


def have_reference() -> bool:
    return os.path.exists(REF_BIN) and os.access(REF_BIN, os.X_OK)


def run_reference(ini_path: str, cwd: str, threads: int | None = None, timeout: float = 3600.0) -> str:
    """Run the UNMODIFIED reference binary on an ini file inside `cwd`; returns its stdout."""
    env = dict(os.environ)
    env["OMP_NUM_THREADS"] = str(threads or os.cpu_count() or 1)
    env.setdefault("OMP_PROC_BIND", "spread")
    env.setdefault("OMP_PLACES", "threads")
    r = subprocess.run([REF_BIN, ini_path], cwd=cwd, env=env, capture_output=True, text=True, timeout=timeout)
    if r.returncode != 0:
        raise RuntimeError(f"reference binary failed ({r.returncode}):\n{r.stdout[-2000:]}\n{r.stderr[-2000:]}")
    return r.stdout
