"""TEST INFRASTRUCTURE ONLY: ctypes loader of the host build of the per-cell CUDA stages (mhd_emul.cpp)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def lib():
    global _lib
    if _lib is None:
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
        _lib = C.CDLL(os.path.join(_HERE, "_build", "libek_emul.so"))
        dp = C.POINTER(C.c_double)
        _lib.emul_mhd_invdt.restype = C.c_double
        _lib.emul_mhd_invdt.argtypes = [C.c_void_p, dp, dp]
        _lib.emul_mhd_work_len.restype = C.c_longlong
        _lib.emul_mhd_work_len.argtypes = [C.c_void_p]
        _lib.emul_mhd_step.argtypes = [C.c_void_p, dp, dp, dp, C.c_double]
        _lib.emul_mhd_update_invdt.restype = C.c_double
        _lib.emul_mhd_update_invdt.argtypes = [C.c_void_p, dp, dp, dp, C.c_double]
        _lib.emul_mhd_invdt_layers.restype = C.c_double
        _lib.emul_mhd_invdt_layers.argtypes = [C.c_void_p, dp]
        _lib.emul_hydro_step.argtypes = [C.c_void_p, dp, dp, dp, C.c_double]
        _lib.emul_hydro_invdt.restype = C.c_double
        _lib.emul_hydro_invdt.argtypes = [C.c_void_p, dp, dp]
        _lib.emul_chunk_map_check.restype = C.c_longlong
        _lib.emul_chunk_map_check.argtypes = [C.c_int] * 4
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def mhd_invdt(p, U):
    work = np.zeros(lib().emul_mhd_work_len(C.byref(p)))
    return lib().emul_mhd_invdt(C.byref(p), _dp(U), _dp(work))


def mhd_step(p, U, dt):
    """U: ghost-filled state.  Returns the new state (ghosts copied from U)."""
    work = np.zeros(lib().emul_mhd_work_len(C.byref(p)))
    out = U.copy()
    rc = lib().emul_mhd_step(C.byref(p), _dp(U), _dp(out), _dp(work), float(dt))
    assert rc == 0
    return out


def mhd_step_with_partial_invdt(p, U, dt):
    """3-D step through the update stage of mhd_update_kernel<true>: (new state, partial max of the next inverse dt)."""
    work = np.zeros(lib().emul_mhd_work_len(C.byref(p)))
    out = U.copy()
    assert lib().emul_mhd_step(C.byref(p), _dp(U), _dp(out), _dp(work), float(dt)) == 0
    out2 = U.copy()
    partial = lib().emul_mhd_update_invdt(C.byref(p), _dp(U), _dp(out2), _dp(work), float(dt))
    assert np.array_equal(out, out2)
    return out, partial


def mhd_invdt_layers(p, U):
    return lib().emul_mhd_invdt_layers(C.byref(p), _dp(U))


def chunk_map_violations(isize, jsize, ksize, yc):
    return lib().emul_chunk_map_check(isize, jsize, ksize, yc)


def hydro_step(p, U, dt, grav=None):
    """U: ghost-filled state.  Returns the new state (ghosts copied from U) computed by the host build of the
    per-cell functions of hydro_step_kernel, in the kernel's order of operations."""
    out = U.copy()
    g = _dp(grav) if grav is not None else None
    rc = lib().emul_hydro_step(C.byref(p), _dp(U), _dp(out), g, float(dt))
    assert rc == 0
    return out


def hydro_invdt(p, U, grav=None):
    g = _dp(grav) if grav is not None else None
    return lib().emul_hydro_invdt(C.byref(p), _dp(U), g)
