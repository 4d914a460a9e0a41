"""ctypes binding of the C++ host layer (lib/libek_host.so, csrc/host/host_capi.cpp): ini parsing with
the reference's ConfigMap semantics, HydroParams derivation, domain decomposition, initial conditions
and the euler_b200 driver as a function.  Used by bench.py and the tests."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import capi

_HERE = os.path.dirname(os.path.abspath(__file__))
HOST_LIB_PATH = os.path.join(_HERE, "lib", "libek_host.so")
SETTINGS_DIR = os.path.join(_HERE, "settings")
if not os.path.exists(HOST_LIB_PATH):
    raise ImportError(f"{HOST_LIB_PATH} not found: build it with `make -C euler_kokkos_b200/csrc/host`")

hlib = C.CDLL(HOST_LIB_PATH, mode=C.RTLD_GLOBAL)
_vp = C.c_void_p
hlib.ekh_case_create.restype = _vp
hlib.ekh_case_create.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int]
hlib.ekh_case_destroy.argtypes = [_vp]
hlib.ekh_case_params.argtypes = [_vp, C.POINTER(capi.EkParams)]
hlib.ekh_case_neighbors.argtypes = [_vp, C.POINTER(C.c_int32)]
hlib.ekh_case_run_info.argtypes = [_vp, C.POINTER(C.c_double)]
hlib.ekh_case_get_string.argtypes = [_vp, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int]
hlib.ekh_case_get_float.restype = C.c_double
hlib.ekh_case_get_float.argtypes = [_vp, C.c_char_p, C.c_char_p, C.c_float]
hlib.ekh_case_init_state.argtypes = [_vp, _vp, _vp]
hlib.ekh_case_write_ini.argtypes = [_vp, C.c_char_p]
hlib.ekh_case_save_vti.argtypes = [_vp, _vp, C.c_int, C.c_char_p, C.c_int]
hlib.ekh_solver_registered.argtypes = [C.c_char_p]
hlib.ekh_run_ini.argtypes = [C.c_char_p, C.c_char_p]

hlib.ekh_solver_create.restype = _vp
hlib.ekh_solver_create.argtypes = [C.c_char_p, C.c_char_p]
hlib.ekh_solver_destroy.argtypes = [_vp]
hlib.ekh_solver_finished.argtypes = [_vp]
hlib.ekh_solver_next_iteration.argtypes = [_vp]
hlib.ekh_solver_advance.argtypes = [_vp, C.c_int]
hlib.ekh_solver_save_solution.argtypes = [_vp]
hlib.ekh_solver_info.argtypes = [_vp, C.POINTER(C.c_double)]
hlib.ekh_solver_state.argtypes = [_vp, _vp, C.c_longlong]

HOST_SYMBOLS = ("ekh_case_create", "ekh_case_destroy", "ekh_case_params", "ekh_case_neighbors", "ekh_case_run_info",
                "ekh_case_get_string", "ekh_case_get_float", "ekh_case_init_state", "ekh_case_write_ini",
                "ekh_solver_registered", "ekh_run_ini", "ekh_solver_create", "ekh_solver_destroy", "ekh_solver_finished",
                "ekh_solver_next_iteration", "ekh_solver_advance", "ekh_solver_save_solution", "ekh_solver_info",
                "ekh_solver_state")


def _ov(overrides) -> bytes:
    if not overrides:
        return b""
    return "\n".join(f"{k}={v}" for k, v in overrides.items()).encode()


def settings(name: str) -> str:
    """Path of one of the shipped parameter files (euler_kokkos_b200/settings)."""
    return os.path.join(SETTINGS_DIR, name)


class Case:
    """An ini file (+ overrides 'section.name' -> value) seen through the C++ host layer."""

    def __init__(self, ini_path: str, overrides=None, rank: int = 0, nranks: int = 1):
        self._h = hlib.ekh_case_create(ini_path.encode(), _ov(overrides), rank, nranks)
        if not self._h:
            raise FileNotFoundError(ini_path)
        self.params = capi.EkParams()
        hlib.ekh_case_params(self._h, C.byref(self.params))
        nb = (C.c_int32 * 6)()
        hlib.ekh_case_neighbors(self._h, nb)
        self.neighbors = list(nb)
        info = (C.c_double * 16)()
        hlib.ekh_case_run_info(self._h, info)
        self.t_end, self.nstepmax, self.noutput, self.nlog = info[0], int(info[1]), int(info[2]), int(info[3])
        self.origin = (info[4], info[5], info[6])
        self.mpi_pos = (int(info[7]), int(info[8]), int(info[9]))
        self.mpi_dims = (int(info[10]), int(info[11]), int(info[12]))

    def __del__(self):
        try:
            if self._h:
                hlib.ekh_case_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def get_string(self, section, name) -> str:
        buf = C.create_string_buffer(512)
        hlib.ekh_case_get_string(self._h, section.encode(), name.encode(), buf, 512)
        return buf.value.decode()

    def get_float(self, section, name, default=0.0) -> float:
        return hlib.ekh_case_get_float(self._h, section.encode(), name.encode(), default)

    def init_state(self, U: np.ndarray, grav: np.ndarray | None = None):
        """Fill the ghosted SoA state (and the gravity field) with the ini's initial condition."""
        assert U.dtype == np.float64 and U.flags["C_CONTIGUOUS"] and U.shape == tuple(self.params.shape)
        U[...] = 0.0
        gp = None
        if grav is not None:
            assert grav.dtype == np.float64 and grav.flags["C_CONTIGUOUS"]
            grav[...] = 0.0
            gp = grav.ctypes.data_as(_vp)
        elif self.params.gravity_enabled:
            self._grav = np.zeros((self.params.dim,) + tuple(self.params.shape[1:]))
            gp = self._grav.ctypes.data_as(_vp)
        rc = hlib.ekh_case_init_state(self._h, U.ctypes.data_as(_vp), gp)
        if rc != 0:
            raise RuntimeError(f"initial condition '{self.get_string('hydro', 'problem')}' not available (rc={rc})")
        return U

    def init_gravity(self, grav: np.ndarray):
        tmp = np.zeros(self.params.shape)
        self.init_state(tmp, grav)
        return grav

    def new_state(self):
        U = np.zeros(self.params.shape)
        grav = np.zeros((self.params.dim,) + tuple(self.params.shape[1:])) if self.params.gravity_enabled else None
        self.init_state(U, grav)
        return U, grav

    def save_vti(self, U: np.ndarray, step: int) -> str:
        """Write dump `step` as the C++ driver does: <prefix>_%07d.vti, or for a decomposed case this rank's piece
        <prefix>_time%07d_mpi%05d.vti (+ the .pvti header from rank 0).  Returns the file written."""
        assert U.dtype == np.float64 and U.flags["C_CONTIGUOUS"] and U.shape == tuple(self.params.shape)
        buf = C.create_string_buffer(1024)
        rc = hlib.ekh_case_save_vti(self._h, U.ctypes.data_as(_vp), int(step), buf, 1024)
        if rc != 0:
            raise OSError(f"cannot write the .vti dump (rc={rc})")
        return buf.value.decode()

    def write_ini(self, path: str):
        rc = hlib.ekh_case_write_ini(self._h, path.encode())
        if rc != 0:
            raise OSError(f"cannot write {path}")

    def describe(self) -> str:
        p = self.params
        name = self.get_string("run", "solver_name")
        rs = {v: k for k, v in capi.RIEMANN.items()}[p.riemann]
        return f"{name}, riemann={rs}, impl={p.impl_version}, problem={self.get_string('hydro', 'problem')}"

    def kernel_name(self) -> str:
        if self.params.mhd:
            return ("mhd step kernels (mhd_pet_tma + mhd_faces_edges_call + mhd_update)" if self.params.dim == 3
                    else "mhd2d step kernels (prims + elec + trace + faces_edge + update)")
        return f"hydro_step_kernel<{self.params.dim}>"


class Solver:
    """SolverFactory::create(run/solver_name, ...) on an ini file: the object the reference's main() drives
    (src/main.cpp:112-153) -- finished(), next_iteration(), save_solution(), m_t / m_dt / m_iteration -- plus the
    device-resident advance(n) and a copy of the current state."""

    def __init__(self, ini_path: str, overrides=None):
        self._h = hlib.ekh_solver_create(ini_path.encode(), _ov(overrides))
        if not self._h:
            raise ValueError(f"{ini_path}: unreadable, or run/solver_name is not a registered solver")
        self.case = Case(ini_path, overrides)

    def close(self):
        if self._h:
            hlib.ekh_solver_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _info(self):
        out = (C.c_double * 8)()
        hlib.ekh_solver_info(self._h, out)
        return out

    t = property(lambda self: self._info()[0])
    dt = property(lambda self: self._info()[1])
    iteration = property(lambda self: int(self._info()[2]))
    times_saved = property(lambda self: int(self._info()[3]))

    def finished(self) -> bool:
        return bool(hlib.ekh_solver_finished(self._h))

    def next_iteration(self):
        hlib.ekh_solver_next_iteration(self._h)

    def advance(self, n: int) -> int:
        return hlib.ekh_solver_advance(self._h, int(n))

    def save_solution(self):
        hlib.ekh_solver_save_solution(self._h)

    def state(self) -> np.ndarray:
        U = np.zeros(self.case.params.shape)
        rc = hlib.ekh_solver_state(self._h, U.ctypes.data_as(_vp), U.size)
        if rc != 0:
            raise RuntimeError(f"ekh_solver_state failed (rc={rc})")
        return U


def write_ini(src_ini: str, overrides, dst_path: str):
    Case(src_ini, overrides).write_ini(dst_path)


def solver_registered(name: str) -> bool:
    return bool(hlib.ekh_solver_registered(name.encode()))


def run_ini(ini_path: str, overrides=None) -> int:
    """euler_b200's main() as a function (prints to stdout like the reference)."""
    return hlib.ekh_run_ini(ini_path.encode(), _ov(overrides))
