"""ctypes binding of the C ABI declared in include/ek_b200.h (libek_b200.so).

Fails loudly (ImportError) when the CUDA library is missing: there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("EK_B200_LIB") or os.path.join(_HERE, "lib", "libek_b200.so")  # env override: kernel-tuning builds

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "or `make -C euler_kokkos_b200/csrc`. There is no CPU fallback."
    )

lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)

EK_OK, EK_ERR_INVALID, EK_ERR_CUDA, EK_ERR_NCCL, EK_ERR_STATE = 0, 1, 2, 3, 4
RIEMANN = {"approx": 0, "llf": 1, "hll": 2, "hllc": 3, "hlld": 4}
BC_DIRICHLET, BC_NEUMANN, BC_PERIODIC, BC_COPY = 1, 2, 3, 4
COMM_ID_BYTES = 128


class EkParams(C.Structure):
    """struct ek_params (include/ek_b200.h)."""

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

    def copy(self) -> "EkParams":
        p = EkParams()
        C.memmove(C.byref(p), C.byref(self), C.sizeof(EkParams))
        return p

    @property
    def shape(self):
        """(nbvar, ksize, jsize, isize) of the ghosted SoA state."""
        g = self.ghost
        ks = self.nz + 2 * g if self.dim == 3 else 1
        return (self.nbvar, ks, self.ny + 2 * g, self.nx + 2 * g)


_dp = C.POINTER(C.c_double)
_vp = C.c_void_p

_protos = {
    "ek_abi_version": (C.c_int, []),
    "ek_create": (C.c_int, [C.POINTER(EkParams), C.c_int, C.POINTER(_vp)]),
    "ek_destroy": (C.c_int, [_vp]),
    "ek_last_error": (C.c_char_p, [_vp]),
    "ek_set_stream": (C.c_int, [_vp, _vp]),
    "ek_sync": (C.c_int, [_vp]),
    "ek_state_len": (C.c_int64, [_vp]),
    "ek_upload_state": (C.c_int, [_vp, _vp]),
    "ek_download_state": (C.c_int, [_vp, _vp]),
    "ek_download_state_async": (C.c_int, [_vp, _vp]),
    "ek_download_wait": (C.c_int, [_vp]),
    "ek_alloc_host": (C.c_int, [C.c_size_t, C.POINTER(_vp)]),
    "ek_free_host": (C.c_int, [_vp]),
    "ek_device_state": (C.c_int, [_vp, C.POINTER(_vp)]),
    "ek_set_gravity": (C.c_int, [_vp, _vp]),
    "ek_fill_boundaries": (C.c_int, [_vp]),
    "ek_compute_dt": (C.c_int, [_vp, _dp]),
    "ek_step": (C.c_int, [_vp, C.c_double]),
    "ek_run": (C.c_int, [_vp, C.c_int, C.c_double, C.c_double, _dp, _dp, C.POINTER(C.c_int)]),
    "ek_conserved_sums": (C.c_int, [_vp, _dp]),
    "ek_max_divb": (C.c_int, [_vp, _dp]),
    "ek_launch_count": (C.c_int64, [_vp]),
    "ek_kernel_timing": (C.c_int, [_vp, C.c_int]),
    "ek_kernel_time": (C.c_int, [_vp, _dp, C.POINTER(C.c_int)]),
    "ek_selftest_math": (C.c_int, [C.c_int, _vp, C.c_int, _vp]),
    "ek_comm_unique_id": (C.c_int, [_vp]),
    "ek_comm_attach": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.POINTER(C.c_int32)]),
    "ek_comm_detach": (C.c_int, [_vp]),
}
ABI_SYMBOLS = tuple(_protos)
for _name, (_res, _args) in _protos.items():
    _f = getattr(lib, _name)  # AttributeError here == the library does not export a declared symbol
    _f.restype = _res
    _f.argtypes = _args


class EkError(RuntimeError):
    pass


def _ptr(a):
    """host pointer of a numpy array / raw int address (e.g. pinned torch tensor .data_ptr())."""
    if isinstance(a, np.ndarray):
        assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
        return a.ctypes.data_as(_vp)
    return _vp(int(a))


class Context:
    """One sub-domain on one GPU (opaque ek_ctx*)."""

    def __init__(self, params: EkParams, device: int = 0):
        self.params = params.copy()
        self._h = _vp()
        rc = lib.ek_create(C.byref(self.params), device, C.byref(self._h))
        if rc != EK_OK:
            raise EkError(f"ek_create failed ({rc}): {lib.ek_last_error(None).decode()}")

    def _ck(self, rc):
        if rc != EK_OK:
            raise EkError(f"libek_b200 error {rc}: {lib.ek_last_error(self._h).decode()}")

    def close(self):
        if self._h:
            lib.ek_destroy(self._h)
            self._h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def shape(self):
        return self.params.shape

    def set_stream(self, cuda_stream: int):
        self._ck(lib.ek_set_stream(self._h, _vp(cuda_stream)))

    def sync(self):
        self._ck(lib.ek_sync(self._h))

    def upload(self, U):
        self._ck(lib.ek_upload_state(self._h, _ptr(U)))

    def download_async(self, out):
        """snapshot the current level and start its copy to `out` (pinned memory overlaps with later steps)"""
        self._ck(lib.ek_download_state_async(self._h, _ptr(out)))

    def download_wait(self):
        self._ck(lib.ek_download_wait(self._h))

    def download(self, out=None):
        if out is None:
            out = np.empty(self.shape, dtype=np.float64)
        self._ck(lib.ek_download_state(self._h, _ptr(out)))
        return out

    def set_gravity(self, g):
        self._ck(lib.ek_set_gravity(self._h, _ptr(g)))

    def fill_boundaries(self):
        self._ck(lib.ek_fill_boundaries(self._h))

    def compute_dt(self) -> float:
        dt = C.c_double()
        self._ck(lib.ek_compute_dt(self._h, C.byref(dt)))
        return dt.value

    def step(self, dt: float):
        self._ck(lib.ek_step(self._h, float(dt)))

    def run(self, nsteps: int, t: float = 0.0, t_end: float = 1e300):
        t_out, dt_out, n = C.c_double(), C.c_double(), C.c_int()
        self._ck(lib.ek_run(self._h, int(nsteps), float(t), float(t_end), C.byref(t_out), C.byref(dt_out), C.byref(n)))
        return t_out.value, dt_out.value, n.value

    def conserved_sums(self):
        out = np.zeros(self.params.nbvar)
        self._ck(lib.ek_conserved_sums(self._h, out.ctypes.data_as(_dp)))
        return out

    def max_divb(self) -> float:
        v = C.c_double()
        self._ck(lib.ek_max_divb(self._h, C.byref(v)))
        return v.value

    @property
    def launch_count(self) -> int:
        return int(lib.ek_launch_count(self._h))

    def kernel_timing(self, enable: bool):
        self._ck(lib.ek_kernel_timing(self._h, int(enable)))

    def kernel_time(self):
        """(summed ms, launches) of the fused step kernel since the last call; resets."""
        ms, n = C.c_double(), C.c_int()
        self._ck(lib.ek_kernel_time(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def device_state_ptr(self) -> int:
        p = _vp()
        self._ck(lib.ek_device_state(self._h, C.byref(p)))
        return p.value

    def comm_attach(self, uid: bytes, rank: int, nranks: int, neighbors):
        nb = (C.c_int32 * 6)(*neighbors)
        buf = C.create_string_buffer(uid, COMM_ID_BYTES)
        self._ck(lib.ek_comm_attach(self._h, C.cast(buf, _vp), rank, nranks, nb))


def comm_unique_id() -> bytes:
    buf = C.create_string_buffer(COMM_ID_BYTES)
    rc = lib.ek_comm_unique_id(C.cast(buf, _vp))
    if rc != EK_OK:
        raise EkError(f"ek_comm_unique_id failed ({rc})")
    return buf.raw


def selftest_math(x: np.ndarray, device: int = 0) -> dict:
    """device values of max(x,0), min(x,0), 1/x, sqrt(x), 1/sqrt(x), (1+x)/x for the operands x (ek_selftest_math)"""
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty((6, x.size))
    rc = lib.ek_selftest_math(device, _ptr(x), x.size, _ptr(out))
    if rc != EK_OK:
        raise EkError(f"ek_selftest_math failed ({rc}): {lib.ek_last_error(None).decode()}")
    return dict(zip(("max0", "min0", "rcp", "sqrt", "rsqrt", "div"), out))
