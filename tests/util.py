"""Shared helpers of the test-suite: parameter construction with the reference's float-parsing
semantics, synthetic states, and the parity metric of SURVEY.md section 8(c)."""
from __future__ import annotations

import re

import numpy as np

from euler_kokkos_b200 import capi


def f32(x) -> float:
    """ConfigMap::getFloat semantics: strtof then widened to double (ConfigMap.cpp:38-46)."""
    return float(np.float32(x))


def make_params(dim=3, mhd=False, n=(16, 12, 8), box=((0.0, 1.0), (0.0, 1.0), (0.0, 1.0)), riemann="hllc",
                bc=(1, 1, 1, 1, 1, 1), gamma0=1.666, cfl=0.8, slope_type=2.0, smallr=1e-10, smallc=1e-10,
                impl=0, gravity=False, hancock=False) -> capi.EkParams:
    p = capi.EkParams()
    p.dim, p.mhd = dim, int(mhd)
    p.nbvar = 8 if mhd else (5 if dim == 3 else 4)
    p.ghost = 3 if mhd else 2
    p.nx, p.ny, p.nz = n[0], n[1], (n[2] if dim == 3 else 1)
    p.riemann = capi.RIEMANN[riemann]
    p.impl_version = impl
    for i in range(6):
        p.bc[i] = bc[i]
    p.gravity_enabled, p.gravity_hancock = int(gravity), int(hancock)
    p.dx = (f32(box[0][1]) - f32(box[0][0])) / n[0]
    p.dy = (f32(box[1][1]) - f32(box[1][0])) / n[1]
    p.dz = (f32(box[2][1]) - f32(box[2][0])) / (n[2] if dim == 3 else 1)
    p.gamma0, p.cfl, p.slope_type = f32(gamma0), f32(cfl), f32(slope_type)
    p.smallr, p.smallc = f32(smallr), f32(smallc)
    p.smallp = p.smallc * p.smallc / p.gamma0
    p.smallpp = p.smallr * p.smallp
    p.gamma6 = (p.gamma0 + 1.0) / (2.0 * p.gamma0)
    return p


def random_hydro_state(p, seed=0, mach=0.8, contrast=4.0):
    """Smooth-ish random conserved state with shocks-to-be: positive density/pressure, O(1) velocities."""
    rng = np.random.default_rng(seed)
    nv, ks, js, is_ = p.shape
    dim = p.dim

    def field(lo, hi):
        a = rng.uniform(lo, hi, size=(ks, js, is_))
        # a little smoothing so that slopes are not all limited to zero
        for ax in range(3):
            if a.shape[ax] > 2:
                a = 0.5 * a + 0.25 * (np.roll(a, 1, ax) + np.roll(a, -1, ax))
        return a

    rho = field(1.0, contrast)
    pr = field(0.5, contrast)
    vel = [field(-mach, mach) for _ in range(dim)]
    U = np.zeros(p.shape)
    U[0] = rho
    ek = 0.0
    for c in range(dim):
        U[2 + c] = rho * vel[c]
        ek = ek + 0.5 * rho * vel[c] ** 2
    U[1] = pr / (p.gamma0 - 1.0) + ek
    if p.mhd:
        for c in range(3):
            U[5 + c] = field(-0.7, 0.7)
        # total energy includes the magnetic energy of the cell-centred field (approximately)
        U[1] += 0.5 * (U[5] ** 2 + U[6] ** 2 + U[7] ** 2)
    return np.ascontiguousarray(U)


def blast_state(p, radius_cells=4.0):
    """Spherical blast (InitBlastFunctor semantics, not bit-identical: synthetic test input)."""
    nv, ks, js, is_ = p.shape
    g = p.ghost
    k, j, i = np.meshgrid(np.arange(ks), np.arange(js), np.arange(is_), indexing="ij")
    cx, cy, cz = (is_ - 1) / 2.0, (js - 1) / 2.0, (ks - 1) / 2.0
    r2 = (i - cx) ** 2 + (j - cy) ** 2 + ((k - cz) ** 2 if p.dim == 3 else 0.0)
    inside = r2 < radius_cells**2
    U = np.zeros(p.shape)
    U[0] = np.where(inside, 1.0, 1.2)
    U[1] = np.where(inside, 10.0, 0.1) / (p.gamma0 - 1.0)
    return np.ascontiguousarray(U)


def group_rel_linf(a, b, p):
    """max|a-b| / max|b| per physical group (rho; E; |rho v| over all components; |B|) over interior
    cells -- SURVEY.md 8(c): several components are identically zero for whole runs."""
    from euler_kokkos_b200.vti import interior_of

    ai, bi = interior_of(a, p.ghost, p.dim), interior_of(b, p.ghost, p.dim)
    nm = 2 + p.dim if not p.mhd else 5
    groups = {"rho": [0], "E": [1], "mom": list(range(2, nm))}
    if p.mhd:
        groups["B"] = [5, 6, 7]
    out = {}
    for name, idx in groups.items():
        den = max(np.abs(bi[idx]).max(), 1e-300)
        out[name] = float(np.abs(ai[idx] - bi[idx]).max() / den)
    return out


def edit_ini(text: str, **kv) -> str:
    """replace 'key=value' lines (keys are unique across sections in these files)"""
    for k, v in kv.items():
        text, n = re.subn(rf"(?m)^{k}=.*$", f"{k}={v}", text)
        assert n == 1, (k, n)
    return text


def log_lines(out: str):
    return [ln.rstrip() for ln in out.splitlines() if ln.startswith("time step=") or ln.startswith("Output results at time")
            or ln.startswith("final time is")]
