#!/usr/bin/env python
"""Fixture for tests/test_shipped_inis_gpu.py: the UNMODIFIED reference binary (oracle/_ref) on EVERY parameter file the
reference ships (settings/*.ini), sizes, solver, Riemann solver and implementationVersion AS SHIPPED, 5 steps; of the final
dump only a fingerprint is kept (per variable: sum, sum of squares, min, max and 64 values at fixed strided positions), so
that the fixture stays a few KB per file.

    python tests/golden/make_shipped_fixture.py        # needs oracle/_ref and /root/reference
"""
from __future__ import annotations

import glob
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from euler_kokkos_b200 import hostapi  # noqa: E402
from euler_kokkos_b200.vti import read_vti  # noqa: E402
from oracle import oracle_lib as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SETTINGS = "/root/reference/settings"
NSTEPS = 5
OVERRIDES = {"run.noutput": -1, "run.nstepmax": NSTEPS, "run.tend": 1e6, "run.nlog": 1000, "output.outputprefix": "g",
             "output.hdf5_enabled": "no", "output.vtk_enabled": "yes", "output.outputvtkascii": "false", "run.restart_enabled": 0}


def fingerprint(u: np.ndarray) -> dict:
    """u: (nvar, nz, ny, nx)"""
    out = {"shape": list(u.shape), "vars": []}
    for v in range(u.shape[0]):
        a = u[v].ravel()
        idx = np.linspace(0, a.size - 1, 64).astype(np.int64)
        out["vars"].append({"sum": float(a.sum(dtype=np.float64)), "sumsq": float((a * a).sum(dtype=np.float64)), "min": float(a.min()),
                            "max": float(a.max()), "sample": [float(x) for x in a[idx]]})
    return out


def run_one(binary_runner, ini_path, extra=None):
    with tempfile.TemporaryDirectory() as td:
        ov = dict(OVERRIDES)
        ov["output.outputdir"] = td
        ov.update(extra or {})
        ini = os.path.join(td, "case.ini")
        hostapi.write_ini(ini_path, ov, ini)
        binary_runner(ini, td)
        files = sorted(glob.glob(os.path.join(td, "g_*.vti")))
        assert len(files) == NSTEPS + 2, (ini_path, len(files))
        names, u = read_vti(files[-1])
        return names, u


def main():
    fix = {"nsteps": NSTEPS, "files": {}}
    for path in sorted(glob.glob(os.path.join(REF_SETTINGS, "*.ini"))):
        name = os.path.basename(path)
        names, u = run_one(lambda ini, td: O.run_reference(ini, td, threads=8), path)
        fix["files"][name] = {"names": list(names), "final": fingerprint(u)}
        print(name, u.shape)
    with open(os.path.join(HERE, "shipped_ref.json"), "w") as f:
        json.dump(fix, f)


if __name__ == "__main__":
    if not O.have_reference():
        raise SystemExit("oracle/_ref/euler_kokkos_ref missing")
    main()
