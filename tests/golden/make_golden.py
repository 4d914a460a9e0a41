#!/usr/bin/env python
"""Generate the golden fixtures of tests/golden/*.npz by RUNNING THE UNMODIFIED REFERENCE BINARY
(oracle/_ref/euler_kokkos_ref, built from /root/reference by oracle/ref_build/Makefile) on small
versions of the BASELINE.json configs with [run] nOutput=-1 (one .vti dump per step,
SolverBase.cpp:240-265) and reading its full-precision Float64 dumps back.

The reference has no golden vectors of its own; these fixtures are the pin of
the CPU oracle (tests/test_oracle_golden.py) and, through it, of the CUDA path.

    python tests/golden/make_golden.py            # needs oracle/_ref (i.e. /root/reference at build time)

Each fixture holds: the ini text actually run, the interior state of the initial condition and of
selected later steps, shape (nvar, nz, ny, nx).
"""
from __future__ import annotations

import glob
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

R = "/root/reference/settings/"
# name -> (settings file, overrides, number of steps, steps kept)
CASES = {
    "implode2d_hllc": ("test_implode_2D.ini", {"mesh.nx": 48, "mesh.ny": 24}, 8, (1, 2, 8)),
    "blast3d_hllc": ("test_blast_3D.ini", {"mesh.nx": 24, "mesh.ny": 18, "mesh.nz": 12}, 6, (1, 2, 6)),
    "implode3d_hllc": ("test_implode_3D.ini", {"mesh.nx": 20, "mesh.ny": 16, "mesh.nz": 12}, 4, (1, 4)),
    "rt3d_approx_grav_impl1": ("test_rayleigh_taylor_3d.ini", {"mesh.nx": 12, "mesh.ny": 10, "mesh.nz": 24}, 5, (1, 2, 5)),
    "ot3d_hlld": ("test_mhd_orszag_tang_3D.ini", {"mesh.nx": 20, "mesh.ny": 16, "mesh.nz": 8}, 6, (1, 2, 6)),
    "ot3d_hlld_slope1": ("test_mhd_orszag_tang_3D.ini", {"mesh.nx": 16, "mesh.ny": 16, "mesh.nz": 8, "hydro.slope_type": 1}, 3, (1, 3)),
    "ot3d_hll": ("test_mhd_orszag_tang_3D.ini", {"mesh.nx": 16, "mesh.ny": 12, "mesh.nz": 8, "hydro.riemann": "hll"}, 3, (1, 3)),
    # non-smooth MHD (shock branches of HLLD / 2-D HLLD) with periodic, reflective and outflow faces; run from the
    # reference's own ini (its [blast] bz values are uninitialised memory in the reference, BlastParams.h:56-62,
    # so the test starts from the reference's dump 0 rather than from our initial condition)
    "mhdblast3d_hlld": ("/root/reference/settings/test_mhd_blast_3D.ini",
                        {"mesh.nx": 16, "mesh.ny": 14, "mesh.nz": 12, "blast.radius": 0.3, "other.implementationversion": 0}, 8, (1, 4, 8)),
    "mhdblast3d_hlld_bc": ("/root/reference/settings/test_mhd_blast_3D.ini",
                           {"mesh.nx": 14, "mesh.ny": 12, "mesh.nz": 12, "blast.radius": 0.3, "other.implementationversion": 0,
                            "mesh.boundary_type_xmin": 1, "mesh.boundary_type_xmax": 2, "mesh.boundary_type_ymin": 2,
                            "mesh.boundary_type_ymax": 1, "mesh.boundary_type_zmin": 1, "mesh.boundary_type_zmax": 1}, 8, (1, 4, 8)),
    "blast3d_hllc_big": ("test_blast_3D.ini", {"mesh.nx": 20, "mesh.ny": 16, "mesh.nz": 12, "blast.radius": 0.35,
                                               "mesh.boundary_type_xmin": 2, "mesh.boundary_type_ymax": 3, "mesh.boundary_type_ymin": 3}, 8, (1, 8)),
    "ot2d_hlld": ("test_mhd_orszag_tang_2D.ini", {"mesh.nx": 48, "mesh.ny": 40}, 10, (1, 2, 10)),
    "ot2d_hlld_slope1": ("test_mhd_orszag_tang_2D.ini", {"mesh.nx": 32, "mesh.ny": 32, "hydro.slope_type": 1}, 4, (1, 4)),
    "ot2d_hll": ("test_mhd_orszag_tang_2D.ini", {"mesh.nx": 32, "mesh.ny": 24, "hydro.riemann": "hll"}, 3, (1, 3)),
    "mhdblast2d_hlld_bc": ("/root/reference/settings/test_mhd_blast_2D.ini",
                           {"mesh.nx": 32, "mesh.ny": 28, "blast.radius": 0.3, "other.implementationversion": 0, "run.solver_name": "MHD_Muscl_2D",
                            "mesh.boundary_type_xmin": 1, "mesh.boundary_type_xmax": 2, "mesh.boundary_type_ymin": 3,
                            "mesh.boundary_type_ymax": 3}, 10, (1, 5, 10)),
    "ot3d_llf": ("test_mhd_orszag_tang_3D.ini", {"mesh.nx": 16, "mesh.ny": 12, "mesh.nz": 8, "hydro.riemann": "llf"}, 3, (1, 3)),
    # ---- the remaining problems of the reference, from its own settings files ----------------
    "fourquad2d_hllc": (R + "test_four_quadrant_2D.ini", {"mesh.nx": 40, "mesh.ny": 36}, 6, (1, 6)),
    "fourquad3d_hllc": (R + "test_four_quadrant_3D.ini", {"mesh.nx": 16, "mesh.ny": 14, "mesh.nz": 12}, 4, (1, 4)),
    "gresho2d_hllc": (R + "test_muscl_gresho_vortex_2D.ini", {"mesh.nx": 32, "mesh.ny": 32}, 4, (1, 4)),
    "isentropic2d_hllc": (R + "test_muscl_isentropic_vortex_2d.ini", {"mesh.nx": 32, "mesh.ny": 28}, 4, (1, 4)),
    "kh2d_hllc": (R + "test_muscl_kelvin_helmholtz_2D.ini", {"mesh.nx": 32, "mesh.ny": 32}, 4, (1, 4)),
    "kh3d_hllc": (R + "test_muscl_kelvin_helmholtz_3D.ini", {"mesh.nx": 16, "mesh.ny": 16, "mesh.nz": 16}, 3, (1, 3)),
    "mhdkh2d_hlld": (R + "test_muscl_kelvin_helmholtz_2D_mhd.ini", {"mesh.nx": 32, "mesh.ny": 32}, 4, (1, 4)),
    "briowu2d_x_hlld": (R + "test_mhd_brio_wu_2d_x.ini", {"mesh.nx": 64, "mesh.ny": 8, "other.implementationversion": 0}, 8, (1, 8)),
    "briowu2d_y_hlld": (R + "test_mhd_brio_wu_2d_y.ini", {"mesh.nx": 8, "mesh.ny": 48, "other.implementationversion": 0}, 6, (1, 6)),
    "briowu3d_z_hlld": (R + "test_mhd_brio_wu_3d_z.ini", {"mesh.nx": 8, "mesh.ny": 8, "mesh.nz": 48, "other.implementationversion": 0}, 6, (1, 6)),
    # MHD implementationVersion 1 / 2 exactly as the shipped files ask for it (no override of [other]): served by the
    # version-0 kernels (the reference's versions agree to 2e-15 in these configurations)
    "mhdblast3d_hlld_impl1": (R + "test_mhd_blast_3D.ini", {"mesh.nx": 16, "mesh.ny": 12, "mesh.nz": 12, "blast.radius": 0.3}, 6, (1, 6)),
    "briowu2d_x_hlld_impl2": (R + "test_mhd_brio_wu_2d_x.ini", {"mesh.nx": 64, "mesh.ny": 8}, 8, (1, 8)),
    "briowu2d_y_hlld_impl2": (R + "test_mhd_brio_wu_2d_y.ini", {"mesh.nx": 8, "mesh.ny": 48}, 6, (1, 6)),
    "rotor2d_hlld": (R + "test_mhd_rotor.ini", {"mesh.nx": 40, "mesh.ny": 40}, 5, (1, 5)),
    "fieldloop2d_hlld": (R + "mhd_fieldloop2d.ini", {"mesh.nx": 32, "mesh.ny": 20, "hydro.problem": "field_loop"}, 5, (1, 5)),
    "fieldloop3d_hlld": (R + "mhd_fieldloop3d.ini", {"mesh.nx": 24, "mesh.ny": 12, "mesh.nz": 12, "hydro.problem": "field_loop"}, 4, (1, 4)),
}


def make(name):
    ini_name, ov, nsteps, keep = CASES[name]
    with tempfile.TemporaryDirectory() as td:
        overrides = dict(ov)
        overrides.update({"run.noutput": -1, "run.nstepmax": nsteps, "run.tend": 1e6, "run.nlog": 1,
                          "output.outputdir": td, "output.outputprefix": "g", "output.hdf5_enabled": "no",
                          "output.vtk_enabled": "yes", "output.outputvtkascii": "false"})
        ini = os.path.join(td, "case.ini")
        hostapi.write_ini(ini_name if os.path.isabs(ini_name) else hostapi.settings(ini_name), overrides, ini)
        with open(ini) as f:
            ini_text = f.read()
        out = O.run_reference(ini, td, threads=4)
        files = sorted(glob.glob(os.path.join(td, "g_*.vti")))
        # file 0 = initial state (main.cpp:117), file 1 = same state (iteration 0), file n+1 = after n steps
        assert len(files) >= nsteps + 2, (name, len(files), out[-2000:])
        names, u0 = read_vti(files[0])
        _, u1 = read_vti(files[1])
        assert np.array_equal(u0, u1)
        data = {"ini": np.array(ini_text), "names": np.array(names), "steps": np.array(keep), "u0": u0}
        for s in keep:
            _, us = read_vti(files[s + 1])
            data[f"u{s}"] = us
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **data)
    print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB, vars={names}, shape={u0.shape}")


if __name__ == "__main__":
    if not O.have_reference():
        raise SystemExit("oracle/_ref/euler_kokkos_ref missing: run `make -C oracle/ref_build` where /root/reference exists")
    for n in (sys.argv[1:] or CASES):
        make(n)
