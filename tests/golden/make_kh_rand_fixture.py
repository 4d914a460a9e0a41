#!/usr/bin/env python
"""tests/golden/kh_rand/ref.npz: dump 0 of the UNMODIFIED reference binary (oracle/_ref/euler_kokkos_ref, Kokkos OpenMP
back-end) on the Kelvin-Helmholtz problems with [KH] perturbation_rand=1, run with 1, 3 and 4 OpenMP threads.

The reference draws the perturbation from Kokkos::Random_XorShift64_Pool inside an MDRangePolicy loop: the stream
depends on the thread count and on the tile-to-thread map.  The host layer reproduces that stream
(csrc/host/InitConditionsMore.cpp, KokkosPoolStream); tests/test_oracle_golden.py compares it with these dumps.

    python tests/golden/make_kh_rand_fixture.py        # needs oracle/_ref (i.e. /root/reference at build time)
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
RAND = {"kh.perturbation_rand": 1, "kh.perturbation_sine": 0, "kh.perturbation_sine_robertson": 0}
# name -> (settings file, overrides, OpenMP threads)
CASES = {
    "kh2d_t1": (R + "test_muscl_kelvin_helmholtz_2D.ini", {"mesh.nx": 21, "mesh.ny": 14}, 1),
    "kh2d_t3": (R + "test_muscl_kelvin_helmholtz_2D.ini", {"mesh.nx": 21, "mesh.ny": 14}, 3),
    "kh2d_t4_seed7": (R + "test_muscl_kelvin_helmholtz_2D.ini", {"mesh.nx": 32, "mesh.ny": 16, "kh.rand_seed": 7}, 4),
    "kh3d_t1": (R + "test_muscl_kelvin_helmholtz_3D.ini", {"mesh.nx": 9, "mesh.ny": 7, "mesh.nz": 6}, 1),
    "kh3d_t3": (R + "test_muscl_kelvin_helmholtz_3D.ini", {"mesh.nx": 9, "mesh.ny": 7, "mesh.nz": 6}, 3),
    "kh3d_t4": (R + "test_muscl_kelvin_helmholtz_3D.ini", {"mesh.nx": 12, "mesh.ny": 8, "mesh.nz": 8}, 4),
    "mhdkh2d_t1": (R + "test_muscl_kelvin_helmholtz_2D_mhd.ini", {"mesh.nx": 20, "mesh.ny": 12}, 1),
    "mhdkh2d_t3": (R + "test_muscl_kelvin_helmholtz_2D_mhd.ini", {"mesh.nx": 20, "mesh.ny": 12}, 3),
}


def main():
    if not O.have_reference():
        raise SystemExit("oracle/_ref/euler_kokkos_ref missing: run `make -C oracle/ref_build` where /root/reference exists")
    data = {}
    for name, (ini_name, ov, threads) in CASES.items():
        with tempfile.TemporaryDirectory() as td:
            overrides = dict(ov)
            overrides.update(RAND)
            overrides.update({"run.noutput": -1, "run.nstepmax": 1, "run.tend": 1e6, "run.nlog": 1, "output.outputdir": td,
                              "output.outputprefix": "g", "output.hdf5_enabled": "no", "output.vtk_enabled": "yes",
                              "output.outputvtkascii": "false"})
            ini = os.path.join(td, "case.ini")
            hostapi.write_ini(ini_name, overrides, ini)
            with open(ini) as f:
                ini_text = f.read()
            O.run_reference(ini, td, threads=threads)
            files = sorted(glob.glob(os.path.join(td, "g_*.vti")))
            names, u0 = read_vti(files[0])
        data[name + "/ini"] = np.array(ini_text)
        data[name + "/threads"] = np.array(threads)
        data[name + "/u0"] = u0
        print(name, names, u0.shape, "threads", threads)
    os.makedirs(os.path.join(HERE, "kh_rand"), exist_ok=True)
    np.savez_compressed(os.path.join(HERE, "kh_rand", "ref.npz"), **data)


if __name__ == "__main__":
    main()
