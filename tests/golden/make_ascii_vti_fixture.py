#!/usr/bin/env python
"""tests/golden/ascii_vti/*.vti + cases.json: dump 0 of the UNMODIFIED reference binary (oracle/_ref/euler_kokkos_ref) with
[output] outputVtkAscii=true on small 2-D and 3-D problems (hydro and MHD).  The host layer's writer
(csrc/host/IO_VTK.cpp::save_vti) must reproduce these files byte for byte from its own initial condition
(tests/test_abi_and_host.py::test_ascii_vti_dump_is_byte_identical_to_the_reference).

    python tests/golden/make_ascii_vti_fixture.py        # needs oracle/_ref (i.e. /root/reference at build time)
"""
from __future__ import annotations

import glob
import json
import os
import shutil
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from euler_kokkos_b200 import hostapi  # noqa: E402
from oracle import oracle_lib as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
R = "/root/reference/settings/"
CASES = {
    "implode2d": ("test_implode_2D.ini", {"mesh.nx": 12, "mesh.ny": 7}),
    "blast3d": ("test_blast_3D.ini", {"mesh.nx": 7, "mesh.ny": 5, "mesh.nz": 4}),
    "ot2d": (R + "test_mhd_orszag_tang_2D.ini", {"mesh.nx": 9, "mesh.ny": 6}),
    "ot3d": ("test_mhd_orszag_tang_3D.ini", {"mesh.nx": 6, "mesh.ny": 5, "mesh.nz": 3}),
}


def main():
    if not O.have_reference():
        raise SystemExit("oracle/_ref/euler_kokkos_ref missing: run `make -C oracle/ref_build` where /root/reference exists")
    out_dir = os.path.join(HERE, "ascii_vti")
    os.makedirs(out_dir, exist_ok=True)
    meta = {}
    for name, (ini_name, ov) in CASES.items():
        with tempfile.TemporaryDirectory() as td:
            overrides = dict(ov)
            overrides.update({"run.noutput": -1, "run.nstepmax": 1, "run.tend": 1e6, "run.nlog": 1, "output.outputdir": td,
                              "output.outputprefix": "g", "output.hdf5_enabled": "no", "output.vtk_enabled": "yes",
                              "output.outputvtkascii": "true"})
            ini = os.path.join(td, "case.ini")
            hostapi.write_ini(ini_name if os.path.isabs(ini_name) else hostapi.settings(ini_name), overrides, ini)
            with open(ini) as f:
                ini_text = f.read()
            O.run_reference(ini, td, threads=2)
            files = sorted(glob.glob(os.path.join(td, "g_*.vti")))
            shutil.copy(files[0], os.path.join(out_dir, name + ".vti"))
        # the fixture's ini: same problem, output directory left to the test
        meta[name] = {"ini": ini_text.replace(td, "@OUTDIR@")}
        print(name, os.path.getsize(os.path.join(out_dir, name + ".vti")), "bytes")
    with open(os.path.join(out_dir, "cases.json"), "w") as f:
        json.dump(meta, f, indent=1)


if __name__ == "__main__":
    main()
