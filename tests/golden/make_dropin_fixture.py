#!/usr/bin/env python
"""Fixture for tests/test_dropin_gpu.py: what the UNMODIFIED reference binary (oracle/_ref) prints and the XML header
of the .vti files it writes, for four golden cases (hydro 2-D / 3-D, MHD 2-D / 3-D) run with [run] nlog=3:
  * stdout lines 'time step=...' and 'Output results at time ...' of a run with nOutput=-1 (one dump per step) and of a
    run with nOutput=0 (no output: the stretch our driver runs device-resident),
  * the bytes of the first dump up to and including '<AppendedData encoding="raw">\n _' (main.cpp:112-153,
    SolverBase.cpp:240-265, IO_VTK.cpp:213-410).
The states themselves are in the golden .npz files (make_golden.py).

    python tests/golden/make_dropin_fixture.py        # needs oracle/_ref
"""
from __future__ import annotations

import glob
import json
import os
import re
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle_lib as O  # noqa: E402
from tests.util import edit_ini, log_lines  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = ("implode2d_hllc", "blast3d_hllc", "ot2d_hlld", "ot3d_hlld", "rt3d_approx_grav_impl1")
NLOG = 3


def main():
    fix = {"nlog": NLOG, "cases": {}}
    for name in CASES:
        z = np.load(os.path.join(HERE, name + ".npz"))
        rec = {}
        for mode, nout in (("dumps", -1), ("quiet", 0)):
            with tempfile.TemporaryDirectory() as td:
                ini = edit_ini(str(z["ini"]), nlog=NLOG, noutput=nout, outputdir=td)
                path = os.path.join(td, "case.ini")
                with open(path, "w") as f:
                    f.write(ini)
                out = O.run_reference(path, td, threads=2)
                rec[mode + "_stdout"] = log_lines(out)
                if nout:
                    files = sorted(glob.glob(os.path.join(td, "g_*.vti")))
                    rec["nfiles"] = len(files)
                    with open(files[0], "rb") as f:
                        blob = f.read()
                    m = blob.find(b"<AppendedData")
                    end = blob.find(b"_", m) + 1
                    rec["vti_header"] = blob[:end].decode("ascii")
                    rec["vti_size"] = len(blob)
                    rec["vti_tail"] = blob[-64:].decode("ascii", errors="replace")
        fix["cases"][name] = rec
        print(name, rec["nfiles"], len(rec["dumps_stdout"]), len(rec["quiet_stdout"]))
    with open(os.path.join(HERE, "dropin_ref.json"), "w") as f:
        json.dump(fix, f, indent=1)


if __name__ == "__main__":
    if not O.have_reference():
        raise SystemExit("oracle/_ref/euler_kokkos_ref missing")
    main()
