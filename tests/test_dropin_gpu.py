"""GPU tests of the DROP-IN itself: the C++ driver behind the reference's SolverFactory / ini API
(csrc/host: main.cpp -> SolverFactory::create -> SolverMusclB200 -> include/ek_b200.h), run exactly as a user of the
reference runs it -- `euler_b200 file.ini` -- and through the factory object (hostapi.Solver).

Checked against what the UNMODIFIED reference binary did on the same ini files (tests/golden/*.npz states,
tests/golden/dropin_ref.json stdout lines and .vti headers; generators committed beside them):
  * every .vti dump of a run with nOutput=-1 equals the reference's dump (< 1e-11, relative L-inf per physical group),
  * the .vti files have the reference's header byte for byte, the same size and the same trailer,
  * stdout carries the reference's 'time step=' / 'Output results at time' / 'final time is' lines at the nlog cadence,
  * the device-resident stretch (sync_free=yes, main.cpp) ends in the same state as the per-step host loop
    (sync_free=no, the reference's own loop, src/main.cpp:126-131).
"""
import glob
import json
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

from tests import util
from tests.util import edit_ini, log_lines

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
BIN = os.path.join(ROOT, "euler_kokkos_b200", "bin", "euler_b200")
with open(os.path.join(HERE, "golden", "dropin_ref.json")) as _f:
    FIX = json.load(_f)
CASES = sorted(FIX["cases"])
_NUM = re.compile(r"[-+]?\d+\.?\d*(?:[eE][-+]?\d+)?")


def _same_log(ours, ref):
    """same lines, same integers, floats equal to the printed precision (the last printed digit may round differently
    when the GPU's dt differs from the reference's in the 13th digit)"""
    assert len(ours) == len(ref), (ours, ref)
    for a, b in zip(ours, ref):
        assert _NUM.sub("#", a) == _NUM.sub("#", b), (a, b)
        for x, y in zip(_NUM.findall(a), _NUM.findall(b)):
            assert abs(float(x) - float(y)) <= 2e-6 * max(abs(float(y)), 1e-3), (a, b)


def _run_binary(ini_text, td, **kv):
    ini = os.path.join(td, "case.ini")
    with open(ini, "w") as f:
        f.write(edit_ini(ini_text, outputdir=td, **kv))
    r = subprocess.run([BIN, ini], cwd=td, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return r.stdout


@pytest.mark.parametrize("name", CASES)
def test_euler_b200_executable_reproduces_the_reference_run(name):
    from euler_kokkos_b200 import hostapi
    from euler_kokkos_b200.vti import embed_interior, read_vti

    z = np.load(os.path.join(HERE, "golden", name + ".npz"))
    ref = FIX["cases"][name]
    with tempfile.TemporaryDirectory() as td:
        out = _run_binary(str(z["ini"]), td, nlog=FIX["nlog"], noutput=-1)
        _same_log(log_lines(out), ref["dumps_stdout"])
        files = sorted(glob.glob(os.path.join(td, "g_*.vti")))
        assert len(files) == ref["nfiles"]
        with open(files[0], "rb") as f:
            blob = f.read()
        end = blob.find(b"_", blob.find(b"<AppendedData")) + 1
        assert blob[:end].decode("ascii") == ref["vti_header"]
        assert len(blob) == ref["vti_size"]
        assert blob[-64:].decode("ascii", errors="replace") == ref["vti_tail"]
        case = hostapi.Case(os.path.join(td, "case.ini"))
        p = case.params
        names, u0 = read_vti(files[0])
        assert list(names) == [str(n) for n in z["names"]]
        assert np.abs(u0 - z["u0"]).max() <= 1e-15 * np.abs(z["u0"]).max()
        for s in [int(x) for x in z["steps"]]:
            _, us = read_vti(files[s + 1])  # file 0 = initial state, file n+1 = after n steps (main.cpp:117)
            err = util.group_rel_linf(embed_interior(us, p.ghost, p.dim), embed_interior(z[f"u{s}"], p.ghost, p.dim), p)
            assert max(err.values()) < 1e-11, (name, s, err)


@pytest.mark.parametrize("name", CASES)
def test_device_resident_stretches_print_the_reference_log(name):
    """nOutput=0: main.cpp runs stretches of nlog iterations device-resident (ek_run); the log lines are the reference's"""
    z = np.load(os.path.join(HERE, "golden", name + ".npz"))
    ref = FIX["cases"][name]
    with tempfile.TemporaryDirectory() as td:
        out = _run_binary(str(z["ini"]), td, nlog=FIX["nlog"], noutput=0)
        _same_log(log_lines(out), ref["quiet_stdout"])
        assert not glob.glob(os.path.join(td, "*.vti"))
        assert "Perf" in out and "Mcell-updates/s" in out  # solver_utils.h:18-51


@pytest.mark.parametrize("name", CASES)
def test_factory_solver_sync_free_equals_the_host_loop(name):
    """SolverFactory::create -> next_iteration() until finished() (the reference's loop) against advance(nlog) stretches:
    same number of iterations, same time, same state; and both equal the reference's last dump"""
    from euler_kokkos_b200 import hostapi
    from euler_kokkos_b200.vti import embed_interior, interior_of

    z = np.load(os.path.join(HERE, "golden", name + ".npz"))
    last = int(z["steps"][-1])
    with tempfile.TemporaryDirectory() as td:
        ini = os.path.join(td, "case.ini")
        with open(ini, "w") as f:
            f.write(edit_ini(str(z["ini"]), outputdir=td, noutput=0, nstepmax=last, nlog=1000))
        a = hostapi.Solver(ini)
        while not a.finished():
            a.next_iteration()
        b = hostapi.Solver(ini)
        while not b.finished():
            assert b.advance(4) > 0
        p = a.case.params
        assert a.iteration == b.iteration == last
        assert abs(a.t - b.t) <= 1e-13 * abs(a.t)
        Ua, Ub = a.state(), b.state()
        err = util.group_rel_linf(Ub, Ua, p)
        assert max(err.values()) < 1e-13, err
        ref = embed_interior(z[f"u{last}"], p.ghost, p.dim)
        err = util.group_rel_linf(embed_interior(interior_of(Ua, p.ghost, p.dim), p.ghost, p.dim), ref, p)
        assert max(err.values()) < 1e-11, err
        a.close()
        b.close()


def test_asynchronous_output_writes_the_same_files():
    """[output] async_io (default yes): every dump is a device snapshot copied to page-locked memory on its own stream and
    written by a helper thread beside the next iterations; the files are byte-identical to synchronous dumps"""
    z = np.load(os.path.join(HERE, "golden", "ot3d_hlld.npz"))
    blobs = {}
    for mode in ("yes", "no"):
        with tempfile.TemporaryDirectory() as td:
            text = edit_ini(str(z["ini"]), outputdir=td, noutput=-1).replace("[output]\n", f"[output]\nasync_io={mode}\n")
            ini = os.path.join(td, "case.ini")
            with open(ini, "w") as f:
                f.write(text)
            r = subprocess.run([BIN, ini], cwd=td, capture_output=True, text=True, timeout=300)
            assert r.returncode == 0, r.stderr[-1000:]
            files = sorted(glob.glob(os.path.join(td, "g_*.vti")))
            blobs[mode] = [open(fn, "rb").read() for fn in files]
    assert len(blobs["yes"]) == len(blobs["no"]) >= 8
    assert blobs["yes"] == blobs["no"]


def test_restart_is_refused_and_ascii_output_is_written():
    """a configuration the drop-in cannot honour (restart: HDF5 loaders absent) stops with an error instead of silently
    doing something else; [output] outputVtkAscii=true writes the reference's text format (byte-compared with the
    reference binary on the CPU tier: test_abi_and_host.py::test_ascii_vti_dump_is_byte_identical_to_the_reference)"""
    z = np.load(os.path.join(HERE, "golden", "implode2d_hllc.npz"))
    with tempfile.TemporaryDirectory() as td:
        ini = os.path.join(td, "case.ini")
        text = edit_ini(str(z["ini"]), outputdir=td).replace("[run]\n", "[run]\nrestart_enabled=1\n")
        with open(ini, "w") as f:
            f.write(text)
        r = subprocess.run([BIN, ini], cwd=td, capture_output=True, text=True, timeout=120)
        assert r.returncode != 0 and "restart" in r.stderr
        text = edit_ini(str(z["ini"]), outputdir=td, outputvtkascii="true")
        with open(ini, "w") as f:
            f.write(text)
        r = subprocess.run([BIN, ini], cwd=td, capture_output=True, text=True, timeout=120)
        assert r.returncode == 0, r.stderr[-1000:]
        files = sorted(glob.glob(os.path.join(td, "*.vti")))
        assert files
        with open(files[0], "rb") as f:
            blob = f.read()
        assert blob.startswith(b'<?xml version="1.0"?>\n') and b'format="ascii"' in blob and b"AppendedData" not in blob


def test_shipped_brio_wu_3d_z_asks_for_impl_2_and_is_refused_loudly():
    """test_mhd_brio_wu_3d_{x,y,z}.ini ship with [other] implementationVersion=2.  In 3-D that is a different scheme in
    the reference (its own impl 2 and impl 0 differ by 1e-1 on this very file); the drop-in refuses it with a message
    instead of running impl-0 semantics under that name.  With the version set to 1 the same file runs."""
    z = np.load(os.path.join(HERE, "golden", "briowu3d_z_hlld.npz"))
    with tempfile.TemporaryDirectory() as td:
        ini = os.path.join(td, "case.ini")
        with open(ini, "w") as f:
            f.write(edit_ini(str(z["ini"]), outputdir=td, implementationversion=2, noutput=0))
        r = subprocess.run([BIN, ini], cwd=td, capture_output=True, text=True, timeout=120)
        assert r.returncode != 0
        assert "implementationVersion=2 is not provided" in r.stderr, r.stderr[-1000:]
        with open(ini, "w") as f:
            f.write(edit_ini(str(z["ini"]), outputdir=td, implementationversion=1, noutput=0))
        r = subprocess.run([BIN, ini], cwd=td, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-1000:]
