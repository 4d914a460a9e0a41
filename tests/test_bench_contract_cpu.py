"""CPU tests of bench.py's contract: the reference arm prints ONE JSON line with the agreed keys (it times the
reference's own OpenMP build, oracle/_ref, and needs no GPU); our arm refuses to run without a CUDA device instead
of falling back to anything."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, timeout=600):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=timeout,
                          cwd=ROOT)


def test_reference_arm_prints_one_json_line():
    from oracle import oracle_lib as O

    if not O.have_reference():
        pytest.skip("oracle/_ref is not built here")
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-sample-n", "128")  # default sample: 256^3 (29 GB, ~30 s here)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "cell-updates/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["dtype"] == "f64" and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "test_mhd_orszag_tang_3D.ini @ 512x512x512" in d["config"]["workload"]  # the north-star target, same string as our arm
    assert "128x128x128" in d["cpu_baseline"]["sample"]


def test_reference_arm_does_not_load_the_product_libraries():
    """the ini of the sample is edited in plain Python: neither capi (libek_b200.so) nor hostapi (libek_host.so) is
    imported by the process that times the reference"""
    from oracle import oracle_lib as O

    if not O.have_reference():
        pytest.skip("oracle/_ref is not built here")
    code = ("import sys, argparse; sys.path.insert(0, %r); import bench; "
            "a = argparse.Namespace(workload='implode2d', n=None, shape=None); "
            "v, desc, thr = bench.reference_sample(a, None, 5, threads=2); assert v > 0; "
            "bad = [m for m in sys.modules if m.startswith('euler_kokkos_b200')]; assert not bad, bad; "
            "maps = open('/proc/self/maps').read(); assert 'libek_' not in maps; print('ok')") % ROOT
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]


def test_our_arm_has_no_cpu_fallback():
    try:
        import torch

        if torch.cuda.is_available():
            pytest.skip("a CUDA device is present")
    except ImportError:
        pass
    r = _run("--steps", "1", "--warmup", "3", "--no-cpu-baseline", timeout=300)
    assert r.returncode != 0
    assert not [ln for ln in r.stdout.splitlines() if ln.strip().startswith("{")], r.stdout
