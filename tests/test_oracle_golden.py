"""CPU tests: the plain-C oracle (oracle/*.c) against the golden fixtures produced by the UNMODIFIED
reference binary (tests/golden/make_golden.py), and the host layer's initial conditions against the
reference's own.  These pin the oracle that the GPU parity tests use as checker."""
import os
import tempfile

import numpy as np
import pytest

from tests import util

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(f[:-4] for f in os.listdir(GOLDEN) if f.endswith(".npz"))

# The oracle follows the reference's floating-point order; observed agreement is bit-exact or within
# a few ulp (the reference is compiled -O3 without FMA, like the oracle).  Gate: 1e-13 relative L-inf
# per physical group, i.e. 10x tighter than the 1e-12 gate of the GPU path.
TOL = 1e-13


def load_case(name):
    from euler_kokkos_b200 import hostapi

    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    with tempfile.NamedTemporaryFile("w", suffix=".ini", delete=False) as f:
        f.write(str(z["ini"]))
        path = f.name
    try:
        case = hostapi.Case(path)
    finally:
        os.unlink(path)
    return z, case


@pytest.mark.parametrize("name", CASES)
def test_initial_condition_matches_reference(name):
    """Init*Functor restatements of the host layer (csrc/host/InitConditions.cpp) vs the reference's dump 0."""
    from euler_kokkos_b200.vti import interior_of

    z, case = load_case(name)
    p = case.params
    # (MHD blast: the reference never assigns [blast] bz_in/bz_out (BlastParams.h:56-62); its dumps hold denormal garbage
    #  ~1e-309 there, ours 0 -- equal within the tolerance below)
    U, _ = case.new_state()
    ours = interior_of(U, p.ghost, p.dim)
    ref = z["u0"]
    assert ours.shape == ref.shape
    den = np.abs(ref).max()
    assert np.abs(ours - ref).max() <= 1e-15 * den, np.abs(ours - ref).max()


@pytest.mark.parametrize("name", CASES)
def test_oracle_free_run_matches_reference(name):
    """N free-running steps of the oracle (its own dt) from the reference's initial dump."""
    from euler_kokkos_b200.vti import embed_interior
    from oracle import oracle_lib as O

    z, case = load_case(name)
    p = case.params
    op = O.params_from(p)
    grav = None
    if p.gravity_enabled:
        _, grav = case.new_state()
    U = np.ascontiguousarray(embed_interior(z["u0"], p.ghost, p.dim))
    if not p.mhd:
        O.make_boundaries(op, U)
    t, done = 0.0, 0
    for s in [int(x) for x in z["steps"]]:
        U, t, _ = O.run(op, U, s - done, t=t, grav=grav)
        done = s
        ref = np.ascontiguousarray(embed_interior(z[f"u{s}"], p.ghost, p.dim))
        err = util.group_rel_linf(U, ref, p)
        assert max(err.values()) < TOL, (name, s, err)


KH_RAND = os.path.join(GOLDEN, "kh_rand", "ref.npz")
KH_RAND_CASES = sorted({k.split("/")[0] for k in np.load(KH_RAND).files}) if os.path.exists(KH_RAND) else []


@pytest.mark.parametrize("name", KH_RAND_CASES)
def test_kh_random_perturbation_matches_the_reference_stream(name, monkeypatch):
    """[KH] perturbation_rand=1: the reference draws from Kokkos::Random_XorShift64_Pool inside an MDRangePolicy loop
    (HydroInitFunctors2D.h:214-348, HydroInitFunctors3D.h:270-426, MHDInitFunctors2D.h:484-676).  The host layer
    reproduces the stream of the reference's OpenMP back-end for a given thread count -- pool seeding, xorshift64*,
    2 x full / 2 x 2 x full host tiles, schedule(static, 1) tile-to-thread map -- so that dump 0 of the reference binary
    run with 1, 3 and 4 threads (tests/golden/make_kh_rand_fixture.py) is reproduced bit for bit."""
    from euler_kokkos_b200 import hostapi
    from euler_kokkos_b200.vti import interior_of

    z = np.load(KH_RAND)
    monkeypatch.setenv("OMP_NUM_THREADS", str(int(z[name + "/threads"])))
    monkeypatch.delenv("EK_KH_RAND_THREADS", raising=False)
    with tempfile.NamedTemporaryFile("w", suffix=".ini", delete=False) as f:
        f.write(str(z[name + "/ini"]))
        path = f.name
    try:
        case = hostapi.Case(path)
        U, _ = case.new_state()
    finally:
        os.unlink(path)
    p = case.params
    ours = interior_of(U, p.ghost, p.dim)
    ref = z[name + "/u0"]
    assert ours.shape == ref.shape
    assert np.array_equal(ours, ref), np.abs(ours - ref).max()
    # the explicit thread count wins over OMP_NUM_THREADS; another thread count is another stream (except with one tile row)
    monkeypatch.setenv("EK_KH_RAND_THREADS", str(int(z[name + "/threads"]) + 1))
    U2, _ = case.new_state()
    assert not np.array_equal(interior_of(U2, p.ghost, p.dim), ref)


def test_kh_random_perturbation_hash_generator_is_decomposition_independent():
    """[KH] rand_generator=hash: the counter-based generator keyed on (seed, global cell index) -- the same field whatever
    the thread count (the reference's stream is not)"""
    from euler_kokkos_b200 import hostapi
    from euler_kokkos_b200.vti import interior_of

    ov = {"mesh.nx": 16, "mesh.ny": 12, "kh.perturbation_rand": 1, "kh.perturbation_sine": 0, "kh.perturbation_sine_robertson": 0,
          "kh.rand_generator": "hash"}
    out = []
    for t in ("1", "5"):
        os.environ["EK_KH_RAND_THREADS"] = t
        try:
            case = hostapi.Case(hostapi.settings("test_muscl_kelvin_helmholtz_2D.ini"), ov)
            U, _ = case.new_state()
        finally:
            del os.environ["EK_KH_RAND_THREADS"]
        out.append(interior_of(U, case.params.ghost, 2))
    assert np.array_equal(out[0], out[1])
    assert np.std(out[0][2] / out[0][0]) > 0.0
