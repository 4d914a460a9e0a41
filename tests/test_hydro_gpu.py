"""GPU parity tests of the fused hydro step (through the C ABI) against the CPU oracle."""
import numpy as np
import pytest

from tests import util

pytestmark = pytest.mark.gpu

TOL = 1e-12  # north_star: per-step conserved fields agree within a relative L-inf of 1e-12


def _check_steps(p, U0, nsteps, grav=None, tol=TOL):
    from euler_kokkos_b200 import capi
    from oracle import oracle_lib as O

    op = O.params_from(p)
    ctx = capi.Context(p)
    if grav is not None:
        ctx.set_gravity(grav)
    ctx.upload(U0)
    Uo = U0.copy()
    for s in range(nsteps):
        dt_o = op.cfl / O.invdt(op, Uo, grav)
        dt_g = ctx.compute_dt()
        assert abs(dt_g - dt_o) <= 1e-13 * dt_o, (s, dt_g, dt_o)
        # one-step parity: feed the oracle's state n to both, compare n+1
        ctx.upload(Uo)
        ctx.step(dt_o)
        Ug = ctx.download()
        Uo = O.step(op, Uo, dt_o, grav)
        err = util.group_rel_linf(Ug, Uo, p)
        assert max(err.values()) < tol, (s, err)
    ctx.close()


@pytest.mark.parametrize("riemann", ["hllc", "approx", "llf"])
@pytest.mark.parametrize("n", [(16, 12, 8), (37, 29, 19), (64, 30, 40)])
def test_hydro3d_one_step_parity(riemann, n):
    p = util.make_params(dim=3, n=n, riemann=riemann, box=((0, 2.0), (0, 1.5), (0, 1.0)))
    _check_steps(p, util.random_hydro_state(p, seed=1), 3)


@pytest.mark.parametrize("knobs", [{"EK_HYDRO_FOLD": "0"}, {"EK_HYDRO_FOLD": "1"}])
@pytest.mark.parametrize("riemann", ["hllc", "approx", "llf"])
def test_hydro3d_kernel_variants_agree_with_the_oracle(knobs, riemann, monkeypatch):
    """every selectable 3-D schedule (one warp per tile row, folded halo rows with 16 / 15 warps, one barrier per plane) is held to the
    same one-step tolerance, with reflective / outflow / periodic faces and several z chunks."""
    for k, v in knobs.items():
        monkeypatch.setenv(k, v)
    p = util.make_params(dim=3, n=(61, 33, 40), riemann=riemann, bc=(3, 3, 1, 2, 2, 1))
    _check_steps(p, util.random_hydro_state(p, seed=11), 2)


@pytest.mark.parametrize("bc", [(1, 1, 1, 1, 1, 1), (2, 2, 2, 2, 2, 2), (3, 3, 3, 3, 3, 3), (3, 3, 1, 2, 2, 1)])
def test_hydro3d_boundaries(bc):
    p = util.make_params(dim=3, n=(20, 14, 10), bc=bc)
    _check_steps(p, util.random_hydro_state(p, seed=2), 2)


@pytest.mark.parametrize("riemann", ["hllc", "approx", "llf", "hll"])
@pytest.mark.parametrize("n", [(32, 16), (61, 45)])
def test_hydro2d_one_step_parity(riemann, n):
    p = util.make_params(dim=2, n=(n[0], n[1], 1), riemann=riemann, box=((0, 2.0), (0, 1.0), (0, 1.0)))
    _check_steps(p, util.random_hydro_state(p, seed=3), 3)


@pytest.mark.parametrize("slope_type", [0.0, 1.0, 2.0])
def test_hydro3d_slope_types(slope_type):
    p = util.make_params(dim=3, n=(18, 16, 12), slope_type=slope_type)
    _check_steps(p, util.random_hydro_state(p, seed=4), 2)


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("impl,hancock", [(0, False), (1, False), (0, True), (1, True)])
def test_hydro_gravity(dim, impl, hancock):
    p = util.make_params(dim=dim, n=(20, 12, 16 if dim == 3 else 1), riemann="approx", gamma0=1.4, cfl=0.7,
                         bc=(3, 3, 3, 3, 1, 1) if dim == 3 else (3, 3, 1, 1, 1, 1), impl=impl, gravity=True,
                         hancock=hancock)
    nv, ks, js, is_ = p.shape
    grav = np.zeros((dim, ks, js, is_))
    grav[dim - 1] = -util.f32(0.1)
    grav[0] = 0.03
    _check_steps(p, util.random_hydro_state(p, seed=5), 2, grav=np.ascontiguousarray(grav))


def test_hydro3d_blast_free_run():
    """free-running N steps with device-resident dt (ek_run) vs the oracle's own loop."""
    from euler_kokkos_b200 import capi
    from oracle import oracle_lib as O

    p = util.make_params(dim=3, n=(24, 20, 16))
    U0 = util.blast_state(p)
    ctx = capi.Context(p)
    ctx.upload(U0)
    t, dt_last, n = ctx.run(12)
    Ug = ctx.download()
    Uo, to, dts = O.run(O.params_from(p), U0, 12)
    assert n == 12
    assert abs(t - to) < 1e-12 * to
    err = util.group_rel_linf(Ug, Uo, p)
    assert max(err.values()) < 1e-11, err
    # conservation: reflective walls, no gravity -> mass and energy to round-off
    s0 = util.blast_state(p)
    from euler_kokkos_b200.vti import interior_of
    m0 = interior_of(s0, 2, 3)[0].sum()
    sums = ctx.conserved_sums()
    assert abs(sums[0] - m0) < 1e-12 * m0


def test_run_stops_at_tend():
    from euler_kokkos_b200 import capi

    p = util.make_params(dim=3, n=(16, 16, 16))
    ctx = capi.Context(p)
    ctx.upload(util.blast_state(p))
    dt0 = ctx.compute_dt()
    t_end = 2.5 * dt0
    t, dt_last, n = ctx.run(50, 0.0, t_end)
    assert n < 50 and abs(t - t_end) < 1e-14
