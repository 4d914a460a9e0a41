"""GPU parity on EVERY fixture the reference binary produced (tests/golden/*.npz: the BASELINE configs plus the other
problems of the reference -- four-quadrant Riemann problems, Gresho and isentropic vortices, Kelvin-Helmholtz,
Brio-Wu shock tubes, rotor, field loop, MHD blast): through the C ABI,
  * one step from each dumped state vs the CPU oracle (relative L-inf per physical group < 1e-12), and
  * a free run with the device-resident time step vs the reference's own later dumps (< 1e-11)."""
import numpy as np
import pytest

from tests import util
from tests.test_oracle_golden import CASES, load_case

pytestmark = pytest.mark.gpu


def _grav(case, p):
    if not p.gravity_enabled:
        return None
    _, grav = case.new_state()
    return grav


@pytest.mark.parametrize("name", CASES)
def test_one_step_from_reference_states(name):
    from euler_kokkos_b200 import capi
    from euler_kokkos_b200.vti import embed_interior
    from oracle import oracle_lib as O

    z, case = load_case(name)
    p = case.params
    op = O.params_from(p)
    grav = _grav(case, p)
    ctx = capi.Context(p)
    if grav is not None:
        ctx.set_gravity(grav)
    for s in [0] + [int(x) for x in z["steps"]]:
        U = np.ascontiguousarray(embed_interior(z[f"u{s}"], p.ghost, p.dim))
        O.make_boundaries(op, U)
        dt = op.cfl / (O.invdt(op, U, grav) if not p.mhd else O.invdt(op, U))
        ctx.upload(U)
        dt_g = ctx.compute_dt()
        assert abs(dt_g - dt) <= 1e-13 * dt, (name, s, dt_g, dt)
        ctx.step(dt)
        Uo = O.step(op, U, dt, grav) if not p.mhd else O.step(op, U, dt)
        err = util.group_rel_linf(ctx.download(), Uo, p)
        assert max(err.values()) < 1e-12, (name, s, err)
    ctx.close()


@pytest.mark.parametrize("name", CASES)
def test_free_run_matches_reference_dumps(name):
    from euler_kokkos_b200 import capi
    from euler_kokkos_b200.vti import embed_interior

    z, case = load_case(name)
    p = case.params
    grav = _grav(case, p)
    ctx = capi.Context(p)
    if grav is not None:
        ctx.set_gravity(grav)
    ctx.upload(np.ascontiguousarray(embed_interior(z["u0"], p.ghost, p.dim)))
    done, t = 0, 0.0
    for s in [int(x) for x in z["steps"]]:
        t, _, n = ctx.run(s - done, t)
        assert n == s - done
        done = s
        ref = np.ascontiguousarray(embed_interior(z[f"u{s}"], p.ghost, p.dim))
        err = util.group_rel_linf(ctx.download(), ref, p)
        assert max(err.values()) < 1e-11, (name, s, err)
    ctx.close()
