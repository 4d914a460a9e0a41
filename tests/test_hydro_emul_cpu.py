"""CPU tests of the fused CUDA hydro step's arithmetic: the per-cell / per-face functions of hydro_physics.cuh (the
very functions hydro_step_kernel calls) compiled for the host and driven in the kernel's order of operations
(tests/emul/hydro_emul.cpp), against the oracle.  The kernel's tiling / shared-memory logic is the GPU tier's job."""
import numpy as np
import pytest

from tests import util
from tests.test_oracle_golden import load_case

TOL = 1e-12  # north_star: per-step conserved fields agree within a relative L-inf of 1e-12


def _one_step_chain(p, U, nsteps, grav=None):
    from oracle import oracle_lib as O
    from tests.emul import emul_lib as E

    op = O.params_from(p)
    for s in range(nsteps):
        inv_o = O.invdt(op, U, grav)  # hydro: dt from the state before the boundary fill (interior cells)
        inv_e = E.hydro_invdt(p, U, grav)
        assert abs(inv_e - inv_o) <= 1e-13 * inv_o, (s, inv_e, inv_o)
        dt = op.cfl / inv_o
        O.make_boundaries(op, U)
        Ue = E.hydro_step(p, U, dt, grav)
        U = O.step(op, U, dt, grav)
        err = util.group_rel_linf(Ue, U, p)
        assert max(err.values()) < TOL, (s, err)
        O.make_boundaries(op, U)  # Uout's ghosts are zero until the next fill; the next invdt reads the interior only
    return U


@pytest.mark.parametrize("name", ["implode2d_hllc", "blast3d_hllc", "implode3d_hllc", "blast3d_hllc_big", "rt3d_approx_grav_impl1",
                                  "fourquad2d_hllc", "fourquad3d_hllc", "gresho2d_hllc", "kh3d_hllc"])
def test_hydro_emul_on_reference_states(name):
    """one-step parity starting from states dumped by the reference binary."""
    from euler_kokkos_b200.vti import embed_interior

    z, case = load_case(name)
    p = case.params
    _, grav = case.new_state()
    for s in [0] + [int(x) for x in z["steps"]]:
        U = np.ascontiguousarray(embed_interior(z[f"u{s}"], p.ghost, p.dim))
        _one_step_chain(p, U, 1, grav)


@pytest.mark.parametrize("riemann", ["hllc", "approx", "llf"])
@pytest.mark.parametrize("bc", [(1, 1, 1, 1, 1, 1), (3, 3, 1, 2, 2, 1)])
def test_hydro3d_emul_random_states(riemann, bc):
    p = util.make_params(dim=3, n=(17, 13, 11), riemann=riemann, bc=bc, box=((0, 2.0), (0, 1.5), (0, 1.0)))
    _one_step_chain(p, util.random_hydro_state(p, seed=2, mach=2.0), 3)


@pytest.mark.parametrize("riemann", ["hllc", "approx", "llf", "hll"])
def test_hydro2d_emul_random_states(riemann):
    p = util.make_params(dim=2, n=(37, 29, 1), riemann=riemann, bc=(1, 2, 3, 3, 0, 0))
    _one_step_chain(p, util.random_hydro_state(p, seed=3, mach=2.0), 3)


@pytest.mark.parametrize("slope_type", [0.0, 1.0, 2.0])
def test_hydro3d_emul_slope_types(slope_type):
    p = util.make_params(dim=3, n=(12, 10, 9), slope_type=slope_type)
    _one_step_chain(p, util.random_hydro_state(p, seed=4), 2)


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("impl,hancock", [(0, False), (1, False), (0, True), (1, True)])
def test_hydro_emul_gravity(dim, impl, hancock):
    """gravity source term + Hancock gravity predictor, incl. implementationVersion 1's omission on the ZMIN-face
    state (HydroRunFunctors3D.h:1351 vs :1277-1285) and the 2-D floor of the new density."""
    p = util.make_params(dim=dim, n=(14, 10, 12 if dim == 3 else 1), riemann="approx", gamma0=1.4, cfl=0.7,
                         bc=(3, 3, 3, 3, 1, 1) if dim == 3 else (3, 3, 1, 1, 1, 1), impl=impl, gravity=True, hancock=hancock)
    nv, ks, js, is_ = p.shape
    grav = np.zeros((dim, ks, js, is_))
    grav[dim - 1] = -util.f32(0.1)
    grav[0] = 0.03
    _one_step_chain(p, util.random_hydro_state(p, seed=5), 2, grav=np.ascontiguousarray(grav))
