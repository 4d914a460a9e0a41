"""CPU tests of the CUDA MHD path's per-cell stages (mhd_cells.cuh / mhd_physics.cuh compiled for the
host, tests/emul) against the oracle: same code the kernels run, checked without a GPU."""
import numpy as np
import pytest

from tests import util
from tests.test_oracle_golden import load_case

TOL = 1e-12


def _one_step_chain(p, U, nsteps):
    from oracle import oracle_lib as O
    from tests.emul import emul_lib as E

    op = O.params_from(p)
    worst = {}
    for s in range(nsteps):
        O.make_boundaries(op, U)
        inv_o = O.invdt(op, U)
        inv_e = E.mhd_invdt(p, U)
        assert abs(inv_e - inv_o) <= 1e-13 * inv_o
        dt = op.cfl / inv_o
        Ue = E.mhd_step(p, U, dt)
        U = O.step(op, U, dt)
        err = util.group_rel_linf(Ue, U, p)
        assert max(err.values()) < TOL, (s, err)
        for k, v in err.items():
            worst[k] = max(worst.get(k, 0.0), v)
    return worst


@pytest.mark.parametrize("name", ["ot3d_hlld", "ot3d_hlld_slope1", "ot3d_hll", "ot3d_llf", "mhdblast3d_hlld", "mhdblast3d_hlld_bc",
                                  "ot2d_hlld", "ot2d_hlld_slope1", "ot2d_hll", "mhdblast2d_hlld_bc"])
def test_emul_on_reference_states(name):
    """one-step parity starting from states dumped by the reference binary (incl. shocked MHD blast)."""
    from euler_kokkos_b200.vti import embed_interior

    z, case = load_case(name)
    p = case.params
    for s in [0] + [int(x) for x in z["steps"]]:
        U = np.ascontiguousarray(embed_interior(z[f"u{s}"], p.ghost, p.dim))
        _one_step_chain(p, U, 1)


@pytest.mark.parametrize("riemann", ["hlld", "hll", "llf"])
@pytest.mark.parametrize("bc", [(3, 3, 3, 3, 3, 3), (1, 1, 2, 2, 1, 2)])
def test_emul_random_states(riemann, bc):
    """rough random states: exercises the limiter branches, supersonic HLLD branches and all 2-D HLLD cases."""
    p = util.make_params(dim=3, mhd=True, n=(13, 11, 9), riemann=riemann, bc=bc, smallr=1e-8, smallc=1e-8,
                         box=((0, 1.0), (0, 0.8), (0, 0.5)))
    U = util.random_hydro_state(p, seed=11, mach=2.5)
    _one_step_chain(p, U, 3)


@pytest.mark.parametrize("riemann", ["hlld", "hll", "llf"])
@pytest.mark.parametrize("bc", [(3, 3, 3, 3, 0, 0), (1, 2, 2, 1, 0, 0)])
def test_emul_random_states_2d(riemann, bc):
    p = util.make_params(dim=2, mhd=True, n=(23, 17, 1), riemann=riemann, bc=bc, smallr=1e-7, smallc=1e-7,
                         box=((0, 1.0), (0, 0.8), (0, 1.0)))
    U = util.random_hydro_state(p, seed=12, mach=2.5)
    _one_step_chain(p, U, 3)


@pytest.mark.parametrize("shape,yc", [((70, 70, 22), 32), ((38, 37, 14), 8), ((134, 29, 9), 5), ((22, 18, 30), 0), ((22, 18, 30), 18),
                                      ((518, 70, 6), 32), ((33, 65, 7), 64)])
@pytest.mark.parametrize("kg", [1, 4, 8])
def test_chunked_launch_order_covers_every_cell_once(shape, yc, kg):
    """mhd_chunk_cell (the thread -> cell map of the 3-D stencil kernels): a bijection onto the ghosted box that
    keeps warps aligned with the AoSoA tiles."""
    from tests.emul import emul_lib as E

    assert E.chunk_map_violations(*shape, yc, kg) == 0
