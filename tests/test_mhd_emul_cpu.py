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
def test_chunked_launch_order_covers_every_cell_once(shape, yc):
    """mhd_chunk_cell (the thread -> cell map of the 3-D stencil kernels): a bijection onto the ghosted box that
    keeps warps aligned with the AoSoA tiles."""
    from tests.emul import emul_lib as E

    assert E.chunk_map_violations(*shape, yc) == 0


@pytest.mark.parametrize("bc", [(3, 3, 3, 3, 3, 3), (1, 1, 2, 2, 1, 2)])
@pytest.mark.parametrize("layer", ["x", "y", "z", "corner", "none"])
def test_emul_folded_cfl_reduction(layer, bc):
    """mhd_update_kernel<true> + mhd_invdt_layers_kernel (host build of the same per-cell code): the partial maximum
    formed inside the update plus the last-layer pass after the boundary fill equals the stand-alone reduction over
    the ghost-filled new state bit for bit, and the oracle's (ComputeDtFunctor3D_MHD) to round-off -- also when the
    fastest cells sit in the layers whose upper faces are ghosts."""
    from oracle import oracle_lib as O
    from tests.emul import emul_lib as E

    p = util.make_params(dim=3, mhd=True, n=(12, 10, 9), riemann="hlld", bc=bc, smallr=1e-8, smallc=1e-8)
    U = util.random_hydro_state(p, seed=4)
    gw = p.ghost
    sl = {"x": np.s_[:, :, :, gw + p.nx - 1], "y": np.s_[:, :, gw + p.ny - 1, :], "z": np.s_[:, gw + p.nz - 1, :, :],
          "corner": np.s_[:, gw + p.nz - 1, gw + p.ny - 1, gw + p.nx - 1]}
    if layer != "none":
        fast = U[sl[layer]]
        fast[2:5] *= 6.0
        fast[1] += 0.5 * (fast[2] ** 2 + fast[3] ** 2 + fast[4] ** 2) / fast[0]
    op = O.params_from(p)
    O.make_boundaries(op, U)
    dt = op.cfl / O.invdt(op, U)
    U1, partial = E.mhd_step_with_partial_invdt(p, U, dt)
    O.make_boundaries(op, U1)
    layers = E.mhd_invdt_layers(p, U1)
    full = E.mhd_invdt(p, U1)
    assert max(partial, layers) == full, (partial, layers, full)
    if layer != "none":
        assert layers >= partial  # the fast cells are where only the layer pass can see them
    inv_o = O.invdt(op, U1)
    assert abs(full - inv_o) <= 1e-13 * inv_o
