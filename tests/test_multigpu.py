"""2-GPU test of the slab decomposition (NCCL ghost-plane exchange + cross-rank CFL reduction):
the decomposed run must reproduce the single-domain run of the same global grid.  The reference has
no such test (it needs MPI to run decomposed at all).  Needs 2 visible GPUs;
skipped otherwise (run it with `gpurun --gpus 2`)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _worker(rank, world, uid, ini, overrides, nsteps, outdir):
    import torch  # noqa: F401  (CUDA context / device selection)

    from euler_kokkos_b200 import capi, hostapi

    torch.cuda.set_device(rank)
    case = hostapi.Case(ini, overrides, rank=rank, nranks=world)
    p = case.params
    U, grav = case.new_state()
    ctx = capi.Context(p, rank)
    if grav is not None:
        ctx.set_gravity(grav)
    ctx.comm_attach(uid, rank, world, case.neighbors)
    ctx.upload(U)
    t, dt, n = ctx.run(nsteps)
    out = ctx.download()
    sums = ctx.conserved_sums()
    np.savez(os.path.join(outdir, f"rank{rank}.npz"), U=out, t=t, sums=sums)
    ctx.close()


def _run_case(ini_name, overrides_local, nz_local, world, nsteps, tmp_path, tol):
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    from euler_kokkos_b200 import capi, hostapi
    from euler_kokkos_b200.vti import interior_of

    ini = hostapi.settings(ini_name)
    uid = capi.comm_unique_id()
    ov = dict(overrides_local)
    ov["mesh.nz"] = nz_local
    mp.spawn(_worker, args=(world, uid, ini, ov, nsteps, str(tmp_path)), nprocs=world, join=True)

    # single domain with the same GLOBAL grid
    ov1 = dict(overrides_local)
    ov1["mesh.nz"] = nz_local * world
    case = hostapi.Case(ini, ov1)
    p = case.params
    U, grav = case.new_state()
    ctx = capi.Context(p, 0)
    if grav is not None:
        ctx.set_gravity(grav)
    ctx.upload(U)
    t1, _, _ = ctx.run(nsteps)
    ref = interior_of(ctx.download(), p.ghost, p.dim)
    sums1 = ctx.conserved_sums()
    ctx.close()

    parts = []
    for r in range(world):
        z = np.load(os.path.join(str(tmp_path), f"rank{r}.npz"))
        assert abs(float(z["t"]) - t1) <= 1e-14 * abs(t1)
        parts.append(interior_of(z["U"], p.ghost, p.dim))
        np.testing.assert_allclose(z["sums"], sums1, rtol=1e-12, atol=1e-12 * abs(sums1[0]))  # all-reduced sums
    got = np.concatenate(parts, axis=1)  # stack slabs along z
    assert got.shape == ref.shape
    den = np.abs(ref).max(axis=(1, 2, 3), keepdims=True) + 1e-300
    err = (np.abs(got - ref) / den).max()
    assert err <= tol, err


@pytest.mark.parametrize("overlap", ["1", "0"])
def test_hydro3d_two_slabs_match_single_domain(tmp_path, monkeypatch, overlap):
    """overlap=1: the z-halo exchange of step n+1 runs beside the middle z-chunks of step n (ek_api.cu enqueue_step);
    overlap=0: exchange at the start of the step, as the reference's MPI path does.  Same answer either way."""
    monkeypatch.setenv("EK_HALO_OVERLAP", overlap)
    _run_case("test_implode_3D.ini", {"mesh.nx": 40, "mesh.ny": 36}, 20, 2, 10, tmp_path, 1e-13)


def test_hydro3d_two_slabs_periodic_z_with_gravity(tmp_path):
    """both z faces of both slabs are halo faces (periodic wrap), gravity + approx solver (config 4 physics)"""
    _run_case("test_rayleigh_taylor_3d.ini", {"mesh.nx": 24, "mesh.ny": 20, "mesh.boundary_type_zmin": 3, "mesh.boundary_type_zmax": 3},
              16, 2, 6, tmp_path, 1e-13)


@pytest.mark.parametrize("overlap", ["1", "0"])
def test_mhd3d_two_slabs_periodic_match_single_domain(tmp_path, monkeypatch, overlap):
    """overlap=1: the update stage does the END planes first and their exchange runs beside the update of the rest"""
    monkeypatch.setenv("EK_HALO_OVERLAP", overlap)
    _run_case("test_mhd_orszag_tang_3D.ini", {"mesh.nx": 32, "mesh.ny": 32}, 12, 2, 8, tmp_path, 1e-13)
