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


@pytest.mark.parametrize("ru", ["0", "1"])
@pytest.mark.parametrize("overlap", ["1", "0"])
def test_mhd3d_two_slabs_periodic_match_single_domain(tmp_path, monkeypatch, overlap, ru):
    """overlap=1: the update stage does the END planes first (ru=1, the fused R+U kernel: the end z-chunks) and their
    exchange runs beside the update of the rest"""
    monkeypatch.setenv("EK_HALO_OVERLAP", overlap)
    monkeypatch.setenv("EK_MHD_RU", ru)
    _run_case("test_mhd_orszag_tang_3D.ini", {"mesh.nx": 32, "mesh.ny": 32}, 12 if ru == "0" else 30, 2, 8, tmp_path, 1e-13)


# ---- mx x my x mz decompositions ([mpi] section of the ini, HydroParams.cpp:186-320): packed x / y halos ------------
def _worker_cart(rank, world, uid, ini, overrides, nsteps, outdir):
    import torch  # noqa: F401

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
    np.savez(os.path.join(outdir, f"rank{rank}.npz"), U=ctx.download(), t=t, pos=np.array(case.mpi_pos), sums=ctx.conserved_sums())
    ctx.close()


def _run_cart(ini_name, ov_local, dims, nsteps, tmp_path, tol):
    """every rank owns an nx x ny x nz box of an (mx*nx) x (my*ny) x (mz*nz) grid; the assembled result must equal the
    single-domain run (the reference's MPI layout: mpiBorderUtils.h:36-330, SolverBase.cpp:583-664)"""
    import torch
    import torch.multiprocessing as mp

    world = dims[0] * dims[1] * dims[2]
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    from euler_kokkos_b200 import capi, hostapi
    from euler_kokkos_b200.vti import interior_of

    ini = hostapi.settings(ini_name)
    uid = capi.comm_unique_id()
    ov = dict(ov_local)
    ov.update({"mpi.mx": dims[0], "mpi.my": dims[1], "mpi.mz": dims[2]})
    mp.spawn(_worker_cart, args=(world, uid, ini, ov, nsteps, str(tmp_path)), nprocs=world, join=True)

    ov1 = dict(ov_local)
    ov1.update({"mesh.nx": ov_local["mesh.nx"] * dims[0], "mesh.ny": ov_local["mesh.ny"] * dims[1], "mesh.nz": ov_local["mesh.nz"] * dims[2]})
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

    nx, ny, nz = ov_local["mesh.nx"], ov_local["mesh.ny"], ov_local["mesh.nz"]
    got = np.full_like(ref, np.nan)
    for r in range(world):
        z = np.load(os.path.join(str(tmp_path), f"rank{r}.npz"))
        assert abs(float(z["t"]) - t1) <= 1e-14 * abs(t1)
        np.testing.assert_allclose(z["sums"], sums1, rtol=1e-12, atol=1e-12 * abs(sums1[0]))
        px, py, pz = (int(v) for v in z["pos"])
        got[:, pz * nz:(pz + 1) * nz, py * ny:(py + 1) * ny, px * nx:(px + 1) * nx] = interior_of(z["U"], p.ghost, p.dim)
    assert not np.isnan(got).any()
    den = np.abs(ref).max(axis=(1, 2, 3), keepdims=True) + 1e-300
    err = (np.abs(got - ref) / den).max()
    assert err <= tol, err


@pytest.mark.parametrize("dims", [(2, 1, 1), (1, 2, 1)])
def test_hydro3d_x_and_y_halos_on_two_gpus(tmp_path, dims):
    """packed (non-contiguous) ghost slabs: halo_copy_kernel + ncclSend/Recv along x, along y"""
    _run_cart("test_implode_3D.ini", {"mesh.nx": 24, "mesh.ny": 20, "mesh.nz": 16}, dims, 8, tmp_path, 1e-13)


@pytest.mark.parametrize("dims", [(2, 1, 1), (1, 2, 1)])
def test_mhd3d_x_and_y_halos_on_two_gpus(tmp_path, dims):
    _run_cart("test_mhd_orszag_tang_3D.ini", {"mesh.nx": 20, "mesh.ny": 16, "mesh.nz": 12}, dims, 6, tmp_path, 1e-13)


def test_hydro3d_2x2x2_on_eight_gpus(tmp_path):
    """the layout of test_implode_3D_mpi.ini ([mpi] mx, my, mz > 1 in every direction), reflective outer faces"""
    _run_cart("test_implode_3D.ini", {"mesh.nx": 20, "mesh.ny": 16, "mesh.nz": 16}, (2, 2, 2), 10, tmp_path, 1e-13)


def test_mhd3d_2x2x2_periodic_on_eight_gpus(tmp_path):
    """periodic wrap in all three directions between the 8 sub-domains, corner / edge ghosts through the X -> Y -> Z order"""
    _run_cart("test_mhd_orszag_tang_3D.ini", {"mesh.nx": 16, "mesh.ny": 16, "mesh.nz": 12}, (2, 2, 2), 6, tmp_path, 1e-13)


def test_mhd2d_2x2_on_four_gpus(tmp_path):
    _run_cart("test_mhd_orszag_tang_2D.ini", {"mesh.nx": 32, "mesh.ny": 24, "mesh.nz": 1}, (2, 2, 1), 8, tmp_path, 1e-13)


# ---- the C++ driver itself on several ranks (one process per GPU, NCCL id through the rendezvous file) -------------
@pytest.mark.parametrize("ini_name,ov,dims", [
    ("test_mhd_orszag_tang_3D.ini", {"mesh.nx": 24, "mesh.ny": 20, "mesh.nz": 12}, (1, 1, 2)),
    ("test_implode_3D.ini", {"mesh.nx": 20, "mesh.ny": 24, "mesh.nz": 16}, (2, 1, 1)),
])
def test_euler_b200_executable_on_two_ranks_writes_the_single_domain_state(tmp_path, ini_name, ov, dims):
    """`torchrun --no-python --nproc-per-node 2 bin/euler_b200 file.ini` with an [mpi] section: the ranks find each
    other through the per-launch rendezvous file (comm_rendezvous.cpp), exchange halos over NCCL, and write a .pvti
    collection whose assembled last dump equals the single-process run of the same global grid (the reference needs an
    MPI build for this, main.cpp:54-111)."""
    import glob
    import subprocess
    import sys

    import torch

    world = dims[0] * dims[1] * dims[2]
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    from euler_kokkos_b200 import hostapi
    from euler_kokkos_b200.vti import read_pvti, read_vti

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    binary = os.path.join(root, "euler_kokkos_b200", "bin", "euler_b200")
    common = {"run.nstepmax": 6, "run.tend": 1e6, "run.noutput": -1, "run.nlog": 2, "output.outputprefix": "g",
              "output.hdf5_enabled": "no", "output.vtk_enabled": "yes"}
    # decomposed
    dmulti = tmp_path / "multi"
    dmulti.mkdir()
    o = dict(ov)
    o.update(common)
    o.update({"mpi.mx": dims[0], "mpi.my": dims[1], "mpi.mz": dims[2], "output.outputdir": str(dmulti)})
    ini_m = str(dmulti / "case.ini")
    hostapi.write_ini(hostapi.settings(ini_name), o, ini_m)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--no-python", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", "29577", binary, ini_m],
                       cwd=str(dmulti), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    assert not glob.glob("/tmp/ek_b200_rdv_29577_*"), "rank 0 removes the rendezvous file once the communicator is up"
    # single domain, same global grid
    dsingle = tmp_path / "single"
    dsingle.mkdir()
    o1 = dict(ov)
    o1.update(common)
    o1.update({"mesh.nx": ov["mesh.nx"] * dims[0], "mesh.ny": ov["mesh.ny"] * dims[1], "mesh.nz": ov["mesh.nz"] * dims[2],
               "output.outputdir": str(dsingle)})
    ini_s = str(dsingle / "case.ini")
    hostapi.write_ini(hostapi.settings(ini_name), o1, ini_s)
    r1 = subprocess.run([binary, ini_s], cwd=str(dsingle), capture_output=True, text=True, timeout=600)
    assert r1.returncode == 0, r1.stdout[-1500:] + r1.stderr[-1500:]
    # same log lines from rank 0 as from the single process
    from tests.util import log_lines
    assert log_lines(r.stdout) == log_lines(r1.stdout)
    pv = sorted(glob.glob(str(dmulti / "g_time*.pvti")))
    sv = sorted(glob.glob(str(dsingle / "g_*.vti")))
    assert len(pv) == len(sv) and len(pv) >= 7
    names_m, got = read_pvti(pv[-1])
    names_s, ref = read_vti(sv[-1])
    assert names_m == names_s
    den = np.abs(ref).max(axis=(1, 2, 3), keepdims=True) + 1e-300
    assert (np.abs(got - ref) / den).max() <= 1e-13
