"""world_size-2 CPU test (gloo) of the N>1 host logic: slab decomposition through the C++ host layer,
the ghost-plane exchange protocol of csrc/ek_comm.cu restated with torch.distributed send/recv (same
neighbour table, same plane offsets: gw contiguous planes per variable), physical BCs on the outer
faces, and the cross-rank max of the inverse time step.  The assembled, ghost-filled slabs must equal
the ghost-filled single-domain state bit for bit."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _exchange_z(dist, U, p, neighbors, rank):
    """comm_exchange_dir(dir = z) of ek_comm.cu: my planes [gw,2gw) -> min neighbour's max ghosts, my planes
    [n,n+gw) -> max neighbour's min ghosts."""
    import torch

    gw, nz = p.ghost, p.nz
    has_min, has_max = p.bc[4] == 4, p.bc[5] == 4
    reqs, recv = [], []
    t = torch.from_numpy(U)
    if has_min:
        reqs.append(dist.isend(t[:, gw:2 * gw].contiguous(), neighbors[4], tag=1))
    if has_max:
        reqs.append(dist.isend(t[:, nz:nz + gw].contiguous(), neighbors[5], tag=2))
    if has_max:
        buf = torch.empty_like(t[:, :gw])
        dist.recv(buf, neighbors[5], tag=1)
        recv.append((slice(nz + gw, nz + 2 * gw), buf))
    if has_min:
        buf = torch.empty_like(t[:, :gw])
        dist.recv(buf, neighbors[4], tag=2)
        recv.append((slice(0, gw), buf))
    for r in reqs:
        r.wait()
    for sl, buf in recv:
        U[:, sl] = buf.numpy()


def _worker(rank, world, port, ini, nz_local, outdir):
    import torch.distributed as dist

    from euler_kokkos_b200 import capi, hostapi
    from oracle import oracle_lib as O

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    case = hostapi.Case(ini, {"mesh.nx": 12, "mesh.ny": 10, "mesh.nz": nz_local}, rank=rank, nranks=world)
    p = case.params
    U, _ = case.new_state()
    # X, Y: physical BCs (not decomposed); Z: exchange where EK_BC_COPY, physical BC elsewhere
    q = p.copy()
    bc_z = (q.bc[4], q.bc[5])
    op = O.params_from(q)
    op.bc[4] = op.bc[5] = capi.BC_COPY
    O.make_boundaries(op, U)            # x and y faces only (copy faces are skipped by the oracle)
    _exchange_z(dist, U, p, case.neighbors, rank)
    op.bc[4], op.bc[5] = bc_z
    op.bc[0] = op.bc[1] = op.bc[2] = op.bc[3] = capi.BC_COPY
    O.make_boundaries(op, U)            # physical z faces of the outer slabs
    # CFL: max over ranks of the local inverse dt (ncclAllReduce(max) in the product)
    import torch

    oq = O.params_from(p)
    inv = torch.tensor([O.invdt(oq, U)], dtype=torch.float64)
    dist.all_reduce(inv, op=dist.ReduceOp.MAX)
    np.savez(os.path.join(outdir, f"r{rank}.npz"), U=U, inv=inv.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("ini_name", ["test_implode_3D.ini", "test_mhd_orszag_tang_3D.ini"])
def test_two_slabs_ghost_exchange_matches_single_domain(ini_name, tmp_path):
    import torch.multiprocessing as mp

    from euler_kokkos_b200 import hostapi
    from oracle import oracle_lib as O

    world, nz_local = 2, 6
    ini = hostapi.settings(ini_name)
    mp.spawn(_worker, args=(world, _free_port(), ini, nz_local, str(tmp_path)), nprocs=world, join=True)

    case = hostapi.Case(ini, {"mesh.nx": 12, "mesh.ny": 10, "mesh.nz": nz_local * world})
    p = case.params
    G, _ = case.new_state()
    op = O.params_from(p)
    O.make_boundaries(op, G)
    inv_global = O.invdt(op, G)
    gw = p.ghost
    for r in range(world):
        z = np.load(os.path.join(str(tmp_path), f"r{r}.npz"))
        k0 = r * nz_local
        want = G[:, k0:k0 + nz_local + 2 * gw]
        assert np.array_equal(z["U"], want), (ini_name, r, np.abs(z["U"] - want).max())
        assert float(z["inv"][0]) == inv_global
