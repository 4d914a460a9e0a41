"""The device-resident time loop replayed as a CUDA graph (ek_run, single GPU, launch-bound grids: two consecutive steps
captured once per starting level) against the same loop as plain launches: the same kernels in the same order, so the
states must be bitwise equal -- for odd and even step counts, repeated calls (cached graph, both level parities), an
early finish on t_end inside a replayed pair, and every solver family."""
import numpy as np
import pytest

from tests import util

pytestmark = pytest.mark.gpu

CASES = [dict(dim=2, n=(96, 48, 1), riemann="hllc"), dict(dim=3, n=(40, 24, 20), riemann="hllc"),
         dict(dim=3, n=(28, 20, 12), riemann="approx", gravity=True, bc=(3, 3, 3, 3, 1, 1)),
         dict(dim=3, mhd=True, n=(26, 16, 12), riemann="hlld", bc=(3, 3, 3, 3, 3, 3), smallr=1e-8, smallc=1e-8),
         dict(dim=2, mhd=True, n=(48, 40, 1), riemann="hlld", bc=(3, 3, 3, 3, 3, 3), smallr=1e-8, smallc=1e-8)]


def _run(p, U0, grav, plan, graph, monkeypatch, t_end=1e300):
    from euler_kokkos_b200 import capi

    monkeypatch.setenv("EK_GRAPH", graph)
    ctx = capi.Context(p)
    if grav is not None:
        ctx.set_gravity(grav)
    ctx.upload(U0)
    t, out = 0.0, []
    l0 = ctx.launch_count
    for n in plan:
        t, dt, done = ctx.run(n, t, t_end)
        out.append((t, dt, done, ctx.download().copy()))
    launches = ctx.launch_count - l0
    ctx.close()
    return out, launches


@pytest.mark.parametrize("kw", CASES, ids=lambda k: f"{'mhd' if k.get('mhd') else 'hydro'}{k['dim']}d")
def test_graph_replay_equals_plain_launches(kw, monkeypatch):
    p = util.make_params(**kw)
    U0 = util.random_hydro_state(p, seed=21)
    grav = None
    if p.gravity_enabled:
        grav = np.zeros((p.dim,) + tuple(p.shape[1:]))
        grav[p.dim - 1] = -0.1
    plan = (7, 1, 6, 2, 9)  # odd / even counts, both starting parities, counts below the graph threshold
    a, la = _run(p, U0, grav, plan, "0", monkeypatch)
    b, lb = _run(p, U0, grav, plan, "1", monkeypatch)
    assert la == lb  # kernels launched: a replay counts the kernels of its graph
    for (ta, dta, na, Ua), (tb, dtb, nb, Ub) in zip(a, b):
        assert na == nb and ta == tb and dta == dtb
        assert np.array_equal(Ua, Ub)


def test_graph_replay_stops_on_t_end_like_the_plain_loop(monkeypatch):
    p = util.make_params(dim=2, n=(64, 32, 1), riemann="hllc")
    U0 = util.random_hydro_state(p, seed=4)
    # find the time after 5 steps, then ask for 20 steps with t_end inside the 6th
    ref, _ = _run(p, U0, None, (5, 1), "0", monkeypatch)
    t_end = 0.5 * (ref[0][0] + ref[1][0])
    a, _ = _run(p, U0, None, (20,), "0", monkeypatch, t_end=t_end)
    b, _ = _run(p, U0, None, (20,), "1", monkeypatch, t_end=t_end)
    assert a[0][2] == b[0][2] == 6 and a[0][0] == b[0][0] == t_end
    assert np.array_equal(a[0][3], b[0][3])
