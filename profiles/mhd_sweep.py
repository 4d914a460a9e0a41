"""Tuning sweep (GPU box): time the MHD step at 512^3 for combinations of the env knobs of mhd_launch.cu,
one process per library variant (EK_B200_LIB).  usage: python profiles/mhd_sweep.py [n] [ini]"""
import itertools, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("EKH_QUIET", "1")
import numpy as np
from euler_kokkos_b200 import capi, hostapi

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
wl = sys.argv[2] if len(sys.argv) > 2 else "test_mhd_orszag_tang_3D.ini"
knobs = json.loads(os.environ.get("SWEEP", '{"EK_MHD_YCHUNK": ["0", "16", "32", "64"], "EK_MHD_FUSE_R": ["1", "0"]}'))
case = hostapi.Case(os.path.join(ROOT, "euler_kokkos_b200", "settings", wl), {"mesh.nx": n, "mesh.ny": n, "mesh.nz": n, "run.noutput": 0})
p = case.params
ctx = capi.Context(p, 0)
U0 = np.empty(p.shape)
case.init_state(U0)
ctx.upload(U0)
ctx.run(2)
cells = p.nx * p.ny * p.nz
names = sorted(knobs)
for combo in itertools.product(*[knobs[k] for k in names]):
    for k, v in zip(names, combo):
        os.environ[k] = v
    ctx.run(2)
    ctx.sync()
    t0 = time.time()
    ctx.run(5)
    ctx.sync()
    ms = (time.time() - t0) * 1e3 / 5
    print(json.dumps({"lib": os.path.basename(os.path.dirname(capi.LIB_PATH)), **dict(zip(names, combo)), "ms_per_step": round(ms, 3),
                      "gcell_s": round(cells / ms / 1e6, 3)}), flush=True)
ctx.close()
