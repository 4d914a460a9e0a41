#!/usr/bin/env python
"""bench.py -- cell-updates/s (FP64) of the MUSCL-Hancock time step on B200.

    python bench.py --gpus N --steps K --warmup W [--workload blast3d|ot3d|implode2d|rt3d] [--n 512]
    python bench.py --impl reference ...      # the reference's own CPU (OpenMP) build, same metric

A "step" is one godunov_unsplit over the whole (local) grid.  N=1 runs BASELINE.json configs[1]
(test_blast_3D.ini scaled to 512^3, Hydro_Muscl_3D, HLLC).  N>1 (one process per GPU under torchrun)
stacks one 512^3 slab per GPU along z (weak scaling, configs[4]); ghost planes travel over
NCCL/NVLink.  Rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
os.environ.setdefault("EKH_QUIET", "1")  # stdout carries exactly one JSON line
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ALGO_BYTES = {"hydro2d": 64, "hydro3d": 80, "hydro3d_grav": 104, "mhd": 128, "mhd2d": 128}

WORKLOADS = {
    # name: (ini file in euler_kokkos_b200/settings, solver family, default n)
    "blast3d": dict(ini="test_blast_3D.ini", fam="hydro3d", n=512, cubic_box=False),
    "implode3d": dict(ini="test_implode_3D.ini", fam="hydro3d", n=512, cubic_box=False),
    "ot3d": dict(ini="test_mhd_orszag_tang_3D.ini", fam="mhd", n=512, cubic_box=False),
    "rt3d": dict(ini="test_rayleigh_taylor_3d.ini", fam="hydro3d_grav", n=512, cubic_box=False),
    "implode2d": dict(ini="test_implode_2D.ini", fam="hydro2d", n=None, cubic_box=False),
    "ot2d": dict(ini="test_mhd_orszag_tang_2D.ini", fam="mhd2d", n=8192, cubic_box=False),
}


def recorded_traffic(fam, p):
    """per-launch DRAM traffic of the dominant kernel from the committed ncu capture (profiles/traffic.json), or None
    when this workload/size has not been captured"""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            d = json.load(f)
        key = f"{'mhd' if fam == 'mhd' else fam.replace('_grav', '')}@{p.nx}"
        if p.nx == p.ny == p.nz and key in d and not (fam.endswith("_grav")):
            return d[key]
    except Exception:
        pass
    return None


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.proc = None
        self.lines = []
        self.idx = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smmax, reasons = [], None, set()
        for ts, line in self.lines:
            if ts < t0 or ts > t1 + 0.2:
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[1]))
                smmax = float(f[2])
            except Exception:
                continue
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smmax, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
def setup_case(args, rank, world):
    """Build ek_params + the local initial state through the C++ host layer (same ini semantics as
    the reference) for this rank's slab."""
    from euler_kokkos_b200 import hostapi

    w = WORKLOADS[args.workload]
    overrides = {}
    if w["n"] is not None:
        n = args.n or w["n"]
        overrides.update({"mesh.nx": n, "mesh.ny": n, "mesh.nz": n})
        if getattr(args, "shape", None):  # per-GPU interior box, e.g. 1024,1024,128 (configs[3] on 8 GPUs)
            sx, sy, sz = (int(v) for v in args.shape.split(","))
            overrides.update({"mesh.nx": sx, "mesh.ny": sy, "mesh.nz": sz})
    overrides["run.noutput"] = 0
    case = hostapi.Case(os.path.join(ROOT, "euler_kokkos_b200", "settings", w["ini"]), overrides, rank=rank, nranks=world)
    return case, w


def run_ours(args):
    import torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch N>1 with: python -m torch.distributed.run --nproc-per-node N bench.py --gpus N ...")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from euler_kokkos_b200 import capi

    case, w = setup_case(args, rank, world)
    p = case.params
    ctx = capi.Context(p, local)
    stream = torch.cuda.Stream(device=local)
    ctx.set_stream(stream.cuda_stream)
    if world > 1:
        uid = torch.zeros(capi.COMM_ID_BYTES, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(capi.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        ctx.comm_attach(bytes(uid.cpu().numpy().tobytes()), rank, world, case.neighbors)

    # initial state in PINNED host memory (the e2e leg copies it in every repetition)
    nbytes = int(np.prod(p.shape)) * 8
    pinned = torch.empty(int(np.prod(p.shape)), dtype=torch.float64, pin_memory=True)
    U0 = pinned.numpy().reshape(p.shape)
    case.init_state(U0)
    grav_pinned = None
    if p.gravity_enabled:
        grav_pinned = torch.empty(p.dim * int(np.prod(p.shape[1:])), dtype=torch.float64, pin_memory=True)
        case.init_gravity(grav_pinned.numpy().reshape((p.dim,) + tuple(p.shape[1:])))
        ctx.set_gravity(grav_pinned.data_ptr())
    ctx.upload(pinned.data_ptr())

    cells_local = p.nx * p.ny * p.nz
    cells_total = cells_local * world

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing: W warm-up steps, then exactly K steps ---------------------------
    ctx.run(args.warmup)
    ctx.kernel_timing(True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = ctx.launch_count
    t0 = time.time()
    with torch.cuda.stream(stream):
        e0.record(stream)
        t_sim, dt_last, ndone = ctx.run(args.steps)
        e1.record(stream)
    barrier()
    t1 = time.time()
    launches = ctx.launch_count - l0
    ms = e0.elapsed_time(e1)
    kms, kn = ctx.kernel_time()
    ctx.kernel_timing(False)
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    if dist is not None:
        tms = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms.item())
    assert ndone == args.steps, (ndone, args.steps)
    value = cells_total * args.steps / (ms * 1e-3)

    # ---- end-to-end leg through the public C ABI with HOST buffers --------------------------------
    # one job = upload the initial state from pinned host memory, K host-driven steps exactly like the
    # reference's main loop (dt comes back to the host every step), download the final state.
    out_pinned = torch.empty(pinned.numel(), dtype=torch.float64, pin_memory=True)  # empty_like would not be pinned
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e2.record(stream)
        ctx.upload(pinned.data_ptr())
        t = 0.0
        for _ in range(args.steps):
            if p.mhd:
                ctx.fill_boundaries()
            dt = ctx.compute_dt()
            ctx.step(dt)
            t += dt
        ctx.download(out_pinned.data_ptr())
        e3.record(stream)
    barrier()
    ms_e2e = e2.elapsed_time(e3)
    if dist is not None:
        tms = torch.tensor([ms_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms_e2e = float(tms.item())
    e2e_value = cells_total * args.steps / (ms_e2e * 1e-3)
    finite = bool(np.isfinite(out_pinned.numpy()[:: max(1, out_pinned.numel() // 1000003)]).all())

    if rank == 0:
        peak, peak_src = peaks()
        bpc = ALGO_BYTES[w["fam"]]
        k_avg_ms = kms / max(kn, 1)
        achieved = bpc * cells_local / (k_avg_ms * 1e-3) / 1e9 if kn else None
        rec = recorded_traffic(w["fam"], p)
        line = {
            "metric": "cell-updates/s (FP64)", "value": value, "unit": "cell-updates/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic (analytic initial condition of the ini file)",
            "config": {"workload": f"{w['ini']} @ {p.nx}x{p.ny}x{p.nz} per GPU ({case.describe()})", "cells_per_gpu": cells_local,
                       "l2": "inputs (2 x %.1f GB state) larger than L2; no flush needed" % (nbytes / 1e9),
                       "decomposition": "1 z-slab per GPU, NCCL halos" if world > 1 else "single domain"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                         "traffic": (rec["traffic_bytes"] / 1e9) if rec else None, "traffic_unit": "GB per launch (ncu dram bytes read+write)",
                         "traffic_source": rec["source"] if rec else None, "algorithmic_gb_per_launch": bpc * cells_local / 1e9,
                         "fp64_pipe_active_frac_ncu": rec.get("fp64_pipe_active_frac") if rec else None,
                         "fp64_peak_thread_inst_per_s": 18.4e12, "peak_source": peak_src, "kernel": case.kernel_name(), "kernel_ms": k_avg_ms,
                         "kernel_share_of_step": (kms / ms) if ms else None, "algorithmic_bytes_per_cell": bpc},
            "e2e": {"value": e2e_value, "unit": "cell-updates/s", "h2d_bytes_per_step": nbytes / args.steps,
                    "d2h_bytes_per_step": nbytes / args.steps + 8, "what": "upload(pinned) + K x {compute_dt -> host, step} + download(pinned)",
                    "finite": finite},
            "gpu_launches": int(launches), "clocks": clocks, "t_sim": t_sim, "dt_last": dt_last,
        }
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(args)
        emit(line)
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
def reference_sample(args, n_sample, nsteps, threads=None):
    """Run the UNMODIFIED reference binary (oracle/_ref) on a bounded sample of the workload.
    Returns (interior cell-updates/s, description)."""
    from oracle import oracle_lib as O
    from euler_kokkos_b200 import hostapi

    w = WORKLOADS[args.workload]
    if not O.have_reference():
        raise RuntimeError("oracle/_ref/euler_kokkos_ref is missing (build it with __graft_entry__.build() where /root/reference exists)")
    threads = threads or os.cpu_count() or 1
    with tempfile.TemporaryDirectory() as td:
        overrides = {"run.noutput": 0, "run.nstepmax": nsteps, "run.tend": 1e6, "run.nlog": 1000000, "output.outputdir": td}
        if w["n"] is not None:
            overrides.update({"mesh.nx": n_sample, "mesh.ny": n_sample, "mesh.nz": n_sample})
        ini = os.path.join(td, "sample.ini")
        hostapi.write_ini(os.path.join(ROOT, "euler_kokkos_b200", "settings", w["ini"]), overrides, ini)
        out = O.run_reference(ini, td, threads=threads)
    m = re.search(r"Perf\s*:\s*([0-9.eE+-]+)\s*number of Mcell-updates/s", out)
    tt = re.search(r"total\s+time\s*:\s*([0-9.eE+-]+)", out)
    if not m:
        raise RuntimeError("could not parse the reference's Perf line:\n" + out[-1500:])
    cfg = hostapi.Case(os.path.join(ROOT, "euler_kokkos_b200", "settings", w["ini"]),
                       {k: v for k, v in overrides.items() if k.startswith("mesh.")}).params
    ghosted = (cfg.nx + 2 * cfg.ghost) * (cfg.ny + 2 * cfg.ghost) * ((cfg.nz + 2 * cfg.ghost) if cfg.dim == 3 else 1)
    interior = cfg.nx * cfg.ny * cfg.nz
    perf_interior = float(m.group(1)) * 1e6 * interior / ghosted
    desc = (f"{w['ini']} @ {cfg.nx}x{cfg.ny}x{cfg.nz}, {nsteps} steps, reference OpenMP build, {threads} threads, "
            f"total {tt.group(1) if tt else '?'} s; its own Perf line converted to interior cells")
    return perf_interior, desc, threads


def cpu_baseline(args):
    w = WORKLOADS[args.workload]
    n_sample = {"hydro3d": 160, "hydro3d_grav": 160, "mhd": 96, "hydro2d": None, "mhd2d": 1024}[w["fam"]]
    nsteps = {"hydro3d": 6, "hydro3d_grav": 6, "mhd": 4, "hydro2d": 40, "mhd2d": 10}[w["fam"]]
    try:
        v, desc, threads = reference_sample(args, n_sample, nsteps)
        return {"value": v, "unit": "cell-updates/s", "cores": threads, "kind": "reference", "sample": desc}
    except Exception as e:  # noqa: BLE001
        return {"value": None, "unit": "cell-updates/s", "cores": 0, "kind": "reference", "sample": f"failed: {e}"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = WORKLOADS[args.workload]
    n_sample = {"hydro3d": 128, "hydro3d_grav": 128, "mhd": 80, "hydro2d": None, "mhd2d": 768}[w["fam"]]
    nsteps = {"hydro3d": 3, "hydro3d_grav": 3, "mhd": 2, "hydro2d": 20, "mhd2d": 5}[w["fam"]]
    vals, desc, threads = [], "", 0
    t_all = time.time()
    for i in range(args.warmup + args.steps):
        v, desc, threads = reference_sample(args, n_sample, nsteps)
        if i >= args.warmup:
            vals.append(v)
    value = float(np.mean(vals))
    ms = 1e3 * (time.time() - t_all) / max(1, args.warmup + args.steps)
    line = {
        "impl": "reference", "metric": "cell-updates/s (FP64)", "value": value, "unit": "cell-updates/s",
        "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic (analytic initial condition of the ini file)",
        "config": {"workload": f"{w['ini']} (bounded sample per step: {desc})"},
        "cpu_baseline": {"value": value, "unit": "cell-updates/s", "cores": threads, "kind": "reference", "sample": desc},
        "e2e": {"value": value, "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


_REAL_STDOUT = None


def protect_stdout():
    """Everything that libraries print to fd 1 (NCCL's version line, C++ banners) goes to stderr; the
    ONE JSON line is written to the real stdout by emit()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="blast3d", choices=sorted(WORKLOADS))
    ap.add_argument("--n", "--size", dest="n", type=int, default=None, help="cubic interior size per GPU (default 512); use --size under torchrun, whose own parser claims --n")
    ap.add_argument("--shape", default=None, help="per-GPU interior box NX,NY,NZ of a 3-D workload (overrides --n)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    protect_stdout()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
