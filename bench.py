#!/usr/bin/env python
"""bench.py -- cell-updates/s (FP64) of the MUSCL-Hancock time step on B200.

    python bench.py --gpus N --steps K --warmup W [--workload ot3d|blast3d|implode2d|rt3d|ot2d] [--n 512]
    python bench.py --impl reference ...      # the reference's own CPU (OpenMP) build, same metric and config
    python bench.py --impl reference-cuda ... # the reference's own Kokkos-Cuda build on one B200 (same-box GPU comparator)

A "step" is one godunov_unsplit over the whole (local) grid.  The default workload is the north-star target,
BASELINE.json configs[2]: test_mhd_orszag_tang_3D.ini scaled to 512^3 (MHD_Muscl_3D, HLLD + constrained transport).
N>1 (one process per GPU under torchrun) stacks one 512^3 slab per GPU along z (weak scaling); ghost planes travel
over NCCL/NVLink.  Rank 0 prints ONE JSON line; the hydro headline (configs[1], test_blast_3D.ini @ 512^3, HLLC) is
measured in the same run and reported under "secondary".
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import re
import subprocess
import sys
import tempfile
import threading
import time
from collections import OrderedDict

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
os.environ.setdefault("EKH_QUIET", "1")  # stdout carries exactly one JSON line
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SETTINGS = os.path.join(ROOT, "euler_kokkos_b200", "settings")
ALGO_BYTES = {"hydro2d": 64, "hydro3d": 80, "hydro3d_grav": 104, "mhd": 128, "mhd2d": 128}
FP64_PEAK = 18.4e12  # thread-level FP64-pipe instructions/s at 1.965 GHz (profiles/microbench/fp64_peak.cu)

WORKLOADS = {
    # name: (ini file in euler_kokkos_b200/settings, solver family, default n)
    "blast3d": dict(ini="test_blast_3D.ini", fam="hydro3d", n=512),
    "implode3d": dict(ini="test_implode_3D.ini", fam="hydro3d", n=512),
    "ot3d": dict(ini="test_mhd_orszag_tang_3D.ini", fam="mhd", n=512),
    "rt3d": dict(ini="test_rayleigh_taylor_3d.ini", fam="hydro3d_grav", n=512),
    "implode2d": dict(ini="test_implode_2D.ini", fam="hydro2d", n=None),
    "ot2d": dict(ini="test_mhd_orszag_tang_2D.ini", fam="mhd2d", n=8192),
}


# ---- ini files, in plain Python (the reference arm must not load the product libraries) ---------------------------
class Ini:
    """The few ConfigMap semantics bench.py needs: case-insensitive section/key lookup, overrides, write-back."""

    def __init__(self, path):
        self.sections = OrderedDict()
        sec = None
        with open(path) as f:
            for raw in f:
                line = raw.strip()
                if not line or line[0] in "#;":
                    continue
                m = re.match(r"\[(.+)\]$", line)
                if m:
                    sec = self.sections.setdefault(m.group(1).strip().lower(), OrderedDict())
                    continue
                if "=" in line and sec is not None:
                    k, v = line.split("=", 1)
                    v = re.split(r"\s[#;]", v, 1)[0]
                    sec[k.strip().lower()] = v.strip()

    def set(self, dotted, value):
        s, k = dotted.lower().split(".", 1)
        self.sections.setdefault(s, OrderedDict())[k] = str(value)

    def get(self, s, k, default=None):
        return self.sections.get(s.lower(), {}).get(k.lower(), default)

    def write(self, path):
        with open(path, "w") as f:
            for s, kv in self.sections.items():
                f.write(f"[{s}]\n")
                for k, v in kv.items():
                    f.write(f"{k}={v}\n")
                f.write("\n")


def workload_ini(args, extra=None):
    """the ini of the workload with the size overrides of this run (what both arms are configured from)"""
    w = WORKLOADS[args.workload]
    ini = Ini(os.path.join(SETTINGS, w["ini"]))
    if w["n"] is not None:
        n = args.n or w["n"]
        sx = sy = sz = n
        if getattr(args, "shape", None):  # per-GPU interior box, e.g. 1024,1024,128 (configs[3] on 8 GPUs)
            sx, sy, sz = (int(v) for v in args.shape.split(","))
        ini.set("mesh.nx", sx)
        ini.set("mesh.ny", sy)
        ini.set("mesh.nz", sz)
    ini.set("run.noutput", 0)
    for k, v in (extra or {}).items():
        ini.set(k, v)
    return ini, w


def ini_geometry(ini):
    name = ini.get("run", "solver_name", "Unknown")
    dim = 3 if name.endswith("3D") else 2
    gw = 3 if name.startswith("MHD") else 2
    nx, ny = int(ini.get("mesh", "nx")), int(ini.get("mesh", "ny"))
    nz = int(ini.get("mesh", "nz", 1)) if dim == 3 else 1
    return name, dim, gw, nx, ny, nz


def config_dict(args, world):
    """identical for both arms: the workload the metric is quoted on"""
    ini, w = workload_ini(args)
    name, dim, gw, nx, ny, nz = ini_geometry(ini)
    desc = (f"{name}, riemann={ini.get('hydro', 'riemann', 'approx')}, impl={ini.get('other', 'implementationversion', 0)}, "
            f"problem={ini.get('hydro', 'problem', 'unknown')}")
    nbvar = 8 if name.startswith("MHD") else (5 if dim == 3 else 4)
    state_gb = nbvar * 8 * (nx + 2 * gw) * (ny + 2 * gw) * ((nz + 2 * gw) if dim == 3 else 1) / 1e9
    return {"workload": f"{w['ini']} @ {nx}x{ny}x{nz} per GPU ({desc})", "cells_per_gpu": nx * ny * nz,
            "l2": "inputs (2 x %.1f GB state) larger than L2; no flush needed" % state_gb,
            "decomposition": "1 z-slab per GPU, NCCL halos" if world > 1 else "single domain"}


def kernel_sources_hash(fam="mhd"):
    """identifies the kernels a recorded ncu capture belongs to (profiles/traffic.json): the sources of that solver
    family's kernels (mhd_*.cu* or hydro_*.cu*) + the shared arithmetic header"""
    h = hashlib.sha1()
    d = os.path.join(ROOT, "euler_kokkos_b200", "csrc")
    prefix = "mhd_" if fam.startswith("mhd") else "hydro_"
    for fn in sorted(os.listdir(d)):
        if fn.endswith((".cu", ".cuh")) and (fn.startswith(prefix) or fn == "ek_common.cuh"):
            with open(os.path.join(d, fn), "rb") as f:
                h.update(fn.encode() + b"\0" + f.read())
    return h.hexdigest()[:12]


def recorded_capture(fam, nx, ny, nz):
    """per-launch DRAM traffic / instruction counts of the step kernels from the committed ncu capture
    (profiles/traffic.json), only when it was taken on the kernels that are built now and on this size"""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            d = json.load(f)
        key = f"{fam.replace('_grav', '')}@{nx}"
        rec = d.get(key)
        if rec and nx == ny == nz and not fam.endswith("_grav") and rec.get("kernels_sha1") == kernel_sources_hash(fam):
            return rec
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


def bind_near_gpu(local):
    """Pin this rank (and the pinned staging buffers it first-touches) to the CPUs of the GPU's NUMA node: with every rank
    left on node 0, eight concurrent 9 GB uploads share one memory controller / PCIe root.  Best effort."""
    try:
        import torch

        bus = torch.cuda.get_device_properties(local).pci_bus_id
        dom = torch.cuda.get_device_properties(local).pci_domain_id
        dev = torch.cuda.get_device_properties(local).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/local_cpulist"
        with open(path) as f:
            text = f.read().strip()
        cpus = set()
        for part in text.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
            return f"{len(allowed)} cpus of the GPU's NUMA node ({text})"
        return f"not rebound (GPU-local cpus {text} outside this process's cpuset)"
    except Exception as e:  # noqa: BLE001
        return f"not rebound ({type(e).__name__})"


# ------------------------------------------------------------------------------------------------
def measure(args, workload, steps, warmup, rank, world, local, dist, stream, with_e2e, sampler=None):
    """device-resident timing (+ the end-to-end leg) of one workload; returns a dict (rank 0) or None"""
    import torch

    from euler_kokkos_b200 import capi, hostapi

    wargs = argparse.Namespace(**vars(args))
    wargs.workload = workload
    if workload != args.workload:
        wargs.n, wargs.shape = None, None
    ini, w = workload_ini(wargs)
    with tempfile.NamedTemporaryFile("w", suffix=".ini", delete=False) as f:
        tmp_ini = f.name
    ini.write(tmp_ini)
    try:
        case = hostapi.Case(tmp_ini, None, rank=rank, nranks=world)
    finally:
        os.unlink(tmp_ini)
    p = case.params
    ctx = capi.Context(p, local)
    ctx.set_stream(stream.cuda_stream)
    if world > 1:
        uid = torch.zeros(capi.COMM_ID_BYTES, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(capi.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        ctx.comm_attach(bytes(uid.cpu().numpy().tobytes()), rank, world, case.neighbors)

    # initial state in PINNED host memory (the e2e leg copies it in every repetition)
    nelem = int(np.prod(p.shape))
    nbytes = nelem * 8
    pinned = torch.empty(nelem, dtype=torch.float64, pin_memory=True)
    case.init_state(pinned.numpy().reshape(p.shape))
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

    def max_over_ranks(ms):
        if dist is None:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident timing: W warm-up steps, then exactly K steps ---------------------------
    ctx.run(warmup)
    ctx.kernel_timing(True)
    if sampler is not None:
        sampler.start()
        time.sleep(0.3)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = ctx.launch_count
    t0 = time.time()
    with torch.cuda.stream(stream):
        e0.record(stream)
        t_sim, dt_last, ndone = ctx.run(steps)
        e1.record(stream)
    barrier()
    t1 = time.time()
    launches = ctx.launch_count - l0
    ms = max_over_ranks(e0.elapsed_time(e1))
    kms, kn = ctx.kernel_time()
    ctx.kernel_timing(False)
    clocks = sampler.stop(t0, t1) if sampler is not None else None
    assert ndone == steps, (ndone, steps)
    value = cells_total * steps / (ms * 1e-3)

    res = None
    e2e = None
    if with_e2e:
        # ---- end-to-end leg through the public C ABI with HOST buffers --------------------------------
        # one job = upload the initial state from pinned host memory, K host-driven steps exactly like the
        # reference's main loop (dt comes back to the host every step: compute_dt -> step, SolverBase::next_iteration),
        # download the final state.  The stepping part is also timed on its own.
        out_pinned = torch.empty(nelem, dtype=torch.float64, pin_memory=True)  # empty_like would not be pinned
        barrier()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        with torch.cuda.stream(stream):
            ev[0].record(stream)
            ctx.upload(pinned.data_ptr())
            ev[1].record(stream)
            for _ in range(steps):
                dt = ctx.compute_dt()
                ctx.step(dt)
            ev[2].record(stream)
            ctx.download(out_pinned.data_ptr())
            ev[3].record(stream)
        barrier()
        ms_e2e = max_over_ranks(ev[0].elapsed_time(ev[3]))
        ms_up = max_over_ranks(ev[0].elapsed_time(ev[1]))
        ms_steps = max_over_ranks(ev[1].elapsed_time(ev[2]))
        ms_down = max_over_ranks(ev[2].elapsed_time(ev[3]))
        finite = bool(np.isfinite(out_pinned.numpy()[:: max(1, nelem // 1000003)]).all())
        e2e = {"value": cells_total * steps / (ms_e2e * 1e-3), "unit": "cell-updates/s",
               "h2d_bytes_per_step": nbytes / steps, "d2h_bytes_per_step": nbytes / steps + 8,
               "h2d_bytes_total": nbytes, "d2h_bytes_total": nbytes + 8 * steps,
               "what": "upload(pinned) + K x {compute_dt -> host, step} + download(pinned); the state crosses PCIe once "
                       "per job in each direction, so the per-step byte counts are the totals divided by K",
               "upload_ms": ms_up, "stepping_ms": ms_steps, "download_ms": ms_down,
               "stepping_only_value": cells_total * steps / (ms_steps * 1e-3), "finite": finite}
        del out_pinned

    if rank == 0:
        peak, peak_src = peaks()
        bpc = ALGO_BYTES[w["fam"]]
        k_avg_ms = kms / max(kn, 1)
        achieved = bpc * cells_local / (k_avg_ms * 1e-3) / 1e9 if kn else None
        rec = recorded_capture(w["fam"], p.nx, p.ny, p.nz)
        inst_per_cell = rec.get("fp64_thread_inst_per_cell") if rec else None
        fp64_frac = (inst_per_cell * cells_local / (k_avg_ms * 1e-3) / FP64_PEAK) if (inst_per_cell and kn) else None
        cargs = argparse.Namespace(**vars(wargs))
        res = {
            "value": value, "ms_per_step": ms / steps, "steps": steps, "warmup": warmup,
            "config": config_dict(cargs, world),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                         "traffic": (rec["traffic_bytes"] / 1e9) if rec else None, "traffic_unit": "GB per step (ncu dram bytes read+write)",
                         "traffic_source": rec["source"] if rec else "no ncu capture of the kernels built now (profiles/traffic.json is keyed by a hash of csrc/)",
                         "algorithmic_gb_per_launch": bpc * cells_local / 1e9, "algorithmic_bytes_per_cell": bpc,
                         "fp64_frac": fp64_frac, "fp64_thread_inst_per_cell_ncu": inst_per_cell,
                         "fp64_peak_thread_inst_per_s": FP64_PEAK,
                         "fp64_note": "fp64_frac = FP64-pipe thread instructions per cell (ncu) x cells/s / peak: the binding roofline of this path",
                         "peak_source": peak_src, "kernel": case.kernel_name(), "kernel_ms": k_avg_ms,
                         "kernel_share_of_step": (kms / ms) if ms else None},
            "gpu_launches": int(launches), "clocks": clocks, "t_sim": t_sim, "dt_last": dt_last,
        }
        if e2e:
            res["e2e"] = e2e
    ctx.close()
    del pinned
    torch.cuda.empty_cache()
    return res


def run_ours(args):
    import torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch N>1 with: python -m torch.distributed.run --nproc-per-node N bench.py --gpus N ...")
    torch.cuda.set_device(local)
    affinity = bind_near_gpu(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stream = torch.cuda.Stream(device=local)
    sampler = ClockSampler(local) if rank == 0 else None

    main_res = measure(args, args.workload, args.steps, args.warmup, rank, world, local, dist, stream, True, sampler)
    secondary = None
    if not args.no_secondary and args.workload == "ot3d" and not args.n and not args.shape:
        # the hydro headline (BASELINE.json configs[1]) in the same run: 512^3 blast, HLLC, device-resident timing
        sec = measure(args, "blast3d", max(10, args.steps // 2), 3, rank, world, local, dist, stream, False)
        if rank == 0:
            secondary = {"metric": "cell-updates/s (FP64)", "unit": "cell-updates/s", "value": sec["value"],
                         "ms_per_step": sec["ms_per_step"], "steps": sec["steps"], "warmup": sec["warmup"],
                         "config": sec["config"], "roofline": sec["roofline"], "gpu_launches": sec["gpu_launches"]}

    if rank == 0:
        line = {
            "metric": "cell-updates/s (FP64)", "value": main_res["value"], "unit": "cell-updates/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": main_res["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic (analytic initial condition of the ini file)",
            "config": main_res["config"], "roofline": main_res["roofline"], "e2e": main_res["e2e"],
            "gpu_launches": main_res["gpu_launches"], "clocks": main_res["clocks"], "t_sim": main_res["t_sim"],
            "dt_last": main_res["dt_last"], "host_affinity": affinity,
        }
        if secondary:
            line["secondary"] = secondary
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(args)
        emit(line)
    if dist is not None:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
def reference_sample(args, n_sample, nsteps, threads=None):
    """Run the UNMODIFIED reference binary (oracle/_ref) on a bounded sample of the workload.
    Returns (interior cell-updates/s, description, threads).  Loads none of the product libraries: the ini is edited
    in plain Python."""
    from oracle import oracle_lib as O

    w = WORKLOADS[args.workload]
    if not O.have_reference():
        raise RuntimeError("oracle/_ref/euler_kokkos_ref is missing (build it with __graft_entry__.build() where /root/reference exists)")
    threads = threads or os.cpu_count() or 1
    with tempfile.TemporaryDirectory() as td:
        extra = {"run.nstepmax": nsteps, "run.tend": 1e6, "run.nlog": 1000000, "output.outputdir": td}
        if w["n"] is not None:
            extra.update({"mesh.nx": n_sample, "mesh.ny": n_sample, "mesh.nz": n_sample})
        ini, _ = workload_ini(args, extra)
        path = os.path.join(td, "sample.ini")
        ini.write(path)
        out = O.run_reference(path, td, threads=threads)
    m = re.search(r"Perf\s*:\s*([0-9.eE+-]+)\s*number of Mcell-updates/s", out)
    tt = re.search(r"total\s+time\s*:\s*([0-9.eE+-]+)", out)
    if not m:
        raise RuntimeError("could not parse the reference's Perf line:\n" + out[-1500:])
    _, dim, gw, nx, ny, nz = ini_geometry(ini)
    ghosted = (nx + 2 * gw) * (ny + 2 * gw) * ((nz + 2 * gw) if dim == 3 else 1)
    perf_interior = float(m.group(1)) * 1e6 * (nx * ny * nz) / ghosted
    desc = (f"{w['ini']} @ {nx}x{ny}x{nz}, {nsteps} steps, reference OpenMP build, {threads} threads "
            f"(OMP_PROC_BIND=spread OMP_PLACES=threads), total {tt.group(1) if tt else '?'} s; its own Perf line "
            f"converted to interior cells")
    return perf_interior, desc, threads


def cpu_baseline(args):
    """the reference's OpenMP build on the box's host cores: one bounded sample of the workload (for MHD the largest
    cubic size the reference's ~1.7 KB/cell of work arrays allows in a few tens of seconds: 256^3)"""
    w = WORKLOADS[args.workload]
    n_sample = {"hydro3d": 256, "hydro3d_grav": 256, "mhd": 256, "hydro2d": None, "mhd2d": 2048}[w["fam"]]
    nsteps = {"hydro3d": 6, "hydro3d_grav": 6, "mhd": 3, "hydro2d": 40, "mhd2d": 10}[w["fam"]]
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
    # 3-D: 256^3 (the reference's MHD work arrays take ~1.7 KB per cell: 29 GB of host memory; 3-4 s per sample on 16 cores)
    n_sample = {"hydro3d": 256, "hydro3d_grav": 256, "mhd": 256, "hydro2d": None, "mhd2d": 1024}[w["fam"]]
    if getattr(args, "ref_sample_n", None) and n_sample is not None:
        n_sample = args.ref_sample_n
    nsteps = {"hydro3d": 3, "hydro3d_grav": 3, "mhd": 2, "hydro2d": 20, "mhd2d": 5}[w["fam"]]
    vals, desc, threads = [], "", 0
    t_all = time.time()
    for i in range(args.warmup + args.steps):
        v, desc, threads = reference_sample(args, n_sample, nsteps)
        if i >= args.warmup:
            vals.append(v)
    value = float(np.mean(vals))
    ms = 1e3 * (time.time() - t_all) / max(1, args.warmup + args.steps)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    line = {
        "impl": "reference", "metric": "cell-updates/s (FP64)", "value": value, "unit": "cell-updates/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic (analytic initial condition of the ini file)",
        "config": config_dict(args, world),
        "cpu_baseline": {"value": value, "unit": "cell-updates/s", "cores": threads, "kind": "reference",
                         "sample": "each step = one run of: " + desc},
        "e2e": {"value": value, "unit": "cell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def run_reference_cuda_arm(args):
    """The reference's OWN GPU build (oracle/_ref_cuda: unmodified sources, Kokkos Cuda backend, -arch=sm_100) on one
    B200 of the same box: the GPU-vs-GPU comparator.  One run of the binary on the workload at the largest size its
    work arrays allow (MHD impl 0 needs ~1.66 KB per cell: 512^3 does not fit 180 GB, 384^3 does); the figure is its own
    Perf line converted to interior cells.  None of the product libraries is loaded."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    binary = os.path.join(ROOT, "oracle", "_ref_cuda", "euler_kokkos_cuda")
    if not os.path.exists(binary):
        emit({"impl": "reference-cuda", "unavailable": "oracle/_ref_cuda/euler_kokkos_cuda is not built (make -C oracle/ref_build -f Makefile.cuda)"})
        return
    w = WORKLOADS[args.workload]
    n = args.n or {"mhd": 384}.get(w["fam"], w["n"])
    nsteps = args.warmup + args.steps
    with tempfile.TemporaryDirectory() as td:
        extra = {"run.nstepmax": nsteps, "run.tend": 1e6, "run.nlog": 1000000, "output.outputdir": td}
        if w["n"] is not None:
            extra.update({"mesh.nx": n, "mesh.ny": n, "mesh.nz": n})
        ini, _ = workload_ini(args, extra)
        path = os.path.join(td, "case.ini")
        ini.write(path)
        t0 = time.time()
        r = subprocess.run([binary, path], cwd=td, capture_output=True, text=True, timeout=3000)
        wall = time.time() - t0
    out = r.stdout
    m = re.search(r"Perf\s*:\s*([0-9.eE+-]+)\s*number of Mcell-updates/s", out)
    if r.returncode != 0 or not m:
        emit({"impl": "reference-cuda", "unavailable": f"run failed (rc={r.returncode}): {(out + r.stderr)[-400:]}"})
        return
    _, dim, gw, nx, ny, nz = ini_geometry(ini)
    ghosted = (nx + 2 * gw) * (ny + 2 * gw) * ((nz + 2 * gw) if dim == 3 else 1)
    value = float(m.group(1)) * 1e6 * (nx * ny * nz) / ghosted
    tt = re.search(r"total\s+time\s*:\s*([0-9.eE+-]+)", out)
    cargs = argparse.Namespace(**vars(args))
    cargs.n = n
    emit({"impl": "reference-cuda", "metric": "cell-updates/s (FP64)", "value": value, "unit": "cell-updates/s", "n_gpus": 1,
          "steps": nsteps, "warmup": 0, "ms_per_step": (float(tt.group(1)) * 1e3 / nsteps) if tt else None,
          "higher_is_better": True, "dtype": "f64", "data": "synthetic (analytic initial condition of the ini file)",
          "config": config_dict(cargs, 1), "wall_s": wall,
          "what": "unmodified reference, Kokkos 5.1 Cuda backend, nvcc -arch=sm_100 (oracle/ref_build/Makefile.cuda); its own Perf "
                  "line (all steps incl. the first, total-time timer) converted to interior cells"})


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
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference-cuda"])
    ap.add_argument("--workload", default="ot3d", choices=sorted(WORKLOADS))
    ap.add_argument("--n", "--size", dest="n", type=int, default=None, help="cubic interior size per GPU (default 512); use --size under torchrun, whose own parser claims --n")
    ap.add_argument("--shape", default=None, help="per-GPU interior box NX,NY,NZ of a 3-D workload (overrides --n)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-sample-n", type=int, default=None,
                    help="--impl reference: cubic size of the bounded sample each step runs (default 256 for the 3-D workloads)")
    ap.add_argument("--no-secondary", action="store_true", help="skip the hydro blast 512^3 line reported under 'secondary'")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    protect_stdout()
    if args.impl == "reference":
        run_reference_arm(args)
    elif args.impl == "reference-cuda":
        run_reference_cuda_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
