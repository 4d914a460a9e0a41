#!/usr/bin/env python
"""profiles/traffic.json from the ncu metrics pass of profiles/gpu.sh (`metrics:REGEX`): per-step DRAM traffic, FP64-pipe
instruction count and utilisation of the step kernels, stamped with the hash of the kernel sources they were measured on
(bench.py only quotes them while that hash is the one of the tree it runs from).

    python profiles/make_traffic.py mhd@512 gpurun_out/TAG_metrics.csv "source text" [kernel-name-substring ...]
"""
import csv
import json
import os
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    key, path, source = sys.argv[1:4]
    want = sys.argv[4:]
    with open(path, newline="") as f:
        rows = list(csv.DictReader(ln for ln in f if ln.startswith('"')))
    by = OrderedDict()
    for r in rows:
        by.setdefault((r["ID"], r["Kernel Name"]), {})[r["Metric Name"]] = (float(r["Metric Value"].replace(",", "")), r["Metric Unit"])
    # one step = the LAST launch of each wanted kernel
    last = OrderedDict()
    for (i, name), m in by.items():
        short = name.split("(")[0]
        if not want or any(w in short for w in want):
            last[short] = m
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    tscale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
    tot_bytes = tot_fp64 = tot_ms = pipe_ms = 0.0
    kernels = {}
    for name, m in last.items():
        b = sum(m[k][0] * scale[m[k][1]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        ms = m["gpu__time_duration.sum"][0] * tscale[m["gpu__time_duration.sum"][1]]
        fp64 = m["sm__inst_executed_pipe_fp64.sum"][0] * 32.0
        pipe = m["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"][0] / 100.0
        kernels[name] = {"ms_under_ncu": ms, "dram_gb": b / 1e9, "fp64_thread_inst": fp64, "fp64_pipe_active_frac": pipe}
        tot_bytes += b
        tot_fp64 += fp64
        tot_ms += ms
        pipe_ms += pipe * ms
    n = int(key.split("@")[1])
    cells = n ** 3
    out_path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(out_path) as f:
            d = json.load(f)
    except Exception:
        d = {}
    d["_comment"] = ("per-STEP DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum), FP64-pipe thread instructions per cell and "
                     "time-weighted FP64-pipe utilisation of the step kernels, from the ncu metrics pass named in 'source'; "
                     "bench.py quotes an entry only while 'kernels_sha1' equals the hash of that family's kernel sources (euler_kokkos_b200/csrc/mhd_* or hydro_*, + ek_common.cuh)")
    d[key] = {"traffic_bytes": tot_bytes, "fp64_thread_inst_per_cell": tot_fp64 / cells, "fp64_pipe_active_frac": pipe_ms / tot_ms,
              "kernels": kernels, "kernels_sha1": bench.kernel_sources_hash(key.split("@")[0]), "source": source}
    with open(out_path, "w") as f:
        json.dump(d, f, indent=1)
    print(json.dumps(d[key], indent=1))


if __name__ == "__main__":
    main()
