#!/usr/bin/env python
"""Summaries of ncu CSV exports (used by profiles/gpu.sh on the box and here on the merged gpurun_out/ files).
    summarize.py launches FILE.csv   per-kernel totals of a `--metrics gpu__time_duration.sum` launch list
    summarize.py metrics  FILE.csv   per-launch table of the metrics pass of gpu.sh
    summarize.py raw      FILE.csv   selected counters of a `--set full` capture exported with --page raw --csv
"""
import csv
import sys
from collections import OrderedDict, defaultdict


def rows_of(path):
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    return list(csv.DictReader(lines))


def short(name):
    return name.split("(")[0].replace("ek::", "")[:60]


def launches(path):
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in rows_of(path):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        k = short(r["Kernel Name"])
        tot[k] += ms
        cnt[k] += 1
    total = sum(tot.values())
    print(f"{'kernel':60s} {'launches':>8s} {'total ms':>10s} {'ms/launch':>10s} {'share':>6s}")
    for k in sorted(tot, key=tot.get, reverse=True):
        print(f"{k:60s} {cnt[k]:8d} {tot[k]:10.3f} {tot[k] / cnt[k]:10.4f} {100 * tot[k] / total:5.1f}%")
    print(f"{'total':60s} {sum(cnt.values()):8d} {total:10.3f}")


def metrics(path):
    by = OrderedDict()
    for r in rows_of(path):
        by.setdefault((r["ID"], short(r["Kernel Name"])), {})[r["Metric Name"]] = (float(r["Metric Value"].replace(",", "")), r["Metric Unit"])
    for (i, k), m in by.items():
        t = m.get("gpu__time_duration.sum", (0, "ns"))
        ms = t[0] * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(t[1], 1e-6)

        def gb(name):
            v, u = m.get(name, (0, "byte"))
            return v * {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0, "Tbyte": 1e3}.get(u, 1e-9)

        print(f"{i:>4s} {k:44s} {ms:8.3f} ms  dram r/w {gb('dram__bytes_read.sum'):7.2f}/{gb('dram__bytes_write.sum'):7.2f} GB  "
              f"L2 {gb('lts__t_bytes.sum'):8.2f} GB  fp64 pipe {m.get('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', (0,))[0]:5.1f}%  "
              f"warps {m.get('sm__warps_active.avg.pct_of_peak_sustained_active', (0,))[0]:5.1f}%  "
              f"thread inst {m.get('smsp__thread_inst_executed.sum', (0,))[0]:.4g}  warp inst {m.get('smsp__inst_executed.sum', (0,))[0]:.4g}  "
              f"fp64 warp inst {m.get('sm__inst_executed_pipe_fp64.sum', (0,))[0]:.4g}")


RAW_KEYS = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
            "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum",
            "sm__inst_executed_pipe_fp64.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
            "launch__occupancy_limit_shared_mem", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "dram__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
            "smsp__average_warp_latency_issue_stalled_barrier.ratio", "smsp__average_warp_latency_issue_stalled_wait.ratio",
            "smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio", "smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio",
            "smsp__average_warp_latency_issue_stalled_not_selected.ratio", "smsp__average_warp_latency_issue_stalled_no_instruction.ratio",
            "smsp__average_warp_latency_issue_stalled_lg_throttle.ratio", "smsp__average_warp_latency_issue_stalled_mio_throttle.ratio",
            "smsp__average_warp_latency_issue_stalled_dispatch_stall.ratio", "smsp__average_warp_latency_issue_stalled_membar.ratio",
            "smsp__average_warp_latency_issue_stalled_drain.ratio", "smsp__average_warp_latency_issue_stalled_branch_resolving.ratio")


def raw(path):
    with open(path, newline="") as f:
        rd = list(csv.reader(f))
    hdr, units = rd[0], rd[1]
    for row in rd[2:]:
        d = dict(zip(hdr, row))
        print(f"== {d.get('ID')} {short(d.get('Kernel Name', ''))}  grid {d.get('Grid Size')} block {d.get('Block Size')}")
        for k in RAW_KEYS:
            if k in d:
                print(f"   {k:78s} {d[k]:>16s} {units[hdr.index(k)]}")


if __name__ == "__main__":
    {"launches": launches, "metrics": metrics, "raw": raw}[sys.argv[1]](sys.argv[2])
