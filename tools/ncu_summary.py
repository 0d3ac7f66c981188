#!/usr/bin/env python
"""Summarises ncu output for profiles/: a launch list (csv from --metrics gpu__time_duration.sum) into
per-kernel averages and shares, and a .ncu-rep (--set full) into the metrics the roofline quotes.
Usage: python tools/ncu_summary.py launches <launches.csv> | full <prof.ncu-rep>"""
import csv
import subprocess
import sys
from collections import defaultdict

FULL = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    # the match path starts with the bounds / prepare kernels: launches before the first of them belong to the index
    # build (which shares fm_scan_kernel)
    first = next((i for i, r in enumerate(rows) if r[4].startswith(("fm_bounds", "fm_prepare"))), 0)
    rows = rows[first:]
    d = defaultdict(lambda: defaultdict(list))  # kernel -> metric -> values (one row per launch and metric)
    for r in rows:
        d[r[4].split("(")[0]][r[12]].append(float(r[-1].replace(",", "")))
    T = "gpu__time_duration.sum"
    tot = sum(sum(v[T]) for v in d.values())
    extra = sorted({m for v in d.values() for m in v} - {T})
    short = {"smsp__inst_executed.sum": "warp_inst", "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue%",
             "sm__warps_active.avg.pct_of_peak_sustained_active": "warps%"}
    print("# per-kernel device time from `ncu --metrics gpu__time_duration.sum[,...] --clock-control none` (cold-cache, serialised)")
    print("%-32s %4s %10s %7s" % ("kernel", "n", "avg_us", "share") + "".join(" %12s" % short.get(m, m[:12]) for m in extra))
    for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1][T])):
        t = v[T]
        print("%-32s %4d %10.1f %7.3f" % (k[:32], len(t), sum(t) / len(t) / 1e3, sum(t) / tot)
              + "".join(" %12.1f" % (sum(v[m]) / len(v[m])) if v[m] else " %12s" % "-" for m in extra))


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print("# from `ncu --set full --clock-control none --import-source on` (%s)" % path)
    for r in rows[2:]:
        print("kernel: %s" % r[idx["Kernel Name"]].split("(")[0])
        for m in FULL:
            if m in idx:
                print("  %-72s %16s %s" % (m, r[idx[m]], units[idx[m]]))
        rd, wr = float(r[idx["dram__bytes_read.sum"]]), float(r[idx["dram__bytes_write.sum"]])
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
        print("  %-72s %16.1f MB" % ("traffic = dram read + write",
                                      (rd * scale[units[idx["dram__bytes_read.sum"]]] + wr * scale[units[idx["dram__bytes_write.sum"]]]) / 1e6))


def traffic(path):
    """JSON for profiles/ncu_traffic.json: per kernel DRAM bytes, issue-slot utilisation and warp instructions per launch."""
    import json
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
    acc = defaultdict(lambda: defaultdict(list))
    for r in rows[2:]:
        k = r[idx["Kernel Name"]].split("(")[0].split("<")[0]
        rd = float(r[idx["dram__bytes_read.sum"]]) * scale[units[idx["dram__bytes_read.sum"]]]
        wr = float(r[idx["dram__bytes_write.sum"]]) * scale[units[idx["dram__bytes_write.sum"]]]
        acc[k]["traffic_bytes_per_launch"].append(rd + wr)
        acc[k]["issue_active_pct"].append(float(r[idx["smsp__issue_active.avg.pct_of_peak_sustained_active"]]))
        acc[k]["inst_executed"].append(float(r[idx["smsp__inst_executed.sum"]]))
        acc[k]["duration_us"].append(float(r[idx["gpu__time_duration.sum"]]) * (1e-3 if units[idx["gpu__time_duration.sum"]] == "ns" else 1.0))
        acc[k]["l2_hit_pct"].append(float(r[idx["lts__t_sector_hit_rate.pct"]]))
    res = {k: dict({m: sum(v) / len(v) for m, v in d.items()}, launches_sampled=len(d["inst_executed"])) for k, d in acc.items()}
    res["_source"] = "ncu --set full --clock-control none (%s): per launch, config 2 (1M sentences, 100k queries/step)" % path
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic}[sys.argv[1]](sys.argv[2])
