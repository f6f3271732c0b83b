#!/usr/bin/env python
"""Turns gpurun_out/ ncu artefacts into the small tracked summaries under profiles/.
usage: python profiles/summarize.py <tag> [--launches gpurun_out/launches.csv] [--rep gpurun_out/prof_cb.ncu-rep]"""
import argparse
import collections
import csv
import json
import os
import subprocess

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_issued.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__waves_per_multiprocessor", "smsp__thread_inst_executed_per_inst_executed.ratio", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg", "smsp__inst_executed.sum", "sm__inst_executed.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum",
        "sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tmem.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_bytes.sum", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct"]


def launches(path, out):
    rows = [r for r in csv.reader(open(path)) if r and r[0].isdigit()]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        agg[r[4]][0] += 1; agg[r[4]][1] += float(r[-1])
    tot = sum(v[1] for v in agg.values())
    with open(out, "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare shares)\n")
        f.write("kernel,launches,total_us,avg_us,share_pct\n")
        for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"\"{k}\",{v[0]},{v[1] / 1e3:.1f},{v[1] / v[0] / 1e3:.2f},{100 * v[1] / tot:.1f}\n")


def full(rep, out, traffic_json=None):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")], "grid": r[hdr.index("Grid Size")], "block": r[hdr.index("Block Size")]}
        for k in KEYS:
            if k in hdr:
                d[k] = f"{r[hdr.index(k)]} {units[hdr.index(k)]}".strip()
        res.append(d)
    with open(out, "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on (per launch)\n")
        json.dump(res, f, indent=1)
    if traffic_json and res:
        def b(s):
            v, u = s.split()[:2]
            return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        t = b(res[0]["dram__bytes_read.sum"]) + b(res[0]["dram__bytes_write.sum"])
        json.dump({"kernel": res[0]["kernel"], "dram_bytes_per_launch": t, "source": os.path.basename(out)}, open(traffic_json, "w"))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("tag")
    ap.add_argument("--launches"); ap.add_argument("--rep"); ap.add_argument("--traffic", action="store_true")
    a = ap.parse_args()
    here = os.path.dirname(os.path.abspath(__file__))
    if a.launches:
        launches(a.launches, os.path.join(here, f"{a.tag}_launches.csv"))
    if a.rep:
        full(a.rep, os.path.join(here, f"{a.tag}_ncu_full.json"), os.path.join(here, "checkerboard_traffic.json") if a.traffic else None)
