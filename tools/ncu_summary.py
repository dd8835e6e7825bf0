#!/usr/bin/env python
"""Summarise ncu outputs: `launches <csv>` (gpu__time_duration launch list) or `raw <ncu-rep>...` (selected metrics)."""
import collections
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "smsp__inst_executed.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = row["Kernel Name"].split("(")[0]
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0}.get(row["Metric Unit"], 1e-6)
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    print("| kernel | launches | total ms | share |\n|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"| {k} | {v[0]} | {v[1]:.3f} | {100 * v[1] / tot:.1f} % |")


def raw(paths):
    for p in paths:
        out = subprocess.run(["ncu", "-i", p, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        r = list(csv.reader(out.splitlines()))
        hdr, units, vals = r[0], r[1], r[2]
        print("==", p)
        for i, h in enumerate(hdr):
            if h in WANT:
                print(f"  {h:88s} {vals[i]:>16s} {units[i]}")


if __name__ == "__main__" and sys.argv[1] in ("launches", "raw"):
    (launches(sys.argv[2]) if sys.argv[1] == "launches" else raw(sys.argv[2:]))


def table(paths):
    """One markdown row per capture: ms, DRAM R+W (GB), DRAM %, fp64 %, warps %, issue %, regs, L1 hit, L2 hit, top stalls."""
    print("| capture | ms | DRAM R+W GB | DRAM % | fp64 % | warps % | issue % | regs | L1 hit % | L2 hit % | stalls (per issue) |\n|---|---|---|---|---|---|---|---|---|---|---|")
    for p in paths:
        out = open(p).read() if p.endswith(".csv") else subprocess.run(["ncu", "-i", p, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        r = list(csv.reader(out.splitlines()))
        hdr, units, vals = r[0], r[1], r[2]
        d = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}

        def f(k, scale=1.0):
            v, u = d.get(k, ("nan", ""))
            x = float(v.replace(",", ""))
            m = {"Gbyte": 1.0, "Mbyte": 1e-3, "Kbyte": 1e-6, "byte": 1e-9, "ms": 1.0, "us": 1e-3, "ns": 1e-6}.get(u, 1.0)
            return x * m * scale
        st = {k.split("stalled_")[1].split("_per_issue")[0]: f(k) for k in d if "issue_stalled" in k and k.endswith("per_issue_active.ratio")}
        top = ", ".join(f"{k} {v:.1f}" for k, v in sorted(st.items(), key=lambda x: -x[1])[:3])
        name = p.split("/")[-1].replace(".ncu-rep", "").replace("_raw.csv", "").replace("prof_", "")
        print(f"| {name} | {f('gpu__time_duration.sum'):.2f} | {f('dram__bytes_read.sum'):.2f} + {f('dram__bytes_write.sum'):.2f} | "
              f"{f('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | {f('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active'):.1f} | "
              f"{f('sm__warps_active.avg.pct_of_peak_sustained_active'):.1f} | {f('smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f} | "
              f"{int(f('launch__registers_per_thread'))} | {f('l1tex__t_sector_hit_rate.pct'):.0f} | {f('lts__t_sector_hit_rate.pct'):.0f} | {top} |")


if __name__ == "__main__" and sys.argv[1] == "table":
    table(sys.argv[2:])
