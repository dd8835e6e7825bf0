#!/usr/bin/env python
"""profiles/traffic.json from ncu --set full raw pages: DRAM bytes and fp64 flops per column, per kernel and per bench stage.

    python tools/make_traffic_json.py <tag> <ncol> [workload-key]     reads gpurun_out/<tag>_<kernel>_raw.csv (or profiles/)
"""
import csv
import glob
import json
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
STAGE_OF = {"gas_prep_kernel": "gas_optics_lw", "gas_col_kernel": "gas_optics_lw", "gas_lw_band_kernel": "gas_optics_lw", "gas_lw_kernel": "gas_optics_lw",
            "aerosol_optics_kernel": "gas_optics_lw", "gas_sw_kernel": "gas_optics_sw", "gas_sw_band_kernel": "gas_optics_sw",
            "sw_incoming_norm_kernel": "gas_optics_sw", "cloud_prep_kernel": "cloud_optics_generator", "cloud_optics_kernel": "cloud_optics_generator",
            "cloud_gen_warp_kernel": "cloud_optics_generator", "lw_down_kernel": "solver_lw", "lw_up_kernel": "solver_lw", "lw_flux_kernel": "solver_lw",
            "lw_scan_kernel": "solver_lw", "sw_adding_kernel": "solver_sw", "sw_flux_kernel": "solver_sw", "sw_scan_kernel": "solver_sw",
            "sp_sw_layer_kernel": "solver_sw", "sp_sw_sweep_kernel": "solver_sw", "sp_lw_layer_kernel": "solver_lw", "sp_lw_sweep_kernel": "solver_lw",
            "tc_prep_kernel": "cloud_optics_generator", "tc_sw_kernel": "solver_sw", "tc_lw_kernel": "solver_lw"}


def read_raw(path):
    r = list(csv.reader(open(path)))
    hdr, units, vals = r[0], r[1], r[2]
    return {h: (vals[i], units[i]) for i, h in enumerate(hdr)}


def num(d, key):
    v, u = d[key]
    x = float(v.replace(",", ""))
    scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1e-3, "ms": 1.0, "ns": 1e-6, "s": 1e3}.get(u, 1.0)
    return x * scale


def main():
    tag, ncol = sys.argv[1], int(sys.argv[2])
    key = sys.argv[3] if len(sys.argv) > 3 else None
    files = sorted(glob.glob(os.path.join(ROOT, "gpurun_out", f"{tag}_*_raw.csv")) or glob.glob(os.path.join(ROOT, "profiles", f"{tag}_*_raw.csv")))
    kernels, stages = {}, {}
    for f in files:
        k = os.path.basename(f)[len(tag) + 1:-len("_raw.csv")]
        d = read_raw(f)
        cyc = num(d, "sm__cycles_elapsed.avg")
        flop = cyc * (2 * num(d, "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed") +
                      num(d, "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed") +
                      num(d, "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed"))
        byt = num(d, "dram__bytes_read.sum") + num(d, "dram__bytes_write.sum")
        kernels[k] = {"ms_under_ncu": num(d, "gpu__time_duration.sum"), "dram_bytes_per_column": byt / ncol, "fp64_flop_per_column": flop / ncol,
                      "dfma_share_of_fp64_arith": 0.0}
        dfma = num(d, "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed")
        tot = dfma + num(d, "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed") + num(d, "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed")
        kernels[k]["dfma_share_of_fp64_arith"] = dfma / tot if tot else 0.0
        st = STAGE_OF.get(k)
        if st:
            s = stages.setdefault(st, {"dram_bytes_per_column": 0.0, "fp64_flop_per_column": 0.0, "kernels": []})
            s["dram_bytes_per_column"] += byt / ncol
            s["fp64_flop_per_column"] += flop / ncol
            s["kernels"].append(k)
    for s in stages.values():
        s["source"] = f"ncu --set full, one {ncol}-column tile, captures {tag}_<kernel>_raw.csv under profiles/"
    path = os.path.join(ROOT, "profiles", "traffic.json")
    old = json.load(open(path)) if os.path.exists(path) else {}
    out = dict(stages, kernels=kernels)
    if key:
        old[key] = out
    else:
        keep = {k: v for k, v in old.items() if k in ("spartacus_rrtmg", "tripleclouds_ecckd64")}
        old = dict(out, **keep)
    json.dump(old, open(path, "w"), indent=1)
    tot_b = sum(s["dram_bytes_per_column"] for s in stages.values())
    tot_f = sum(s["fp64_flop_per_column"] for s in stages.values())
    print(f"{len(files)} captures: {tot_b / 1e6:.3f} MB DRAM traffic and {tot_f / 1e6:.2f} MFLOP (fp64) per column")
    for k, v in sorted(kernels.items(), key=lambda kv: -kv[1]["ms_under_ncu"]):
        print(f"  {k:28s} {v['ms_under_ncu']:7.3f} ms  {v['dram_bytes_per_column'] / 1e3:9.1f} kB/col  {v['fp64_flop_per_column'] / 1e6:7.3f} MFLOP/col  DFMA share {100 * v['dfma_share_of_fp64_arith']:.0f} %")


if __name__ == "__main__":
    main()
